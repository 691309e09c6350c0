#!/usr/bin/env bash
# weak-scaling bench at N ranks (N = $1), one rank per GPU
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
N=$1
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py > $O/bfinal_n1.json 2> $O/bfinal_n1.err; echo "exit $?" >> $O/bfinal_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29570 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e > $O/bfinal_n$N.json 2> $O/bfinal_n$N.err; echo "exit $?" >> $O/bfinal_n$N.err
fi
python - <<PY
import json
d=json.loads(open("gpurun_out/bfinal_n$N.json").read().strip().splitlines()[-1])
print("N=$N value %.4e ms/step %.2f frac %.3f avg %.2f"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"]), d["breakdown_ms_per_step"], d["clocks"])
if "e2e" in d: print(" e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline",{}).get("value"))
PY
