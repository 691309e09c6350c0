#!/usr/bin/env bash
# 8-GPU pass: multi-rank functional test at 2x2x2, weak-scaling bench at N=8 and N=4
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tests/dist_gpu_worker.py > $O/dist8_43.log 2>&1; echo "exit $?" >> $O/dist8_43.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e > $O/b43_n8_256.json 2> $O/b43_n8_256.err; echo "exit $?" >> $O/b43_n8_256.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > $O/b43_n4_256.json 2> $O/b43_n4_256.err; echo "exit $?" >> $O/b43_n4_256.err
tail -3 $O/dist8_43.log
python - <<'PY'
import json
for f in ["n8","n4"]:
    d=json.loads(open("gpurun_out/b43_%s_256.json"%f).read().strip().splitlines()[-1])
    print(f, "value %.3e ms/step %.2f"%(d["value"], d["ms_per_step"]), d["breakdown_ms_per_step"])
PY
