#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_boundary.py -m gpu -q -x -p no:cacheprovider > $O/pytest_gpu40.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu40.log; tail -3 $O/pytest_gpu40.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/dist_gpu_worker.py > $O/dist2_40.log 2>&1; echo "exit $?" >> $O/dist2_40.log
grep -E "DIST_GPU|Error|exit" $O/dist2_40.log | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > $O/b40_n2_256.json 2> $O/b40_n2_256.err; echo "exit $?" >> $O/b40_n2_256.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b40_n2_256.json").read().strip().splitlines()[-1])
print("N=2 value %.3e ms/step %.2f"%(d["value"], d["ms_per_step"]), d["breakdown_ms_per_step"], d.get("driver"))
PY
tail -2 $O/b40_n2_256.err
