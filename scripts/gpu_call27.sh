#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_particles.py -m gpu -q -x -n 4 -p no:cacheprovider -k "pair or planes" > $O/pytest_gpu27_pair.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu27_pair.log
tail -3 $O/pytest_gpu27_pair.log
run() { # name cps pipe
  VPB_ADVANCE_P_PAIR_CPS=$2 VPB_ADVANCE_P_PAIR_PIPE=$3 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b27_$1.json 2> $O/b27_$1.err
  python - <<PY
import json
d=json.loads(open("$O/b27_$1.json").read().strip().splitlines()[-1])
l=d["advance_p_ms_by_launch"]
print("$1", "value %.3e ms/step %.2f avg %.2f fresh %.2f last %.2f sort %.2f clk %s"%(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], min(l), max(l), d["breakdown_ms_per_step"]["sort_p"], d["clocks"]["sm_mhz"]))
print(l)
PY
}
run c4pipe1 4 1
run c3pipe1 3 1
