#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_particles.py tests/test_gpu_step.py tests/test_golden.py -m gpu -q -x -n 4 -p no:cacheprovider -k "sort or lookahead or native or golden" > $O/pytest_gpu48.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu48.log
tail -5 $O/pytest_gpu48.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b48.json 2> $O/b48.err
python - <<PY
import json
d=json.loads(open("$O/b48.json").read().strip().splitlines()[-1])
l=d["advance_p_ms_by_launch"]
print("value %.3e ms/step %.2f avg %.2f frac %.3f sort %.2f clk %s"%(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["breakdown_ms_per_step"]["sort_p"], d["clocks"]["sm_mhz"]))
PY
