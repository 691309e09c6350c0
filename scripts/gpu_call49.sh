#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu49.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu49.log
tail -4 $O/pytest_gpu49.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $O/b49_default.json 2> $O/b49_default.err; echo "exit $?" >> $O/b49_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b49_default.json").read().strip().splitlines()[-1])
l=d["advance_p_ms_by_launch"]
print("default value %.4e ms/step %.2f avg %.2f frac %.3f min %.2f max %.2f sort %.2f clk %s launches %d"%(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], min(l), max(l), d["breakdown_ms_per_step"]["sort_p"], d["clocks"]["sm_mhz"], d["gpu_launches"]))
print("e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["config"].get("sort_key"))
PY
