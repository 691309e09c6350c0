#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_history.py -m gpu -q -x -n 4 -p no:cacheprovider > $O/pytest_gpu39.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu39.log
tail -12 $O/pytest_gpu39.log
timeout 1200 python bench.py --no-cpu-baseline > $O/b39_native.json 2> $O/b39_native.err; echo "exit $?" >> $O/b39_native.err
timeout 600 python bench.py --driver python --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b39_python.json 2> $O/b39_python.err
python - <<'PY'
import json
for f in ["native","python"]:
    try:
        d=json.loads(open("gpurun_out/b39_%s.json"%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/b39_%s.err"%f).read()[-800:]); continue
    print(f, "value %.3e ms/step %.2f host %.2f avg %.2f frac %.3f sort %.2f launches %s"%(d["value"], d["ms_per_step"], d["host_wall_ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["breakdown_ms_per_step"]["sort_p"], d["gpu_launches"]))
    if "fields_c2" in d: print("  fields", d["fields_c2"]["ms_per_step"], d["fields_c2"]["advance_b"]["frac"], d["fields_c2"]["advance_e"]["frac"], d["fields_c2"]["em_energy_drift_rel"])
PY
