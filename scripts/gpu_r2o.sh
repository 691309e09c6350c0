#!/usr/bin/env bash
# Round 2, GPU call O (1 GPU): rho_p with per-voxel warp aggregation (parity + cost of a cleaning step)
set -u
mkdir -p gpurun_out
S=gpurun_out/r2o_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_particles.py tests/test_gpu_fields.py tests/test_gpu_step.py tests/test_gpu_history.py tests/test_gpu_aniso.py tests/test_gpu_fuzz.py tests/test_golden.py tests/test_gpu_harris.py -q -m gpu -p no:cacheprovider --timeout=600 -rfEs > gpurun_out/r2o_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2o_pytest.log | tail -10 | tee -a $S
timeout 300 python bench.py --no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e > gpurun_out/r2o_bench_default.json 2> gpurun_out/r2o_bench_default.err
echo "bench rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
d = json.loads([l for l in open("gpurun_out/r2o_bench_default.json") if l.startswith("{")][-1])
print("ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "frac", d["roofline"]["frac"], "div_clean", d["div_clean"])
PY
