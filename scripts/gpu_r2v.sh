#!/usr/bin/env bash
# Round 2, GPU call V (1 GPU): device-resident run from the seed alone against the reference deck's energies; load time of configs[0]
set -u
mkdir -p gpurun_out
S=gpurun_out/r2v_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_mt.py tests/test_gpu_step.py -q -m gpu -p no:cacheprovider --timeout=300 -rfEs > gpurun_out/r2v_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2v_pytest.log | tail -20 | tee -a $S
tail -80 gpurun_out/r2v_pytest.log > gpurun_out/r2v_pytest_tail.txt
timeout 300 python - > gpurun_out/r2v_load_mt.json 2> gpurun_out/r2v_load_mt.err <<'PY'
import json
import bench
from old_vpic_b200 import lib
L = lib.load()
L.vpb_init(0)
print(json.dumps(bench.load_mt_measure(L, True)))
PY
echo "load_mt rc=$?" | tee -a $S
cat gpurun_out/r2v_load_mt.json | tee -a $S
tail -5 gpurun_out/r2v_load_mt.err | tee -a $S
