#!/usr/bin/env bash
# Round 2, GPU call U (1 GPU): the Mersenne-Twister stream, ziggurat deviates and the deck load on the device against the oracle
set -u
mkdir -p gpurun_out
S=gpurun_out/r2u_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_mt.py -q -m gpu -p no:cacheprovider --timeout=300 -rfEs > gpurun_out/r2u_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2u_pytest.log | tail -20 | tee -a $S
tail -80 gpurun_out/r2u_pytest.log > gpurun_out/r2u_pytest_tail.txt
