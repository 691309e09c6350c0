#!/usr/bin/env bash
# Round 2, GPU call Q (1 GPU): fused migration rounds through ranks sharing the GPU (deck tests), boundary/step tests
set -u
mkdir -p gpurun_out
S=gpurun_out/r2q_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_deck.py tests/test_gpu_boundary.py tests/test_gpu_step.py -q -m gpu -p no:cacheprovider --timeout=600 -rfEs \
  -k "on_ranks or boundary or native_driver or wall_decks or grows" > gpurun_out/r2q_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2q_pytest.log | tail -20 | tee -a $S
tail -60 gpurun_out/r2q_pytest.log > gpurun_out/r2q_pytest_tail.txt
