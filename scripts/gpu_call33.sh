#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_particles.py tests/test_gpu_history.py -m gpu -q -x -n 4 -p no:cacheprovider -k "sort or history" > $O/pytest_gpu33_sort.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu33_sort.log
tail -2 $O/pytest_gpu33_sort.log
bash scripts/gpu_call32.sh 2>&1 | grep -E "^1[0-4] "
