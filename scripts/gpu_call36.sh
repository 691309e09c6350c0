#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_hydro.py tests/test_abi.py -m gpu -q -x -n 4 -p no:cacheprovider > $O/pytest_gpu36_hydro.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu36_hydro.log
tail -25 $O/pytest_gpu36_hydro.log
