#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu38.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu38.log
tail -12 $O/pytest_gpu38.log
