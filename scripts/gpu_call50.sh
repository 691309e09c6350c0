#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_deck.py -m gpu -q -x -p no:cacheprovider > $O/pytest_gpu50.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu50.log
tail -30 $O/pytest_gpu50.log
