#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_particles.py tests/test_gpu_step.py tests/test_golden.py -m gpu -q -x -n 4 -p no:cacheprovider -k "sort or lookahead or native or golden or harris" > $O/pytest_gpu48.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu48.log
tail -5 $O/pytest_gpu48.log
