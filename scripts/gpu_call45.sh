#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu45.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu45.log
tail -6 $O/pytest_gpu45.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke45.log 2>&1; tail -2 $O/smoke45.log
