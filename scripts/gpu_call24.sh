#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu24.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu24.log
tail -15 $O/pytest_gpu24.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke24.log 2>&1; tail -2 $O/smoke24.log
