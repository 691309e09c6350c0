#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu17.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu17.log
for c in 5 6 4; do
  VPB_ADVANCE_P_STREAM_CPS=$c timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b17_256_cps$c.json 2> $O/b17_256_cps$c.err
done
ls $O | tail -3
