#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu14.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu14.log
VPB_DROPIN_PLANAR=1 timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu14_planar.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu14_planar.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b14_256_wide.json 2> $O/b14_256_wide.err
VPB_SIM_NARROW_INTERPOLATOR=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b14_256_narrow.json 2> $O/b14_256_narrow.err
VPB_ADVANCE_P_STREAM_STORE=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b14_256_wide_s1.json 2> $O/b14_256_wide_s1.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_r1m.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/ncu_launches14.log 2>&1
ls $O | tail -3
