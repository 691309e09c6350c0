#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu13.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu13.log
VPB_DROPIN_PLANAR=1 timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu13_planar.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu13_planar.log
timeout 600 python bench.py --workload fields --field-cells 1024 > $O/b13_fields1024.json 2> $O/b13_fields1024.err
timeout 600 python bench.py --workload fields --field-cells 512 > $O/b13_fields512.json 2> $O/b13_fields512.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b13_256.json 2> $O/b13_256.err
if timeout 300 python bench.py --workload fields --field-cells 512 > $O/plain13.log 2>&1; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'advance_b_kernel|advance_e_kernel' -s 6 -c 3 -o $O/prof_fields_r1l_512_planar \
      python bench.py --workload fields --field-cells 512 > $O/ncu_full18.log 2>&1
fi
ls $O | tail -3
