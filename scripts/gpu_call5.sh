#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu5.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu5.log
for o in 1 0; do
  VPB_ADVANCE_P_ORDERED=$o timeout 600 python bench.py --cells 128 --ppc 64 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $O/b5_128_o$o.json 2> $O/b5_128_o$o.err
  VPB_ADVANCE_P_ORDERED=$o timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/b5_256_o$o.json 2> $O/b5_256_o$o.err
done
VPB_ADVANCE_P_DEPOSIT=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/b5_256_dep0.json 2> $O/b5_256_dep0.err
timeout 900 python bench.py --steps 20 --warmup 3 > $O/b5_256_full.json 2> $O/b5_256_full.err
if timeout 300 python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/plain5.log 2>&1; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_kernel -s 2 -c 1 -o $O/prof_advance_p_r1e_256_fresh \
      python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full7.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_kernel -s 36 -c 1 -o $O/prof_advance_p_r1e_256_drift \
      python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full8.log 2>&1
fi
ls $O | tail -3
