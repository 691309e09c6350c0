#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu6.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu6.log
for t in 1 0; do
  VPB_ADVANCE_P_TMA=$t timeout 600 python bench.py --cells 128 --ppc 64 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $O/b6_128_t$t.json 2> $O/b6_128_t$t.err
done
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/b6_256_t1.json 2> $O/b6_256_t1.err
VPB_ADVANCE_P_TMA_CTAS_PER_SM=2 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/b6_256_t1c2.json 2> $O/b6_256_t1c2.err
if timeout 300 python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/plain6.log 2>&1; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_tma -s 2 -c 1 -o $O/prof_advance_p_r1f_256_fresh \
      python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full9.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_tma -s 36 -c 1 -o $O/prof_advance_p_r1f_256_drift \
      python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full10.log 2>&1
fi
ls $O | tail -3
