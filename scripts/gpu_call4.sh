#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu4.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu4.log
for v in "1 16" "0 16" "1 8" "1 32"; do
  set -- $v
  VPB_ADVANCE_P_ORDERED=$1 VPB_ADVANCE_P_BY=$2 timeout 600 python bench.py --cells 128 --ppc 64 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $O/b4_128_o$1by$2.json 2> $O/b4_128_o$1by$2.err
done
for v in "1 16" "1 8" "1 4"; do
  set -- $v
  VPB_ADVANCE_P_ORDERED=$1 VPB_ADVANCE_P_BY=$2 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/b4_256_o$1by$2.json 2> $O/b4_256_o$1by$2.err
done
if timeout 300 python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/plain4.log 2>&1; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_kernel -s 36 -c 1 -o $O/prof_advance_p_r1d_256_drift \
      python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full6.log 2>&1
fi
ls $O | tail -5
