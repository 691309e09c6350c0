#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu3.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu3.log
for v in "1 0 0" "1 1 0" "1 0 1" "1 1 1" "0 1 1"; do
  set -- $v
  VPB_ADVANCE_P_DEPOSIT=$1 VPB_ADVANCE_P_L2HINT=$2 VPB_ADVANCE_P_PREFETCH=$3 timeout 600 python bench.py --cells 128 --ppc 64 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $O/b3_128_d$1h$2p$3.json 2> $O/b3_128_d$1h$2p$3.err
done
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/b3_256.json 2> $O/b3_256.err; echo "exit $?" >> $O/b3_256.err
VPB_ADVANCE_P_L2HINT=0 VPB_ADVANCE_P_PREFETCH=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/b3_256_plain.json 2> $O/b3_256_plain.err
if timeout 300 python bench.py --cells 128 --ppc 64 --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/plain3.log 2>&1; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:advance_p_kernel -s 2 -c 1 -o $O/prof_advance_p_r1c \
      python bench.py --cells 128 --ppc 64 --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full4.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:advance_p_kernel -s 36 -c 1 -o $O/prof_advance_p_r1c_drift \
      python bench.py --cells 128 --ppc 64 --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full5.log 2>&1
fi
ls $O | tail -5
