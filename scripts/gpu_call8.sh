#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu8.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu8.log
timeout 600 python bench.py --cells 128 --ppc 64 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $O/b8_128.json 2> $O/b8_128.err
for c in 3 2; do
  VPB_ADVANCE_P_STREAM_CTAS_PER_SM=$c timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/b8_256_c$c.json 2> $O/b8_256_c$c.err
done
if timeout 300 python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/plain8.log 2>&1; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_stream -s 2 -c 1 -o $O/prof_advance_p_r1h_256_fresh \
      python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full12.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_stream -s 36 -c 1 -o $O/prof_advance_p_r1h_256_drift \
      python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full13.log 2>&1
fi
ls $O | tail -3
