#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
if timeout 300 python bench.py --steps 11 --warmup 1 --no-e2e --no-cpu-baseline --field-cells 0 > $O/plain12.log 2>&1; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_stream -s 20 -c 1 -o $O/prof_advance_p_r1k_256_step10_c2 \
      python bench.py --steps 11 --warmup 1 --no-e2e --no-cpu-baseline --field-cells 0 > $O/ncu_full16.log 2>&1
  VPB_ADVANCE_P_STREAM_STORE=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_stream -s 20 -c 1 -o $O/prof_advance_p_r1k_256_step10_c2s1 \
      python bench.py --steps 11 --warmup 1 --no-e2e --no-cpu-baseline --field-cells 0 > $O/ncu_full17.log 2>&1
fi
ls $O | tail -3
