#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_pair -s 20 -c 1 -o $O/prof_advance_p_r1p_256_step10_pair4pipe \
      python bench.py --steps 11 --warmup 1 --no-e2e --no-cpu-baseline --field-cells 0 > $O/ncu_full26.log 2>&1
ls -la $O | tail -3
