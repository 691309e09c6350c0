#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
export VPB_ADVANCE_P_PAIR_CPS=4
if timeout 300 python bench.py --steps 11 --warmup 1 --no-e2e --no-cpu-baseline --field-cells 0 > $O/plain20.log 2>&1; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_pair -s 20 -c 1 -o $O/prof_advance_p_r1n_256_step10_pair4 \
      python bench.py --steps 11 --warmup 1 --no-e2e --no-cpu-baseline --field-cells 0 > $O/ncu_full20.log 2>&1
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_r1n.csv python bench.py --steps 6 --warmup 1 --no-cpu-baseline --no-e2e --field-cells 0 > $O/ncu_launches20.log 2>&1
fi
ls $O | tail -3
