#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_particles.py -m gpu -q -x -n 4 -p no:cacheprovider -k "pair or planes" > $O/pytest_gpu21_pair.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu21_pair.log
tail -5 $O/pytest_gpu21_pair.log
run() { # name cps pf dedup
  VPB_ADVANCE_P_PAIR_CPS=$2 VPB_ADVANCE_P_PAIR_PREFETCH=$3 VPB_ADVANCE_P_PAIR_DEDUP=$4 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b21_$1.json 2> $O/b21_$1.err
  python - <<PY
import json
d=json.loads(open("$O/b21_$1.json").read().strip().splitlines()[-1])
l=d["advance_p_ms_by_launch"]
print("$1", "value %.3e ms/step %.2f avg %.2f fresh %.2f last %.2f sort %.2f clk %s"%(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], min(l), max(l), d["breakdown_ms_per_step"]["sort_p"], d["clocks"]["sm_mhz"]))
PY
}
run c4p1d1 4 1 1
run c4p2d1 4 2 1
run c5p1d1 5 1 1
run c3p2d1 3 2 1
run c4p1d0 4 1 0
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_r1o.csv python bench.py --steps 19 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/ncu_launches21.log 2>&1
