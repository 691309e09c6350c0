#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_particles.py -m gpu -q -x -n 4 -p no:cacheprovider -k "sort" > $O/pytest_gpu22_sort.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu22_sort.log
tail -3 $O/pytest_gpu22_sort.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b22.json 2> $O/b22.err
python - <<PY
import json
d=json.loads(open("$O/b22.json").read().strip().splitlines()[-1])
l=d["advance_p_ms_by_launch"]
print("value %.3e ms/step %.2f avg %.2f fresh %.2f last %.2f sort %.2f clk %s"%(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], min(l), max(l), d["breakdown_ms_per_step"]["sort_p"], d["clocks"]["sm_mhz"]))
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_pair -s 2 -c 1 -o $O/prof_advance_p_r1o_256_step1_pair4 \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --field-cells 0 > $O/ncu_full22.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sort_ -c 40 --csv --log-file $O/launches_r1p_sort.csv python bench.py --steps 19 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/ncu_launches22.log 2>&1
grep sort_ $O/launches_r1p_sort.csv | awk -F'","' '{print $5, $NF}' | tail -8
