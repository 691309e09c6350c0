#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1200 python bench.py > $O/b42_default.json 2> $O/b42_default.err; echo "exit $?" >> $O/b42_default.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_r1r.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/ncu_launches42.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b42_default.json").read().strip().splitlines()[-1])
l=d["advance_p_ms_by_launch"]
print("default value %.4e ms/step %.2f avg %.2f frac %.3f fresh %.2f last %.2f sort %.2f clk %s launches %d"%(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], min(l), max(l), d["breakdown_ms_per_step"]["sort_p"], d["clocks"], d["gpu_launches"]))
print("e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
f=d["fields_c2"]; print("fields", f["ms_per_step"], f["advance_b"]["frac"], f["advance_e"]["frac"])
PY
