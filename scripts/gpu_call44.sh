#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
for la in 12 14; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 --sort-lookahead $la > $O/b44_la$la.json 2> $O/b44_la$la.err
  python - <<PY
import json
d=json.loads(open("$O/b44_la$la.json").read().strip().splitlines()[-1])
l=d["advance_p_ms_by_launch"]
print("lookahead $la value %.3e ms/step %.2f avg %.2f frac %.3f min %.2f max %.2f sort %.2f clk %s"%(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], min(l), max(l), d["breakdown_ms_per_step"]["sort_p"], d["clocks"]["sm_mhz"]))
print(l)
PY
done
