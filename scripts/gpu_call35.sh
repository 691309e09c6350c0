#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
for g in 0 32 64 128; do
  VPB_L2_FETCH_GRANULARITY=$g timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b35_g$g.json 2> $O/b35_g$g.err
  python - <<PY
import json
d=json.loads(open("$O/b35_g$g.json").read().strip().splitlines()[-1])
l=d["advance_p_ms_by_launch"]
print("gran $g ->", d["l2_fetch_granularity_bytes"], "value %.3e ms/step %.2f avg %.2f fresh %.2f last %.2f sort %.2f clk %s"%(d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], min(l), max(l), d["breakdown_ms_per_step"]["sort_p"], d["clocks"]["sm_mhz"]))
PY
done
