#!/usr/bin/env bash
# Round 2, after gpu_r2_first.sh has told which advance_p_pair variant and which sort form win:
#   gpurun --timeout 1800 -- 'VARIANT=3 SCATTER=1 bash scripts/gpu_r2_sweep.sh'
# sweeps the species sort interval and the look-ahead of the sort key (steps = two sort periods, so every run sees two
# sorts per species); one JSON line per point in gpurun_out/r2_sweep_<interval>_<lookahead>.json
set -u
mkdir -p gpurun_out
export VPB_ADVANCE_P_PAIR_VARIANT=${VARIANT:-0} VPB_SORT_SCATTER=${SCATTER:-0}
for si in 10 14 20; do
  for la in $((si / 2)) $((si * 6 / 10)) $((si * 7 / 10)); do
    python bench.py --steps $((2 * si)) --warmup 3 --no-e2e --no-cpu-baseline --field-cells 0 --sort-interval $si --sort-lookahead $la \
      > gpurun_out/r2_sweep_${si}_${la}.json 2> gpurun_out/r2_sweep_${si}_${la}.err
    python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_sweep_${si}_${la}.json") if l.startswith("{")][-1])
    print("interval %2d lookahead %2d: %.2f ms/step, advance_p %.2f ms/launch (frac %.3f), sort %.2f ms/step" % (
        $si, $la, d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["breakdown_ms_per_step"]["sort_p"]))
except Exception as e:
    print("interval $si lookahead $la: failed", e)
PY
  done
done | tee gpurun_out/r2_sweep_summary.txt
