#!/usr/bin/env bash
# Round 2, GPU call A (1 GPU):   gpurun --timeout 2100 -- 'bash scripts/gpu_r2a.sh'
# 1. host wall-clock trace of BASELINE configs[0] as an unmodified reference program on the library (where do 12 ms/step go)
# 2. the whole GPU suite with the gated tests on
# 3. bench A/B: advance_p_pair variants, sort.scatter, e2e hot_only, deck-e2e, harris
set -u
mkdir -p gpurun_out
S=gpurun_out/r2a_summary.txt
: > $S
nvidia-smi -L | tee -a $S
( mkdir -p /tmp/deck1 && cd /tmp/deck1 && VPB_TRACE=1 VPB_DECK_STEPS=40 timeout 300 $GRAFT_REPO_ROOT/oracle/_ref/hybrid/thermal_c1.b200.op -tpp=1 \
    > $GRAFT_REPO_ROOT/gpurun_out/r2a_deck_trace.out 2> $GRAFT_REPO_ROOT/gpurun_out/r2a_deck_trace.err )
echo "deck trace rc=$?" | tee -a $S
grep -E "simulation time|vpb trace" gpurun_out/r2a_deck_trace.out gpurun_out/r2a_deck_trace.err | tail -60 | tee -a $S
( mkdir -p /tmp/deck2 && cd /tmp/deck2 && VPB_TRACE=1 VPB_DROPIN_PREFETCH=0 VPB_DECK_STEPS=40 timeout 300 $GRAFT_REPO_ROOT/oracle/_ref/hybrid/thermal_c1.b200.op -tpp=1 \
    > $GRAFT_REPO_ROOT/gpurun_out/r2a_deck_trace_nopf.out 2> $GRAFT_REPO_ROOT/gpurun_out/r2a_deck_trace_nopf.err )
echo "deck trace (no prefetch) rc=$?" | tee -a $S
grep -E "simulation time|vpb trace" gpurun_out/r2a_deck_trace_nopf.out gpurun_out/r2a_deck_trace_nopf.err | tail -60 | tee -a $S

export VPB_RUN_UNVALIDATED=1
timeout 1300 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=700 --durations=40 -rfE > gpurun_out/r2a_all_pytest.log 2>&1
echo "full gpu pytest (gated included) rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2a_all_pytest.log | tail -60 | tee -a $S
unset VPB_RUN_UNVALIDATED

for v in 1 0 3; do
  VPB_ADVANCE_P_PAIR_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e \
    > gpurun_out/r2a_bench_variant$v.json 2> gpurun_out/r2a_bench_variant$v.err
  echo "bench variant $v rc=$?" | tee -a $S
done
VPB_SORT_SCATTER=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e \
  > gpurun_out/r2a_bench_sort_scatter.json 2> gpurun_out/r2a_bench_sort_scatter.err
echo "bench sort.scatter rc=$?" | tee -a $S
VPB_DROPIN_HOT_ONLY=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --field-cells 0 --no-deck-e2e \
  > gpurun_out/r2a_bench_e2e_hot_only.json 2> gpurun_out/r2a_bench_e2e_hot_only.err
echo "bench e2e hot_only rc=$?" | tee -a $S
timeout 400 python bench.py --workload harris --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-deck-e2e > gpurun_out/r2a_bench_harris.json 2> gpurun_out/r2a_bench_harris.err
echo "bench harris rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2a_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "roofline", d.get("roofline", {}).get("frac"),
              "avg_launch", d.get("roofline", {}).get("avg_launch_ms"), "breakdown", d.get("breakdown_ms_per_step"), "e2e", d.get("e2e"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -n 5 gpurun_out/r2a_all_pytest.log
