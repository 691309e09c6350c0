#!/usr/bin/env bash
# Round 2, GPU call P (1 GPU): HEAD after the rho_p change -- whole GPU suite, smoke, default bench; host wall-clock trace of the
# configs[0] deck on the library as it is now (where the 1.15 ms per step go)
set -u
mkdir -p gpurun_out
S=gpurun_out/r2p_summary.txt
: > $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=600 --durations=8 -rfEs > gpurun_out/r2p_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2p_pytest.log | tail -20 | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2p_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $S
( mkdir -p /tmp/deck1 && cd /tmp/deck1 && VPB_TRACE=1 VPB_DECK_STEPS=200 timeout 300 $GRAFT_REPO_ROOT/oracle/_ref/hybrid/thermal_c1.b200.op -tpp=1 \
    > $GRAFT_REPO_ROOT/gpurun_out/r2p_deck_trace.out 2> $GRAFT_REPO_ROOT/gpurun_out/r2p_deck_trace.err )
echo "deck trace rc=$?" | tee -a $S
grep -E "simulation time|vpb trace" gpurun_out/r2p_deck_trace.out gpurun_out/r2p_deck_trace.err | tail -60 | tee -a $S
timeout 600 python bench.py > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench_default.err
echo "bench default rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
d = json.loads([l for l in open("gpurun_out/r2p_bench_default.json") if l.startswith("{")][-1])
print("ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "frac", d["roofline"]["frac"], "e2e", d.get("e2e"))
print("breakdown", d.get("breakdown_ms_per_step"))
print("div_clean", d.get("div_clean"))
print("deck_e2e", json.dumps(d.get("deck_e2e"))[:1500])
PY
