#!/usr/bin/env bash
# Round 2, GPU call M (1 GPU): the round-end sequence -- whole GPU suite, smoke, default bench (both arms), configs[2]
# (thermal device load + the reference's trecon-part deck itself on the library), ncu launch list of the bench command
set -u
mkdir -p gpurun_out
S=gpurun_out/r2m_summary.txt
: > $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=600 --durations=8 -rfEs > gpurun_out/r2m_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2m_pytest.log | tail -20 | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2m_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $S
timeout 600 python bench.py > gpurun_out/r2m_bench_default.json 2> gpurun_out/r2m_bench_default.err
echo "bench default rc=$?" | tee -a $S
timeout 400 python bench.py --impl reference > gpurun_out/r2m_bench_reference.json 2> gpurun_out/r2m_bench_reference.err
echo "bench reference rc=$?" | tee -a $S
timeout 1200 python bench.py --workload harris --trecon-deck --no-e2e --no-cpu-baseline --no-deck-e2e --steps 20 --warmup 3 > gpurun_out/r2m_bench_harris.json 2> gpurun_out/r2m_bench_harris.err
echo "bench harris + trecon deck rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2m_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "frac", (d.get("roofline") or {}).get("frac"),
              "sort", (d.get("sort_p") or {}).get("ms_per_sort"), "e2e", (d.get("e2e") or {}).get("value"))
        for k in ("small_step", "trecon_deck", "trecon_deck_scaled", "cpu_baseline"):
            if k in d: print("   ", k, json.dumps(d[k])[:700])
    except Exception as e:
        print(f, "failed", e)
PY
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
python bench.py --steps 7 --warmup 3 $B > gpurun_out/r2m_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 7 --warmup 3 $B > gpurun_out/r2m_ncu_list.log 2>&1
echo "ncu list rc=$?" | tee -a $S
