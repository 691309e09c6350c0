#!/usr/bin/env bash
# Round 2, round-end sequence (1 GPU; outputs named r2y): the round-end sequence at the head -- whole GPU suite, smoke, default bench (both arms), ncu
# launch list of the bench command
set -u
mkdir -p gpurun_out
S=gpurun_out/r2y_summary.txt
: > $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=600 --durations=8 -rfEs > gpurun_out/r2y_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2y_pytest.log | tail -20 | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2y_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $S
SECONDS=0
timeout 600 python bench.py > gpurun_out/r2y_bench_default.json 2> gpurun_out/r2y_bench_default.err
echo "bench default rc=$? (${SECONDS} s)" | tee -a $S
SECONDS=0
timeout 400 python bench.py --impl reference > gpurun_out/r2y_bench_reference.json 2> gpurun_out/r2y_bench_reference.err
echo "bench reference rc=$? (${SECONDS} s)" | tee -a $S
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2y_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "frac", (d.get("roofline") or {}).get("frac"),
              "sort", (d.get("sort_p") or {}).get("ms_per_sort"), "e2e", (d.get("e2e") or {}).get("value"))
        for k in ("breakdown_ms_per_step", "small_step", "load_mt", "cpu_baseline"):
            if k in d: print("   ", k, json.dumps(d[k])[:700])
        if "fields_c2" in d:
            print("    fields_c2", {k: (round(v["avg_launch_ms"], 3), round(v["frac"], 3)) for k, v in d["fields_c2"].items() if isinstance(v, dict) and "frac" in v})
    except Exception as e:
        print(f, "failed", e)
PY
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
python bench.py --steps 7 --warmup 3 $B > gpurun_out/r2y_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2y_launches.csv python bench.py --steps 7 --warmup 3 $B > gpurun_out/r2y_ncu_list.log 2>&1
echo "ncu list rc=$?" | tee -a $S
