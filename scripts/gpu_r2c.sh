#!/usr/bin/env bash
# Round 2, GPU call C (1 GPU):   gpurun --timeout 2400 -- 'bash scripts/gpu_r2c.sh'
# grouped (brick-Morton) sort: parity tests, bench A/B against the round-1 pipeline, sort-interval sweep, the new bench
# line end to end, ncu launch list + full capture of the sort kernels and of advance_p right after a sort
set -u
mkdir -p gpurun_out
S=gpurun_out/r2c_summary.txt
: > $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=600 -rfEs -k "not trecon and not ranks and not multi" > gpurun_out/r2c_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2c_pytest.log | tail -30 | tee -a $S
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
VPB_SORT_GROUPED=0 timeout 300 python bench.py --steps 20 --warmup 3 $B > gpurun_out/r2c_bench_old_sort.json 2> gpurun_out/r2c_bench_old_sort.err
echo "bench old sort rc=$?" | tee -a $S
for si in 20 10 8 5; do
  timeout 300 python bench.py --steps $((2 * si)) --warmup 3 $B --sort-interval $si > gpurun_out/r2c_bench_grouped_$si.json 2> gpurun_out/r2c_bench_grouped_$si.err
  echo "bench grouped interval $si rc=$?" | tee -a $S
done
timeout 300 python bench.py --steps 20 --warmup 3 $B --sort-interval 10 --sort-lookahead 5 > gpurun_out/r2c_bench_grouped_10_5.json 2> gpurun_out/r2c_bench_grouped_10_5.err
timeout 300 python bench.py --steps 20 --warmup 3 $B --sort-interval 10 --sort-lookahead 0 > gpurun_out/r2c_bench_grouped_10_0.json 2> gpurun_out/r2c_bench_grouped_10_0.err
timeout 900 python bench.py > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
echo "bench default rc=$?" | tee -a $S
timeout 300 python bench.py --impl reference > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err
echo "bench reference rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "frac", d.get("roofline", {}).get("frac"),
              "avg_launch", d.get("roofline", {}).get("avg_launch_ms"), "minmax", d.get("roofline", {}).get("min_launch_ms"), d.get("roofline", {}).get("max_launch_ms"),
              "sort", d.get("sort_p"), "breakdown", d.get("breakdown_ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "divclean", d.get("div_clean"))
        if "fields_c2" in d: print("   fields", {k: (v.get("frac") if isinstance(v, dict) else v) for k, v in d["fields_c2"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
# ncu: launch list of a short run, then the sort kernels of step 10 and the advance_p launch that follows
P="--steps 9 --warmup 3 $B --sort-interval 10"
python bench.py $P > gpurun_out/r2c_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py $P > gpurun_out/r2c_ncu_list.log 2>&1
echo "ncu list rc=$?" | tee -a $S
python bench.py $P > gpurun_out/r2c_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'advance_p_pair|group_move|group_keys' -s 24 -c 5 -o gpurun_out/r2c_prof -f python bench.py $P > gpurun_out/r2c_ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $S
ls -la gpurun_out/*.ncu-rep | tee -a $S
