#!/usr/bin/env bash
# Round 2, GPU call D (1 GPU): second version of the grouped sort's move kernel (bulk async loads, 3 CTAs/SM), new tests
set -u
mkdir -p gpurun_out
S=gpurun_out/r2d_summary.txt
: > $S
timeout 900 python -m pytest tests/test_gpu_particles.py tests/test_gpu_diag.py tests/test_gpu_step.py tests/test_gpu_boundary.py tests/test_gpu_harris.py tests/test_gpu_deck.py -q -m gpu -p no:cacheprovider --timeout=600 -rfEs -k "not trecon_part_deck_as_shipped and not ranks" > gpurun_out/r2d_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2d_pytest.log | tail -30 | tee -a $S
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
for si in 20 10; do
  for ev in 0 1; do
    VPB_SORT_EVICT_LAST=$ev timeout 300 python bench.py --steps $((2 * si)) --warmup 3 $B --sort-interval $si > gpurun_out/r2d_bench_${si}_ev$ev.json 2> gpurun_out/r2d_bench_${si}_ev$ev.err
    echo "bench interval $si evict_last $ev rc=$?" | tee -a $S
  done
done
VPB_SORT_GROUPED=0 timeout 300 python bench.py --steps 20 --warmup 3 $B --sort-interval 10 > gpurun_out/r2d_bench_old_10.json 2> gpurun_out/r2d_bench_old_10.err
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2d_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "avg_launch %.2f" % d["roofline"]["avg_launch_ms"],
              "sort ms %.1f frac %.3f" % (d["sort_p"]["ms_per_sort"], d["sort_p"]["frac"]), "sort/step %.2f" % d["breakdown_ms_per_step"]["sort_p"])
    except Exception as e:
        print(f, "failed", e)
PY
P="--steps 9 --warmup 3 $B --sort-interval 10"
python bench.py $P > gpurun_out/r2d_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'group_move' -s 2 -c 1 -o gpurun_out/r2d_prof -f python bench.py $P > gpurun_out/r2d_ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $S
