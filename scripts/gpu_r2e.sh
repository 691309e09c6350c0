#!/usr/bin/env bash
# Round 2, GPU call E (1 GPU): the grouped sort's move pass three ways (gather / scatter / staged chunks)
set -u
mkdir -p gpurun_out
S=gpurun_out/r2e_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_particles.py -q -m gpu -p no:cacheprovider --timeout=600 -rfEs -k "sort" > gpurun_out/r2e_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2e_pytest.log | tail -30 | tee -a $S
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
for v in 2 1; do
  VPB_SORT_GROUP_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 $B --sort-interval 10 > gpurun_out/r2e_bench_v${v}_10.json 2> gpurun_out/r2e_bench_v${v}_10.err
  echo "bench variant $v interval 10 rc=$?" | tee -a $S
done
for si in 5 8; do
  timeout 300 python bench.py --steps $((2 * si)) --warmup 3 $B --sort-interval $si > gpurun_out/r2e_bench_v2_$si.json 2> gpurun_out/r2e_bench_v2_$si.err
  echo "bench variant 2 interval $si rc=$?" | tee -a $S
done
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2e_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "avg_launch %.2f" % d["roofline"]["avg_launch_ms"],
              "sort ms %.1f frac %.3f" % (d["sort_p"]["ms_per_sort"], d["sort_p"]["frac"]), "sort/step %.2f" % d["breakdown_ms_per_step"]["sort_p"])
    except Exception as e:
        print(f, "failed", e)
PY
P="--steps 9 --warmup 3 $B --sort-interval 10"
python bench.py $P > gpurun_out/r2e_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'group_' -s 6 -c 6 -o gpurun_out/r2e_prof -f python bench.py $P > gpurun_out/r2e_ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $S
VPB_SORT_GROUP_VARIANT=1 ncu --set full --clock-control none --import-source on -k regex:'group_scatter' -s 2 -c 1 -o gpurun_out/r2e_prof_scatter -f python bench.py $P > gpurun_out/r2e_ncu_full1.log 2>&1
echo "ncu full scatter rc=$?" | tee -a $S
