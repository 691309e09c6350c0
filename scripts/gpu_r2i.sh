#!/usr/bin/env bash
# Round 2, GPU call I (1 GPU): key and invert passes with several rows per thread in flight
set -u
mkdir -p gpurun_out
S=gpurun_out/r2i_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_particles.py tests/test_gpu_step.py -q -m gpu -p no:cacheprovider --timeout=600 -rfEs -k "sort or graph or native" > gpurun_out/r2i_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2i_pytest.log | tail -30 | tee -a $S
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
run() {
  local name=$1; shift
  timeout 300 env "$@" > gpurun_out/r2i_bench_$name.json 2> gpurun_out/r2i_bench_$name.err
  echo "bench $name rc=$?" | tee -a $S
}
run i10 python bench.py --steps 20 --warmup 3 $B --sort-interval 10
run i5 python bench.py --steps 20 --warmup 3 $B --sort-interval 5
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2i_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "avg_launch %.2f" % d["roofline"]["avg_launch_ms"],
              "minmax %.2f %.2f" % (d["roofline"]["min_launch_ms"], d["roofline"]["max_launch_ms"]),
              "sort ms %.1f frac %.3f" % (d["sort_p"]["ms_per_sort"], d["sort_p"]["frac"]), "sort/step %.2f" % d["breakdown_ms_per_step"]["sort_p"])
    except Exception as e:
        print(f, "failed", e)
PY
P="--steps 9 --warmup 3 $B --sort-interval 5"
python bench.py $P > gpurun_out/r2i_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'group_' -s 6 -c 3 -o gpurun_out/r2i_prof_keys_invert -f python bench.py $P > gpurun_out/r2i_ncu_full.log 2>&1
echo "ncu keys+invert rc=$?" | tee -a $S
