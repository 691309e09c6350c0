#!/usr/bin/env bash
# Round 2, GPU call J (1 GPU): short sort intervals without / with little look-ahead (the key pass then reads 4 B per
# particle instead of 28), packed key+rank, evict_last on the gather's loads
set -u
mkdir -p gpurun_out
S=gpurun_out/r2j_summary.txt
: > $S
timeout 600 python -m pytest tests/test_gpu_particles.py -q -m gpu -p no:cacheprovider --timeout=600 -rfEs -k "sort" > gpurun_out/r2j_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2j_pytest.log | tail -30 | tee -a $S
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
run() {
  local name=$1; shift
  timeout 300 env "$@" > gpurun_out/r2j_bench_$name.json 2> gpurun_out/r2j_bench_$name.err
  echo "bench $name rc=$?" | tee -a $S
}
run i5_la3 python bench.py --steps 20 --warmup 3 $B --sort-interval 5
run i5_la0 python bench.py --steps 20 --warmup 3 $B --sort-interval 5 --sort-lookahead 0
run i5_la1 python bench.py --steps 20 --warmup 3 $B --sort-interval 5 --sort-lookahead 1
run i6_la0 python bench.py --steps 24 --warmup 3 $B --sort-interval 6 --sort-lookahead 0
run i6_la1 python bench.py --steps 24 --warmup 3 $B --sort-interval 6 --sort-lookahead 1
run i7_la0 python bench.py --steps 21 --warmup 3 $B --sort-interval 7 --sort-lookahead 0
run i7_la2 python bench.py --steps 21 --warmup 3 $B --sort-interval 7 --sort-lookahead 2
run i4_la0 python bench.py --steps 20 --warmup 3 $B --sort-interval 4 --sort-lookahead 0
run i5_la3_keep VPB_SORT_GATHER_KEEP=1 python bench.py --steps 20 --warmup 3 $B --sort-interval 5
run i5_la3_nopack VPB_SORT_PACK_RANK=0 python bench.py --steps 20 --warmup 3 $B --sort-interval 5
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2j_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "avg_launch %.2f" % d["roofline"]["avg_launch_ms"],
              "minmax %.2f %.2f" % (d["roofline"]["min_launch_ms"], d["roofline"]["max_launch_ms"]),
              "sort ms %.1f frac %.3f" % (d["sort_p"]["ms_per_sort"], d["sort_p"]["frac"]), "sort/step %.2f" % d["breakdown_ms_per_step"]["sort_p"])
    except Exception as e:
        print(f, "failed", e)
PY
