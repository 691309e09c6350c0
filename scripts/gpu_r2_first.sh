#!/usr/bin/env bash
# First GPU call of round 2: everything that was written in round 1 after the GPU budget was spent.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_r2_first.sh'
set -u
mkdir -p gpurun_out
export VPB_RUN_UNVALIDATED=1
# 1. the whole GPU suite INCLUDING the gated tests (kernel variants, anisotropic cells, driver hooks, multi-rank decks on
#    one GPU through the host program's mp layer, array growth, the trecon-part deck as shipped); no -x: see everything
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_all_pytest.log 2>&1
echo "full gpu pytest (gated included) rc=$?" | tee -a gpurun_out/r2_summary.txt
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2_all_pytest.log | tail -40 | tee -a gpurun_out/r2_summary.txt
unset VPB_RUN_UNVALIDATED
# 2. advance_p_pair variants in bench.py: 0 default, 1 FULL, 2 LEAN, 3 both (average launch ms in the JSON line)
for v in 0 1 2 3; do
  VPB_ADVANCE_P_PAIR_VARIANT=$v python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --field-cells 0 \
    > gpurun_out/r2_bench_variant$v.json 2> gpurun_out/r2_bench_variant$v.err
  echo "bench variant $v rc=$?" | tee -a gpurun_out/r2_summary.txt
done
# the record-scatter variant of the look-ahead sort (breakdown_ms_per_step.sort_p in the JSON line)
VPB_SORT_SCATTER=1 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --field-cells 0 \
  > gpurun_out/r2_bench_sort_scatter.json 2> gpurun_out/r2_bench_sort_scatter.err
echo "bench sort.scatter rc=$?" | tee -a gpurun_out/r2_summary.txt
# the e2e leg with 2-D copies of the 32 hot bytes of every record (e2e.value / h2d_bytes_per_step in the JSON line)
VPB_DROPIN_HOT_ONLY=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --field-cells 0 \
  > gpurun_out/r2_bench_e2e_hot_only.json 2> gpurun_out/r2_bench_e2e_hot_only.err
echo "bench e2e hot_only rc=$?" | tee -a gpurun_out/r2_summary.txt
# 3. BASELINE configs[0] as an unmodified reference host program: on the library, then on the reference alone
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --field-cells 0 --deck-e2e > gpurun_out/r2_bench_deck.json 2> gpurun_out/r2_bench_deck.err
echo "bench deck-e2e rc=$?" | tee -a gpurun_out/r2_summary.txt
python bench.py --impl reference --steps 3 --warmup 1 --deck-e2e > gpurun_out/r2_bench_deck_ref.json 2> gpurun_out/r2_bench_deck_ref.err
echo "bench deck-e2e reference rc=$?" | tee -a gpurun_out/r2_summary.txt
# 4. BASELINE configs[2] throughput (trecon-part shape)
python bench.py --workload harris --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_harris.json 2> gpurun_out/r2_bench_harris.err
echo "bench harris rc=$?" | tee -a gpurun_out/r2_summary.txt
tail -n 5 gpurun_out/r2_all_pytest.log
