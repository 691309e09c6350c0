#!/usr/bin/env bash
# Round 2, GPU call F (2 GPUs):   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_r2f.sh'
# field-segment graph (1 and 2 ranks), Harris-sheet decomposed parity, N=2 bench with the migration/halo breakdown, harris3d
set -u
mkdir -p gpurun_out
S=gpurun_out/r2f_summary.txt
: > $S
nvidia-smi -L | tee -a $S
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_multi.py tests/test_gpu_particles.py -q -m gpu -p no:cacheprovider --timeout=600 -rfEs -k "step or multi or sort" > gpurun_out/r2f_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2f_pytest.log | tail -30 | tee -a $S
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 3 $B > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
echo "bench N=2 rc=$?" | tee -a $S
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 $B > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
echo "bench N=1 rc=$?" | tee -a $S
timeout 600 $T bench.py --gpus 2 --workload harris3d --steps 25 --warmup 3 $B > gpurun_out/r2f_bench_h3d_n2.json 2> gpurun_out/r2f_bench_h3d_n2.err
echo "bench harris3d N=2 (sort 25) rc=$?" | tee -a $S
timeout 600 $T bench.py --gpus 2 --workload harris3d --steps 20 --warmup 3 --sort-interval 4 $B > gpurun_out/r2f_bench_h3d_n2_i4.json 2> gpurun_out/r2f_bench_h3d_n2_i4.err
echo "bench harris3d N=2 rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2f_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "frac %.3f" % d["roofline"]["frac"],
              "breakdown", {k: round(v, 3) for k, v in d["breakdown_ms_per_step"].items()}, "sort", d["sort_p"]["ms_per_sort"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2f_bench_*.err | tail -30 | tee -a $S
