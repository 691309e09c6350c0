#!/usr/bin/env bash
# Round 2, GPU call L (8 GPUs):   gpurun --gpus 8 --timeout 600 -- 'bash scripts/gpu_r2l.sh'
# BASELINE configs[4] (bench.py --workload harris3d: 1024 x 512 x 512 cells, 2x2x2), its scaled-down parity run against the
# oracle on 2x2x2 ranks, and the headline workload at N=8 with this round's defaults
set -u
mkdir -p gpurun_out
S=gpurun_out/r2l_summary.txt
: > $S
nvidia-smi -L | wc -l | tee -a $S
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
timeout 280 $T --master-port 29531 bench.py --gpus 8 --workload harris3d --steps 15 --warmup 3 $B > gpurun_out/r2l_bench_h3d_n8.json 2> gpurun_out/r2l_bench_h3d_n8.err
echo "bench harris3d N=8 rc=$?" | tee -a $S
VPB_DIST_KIND=harris timeout 150 $T --master-port 29532 tests/dist_gpu_worker.py > gpurun_out/r2l_dist_harris8.log 2>&1
echo "dist harris 2x2x2 rc=$?" | tee -a $S
grep -h "DIST_GPU_OK" gpurun_out/r2l_dist_harris8.log | tee -a $S
timeout 200 $T --master-port 29533 bench.py --gpus 8 --steps 15 --warmup 3 $B > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err
echo "bench thermal N=8 rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2l_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "frac %.3f" % d["roofline"]["frac"],
              "breakdown", {k: round(v, 3) for k, v in d["breakdown_ms_per_step"].items()}, "sort", d["sort_p"]["ms_per_sort"])
    except Exception as e:
        print(f, "failed", e)
PY
for f in gpurun_out/r2l_bench_*.err; do echo "== $f"; tail -n 4 $f; done | tee -a $S
