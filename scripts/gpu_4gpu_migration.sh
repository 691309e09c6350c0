#!/usr/bin/env bash
# Round 2, GPU call S (4 GPUs):   gpurun --gpus 4 --timeout 600 -- 'bash scripts/gpu_r2s.sh'
# fused migration rounds on a 2x2x1 decomposition (rounds 2 and 3 carry particles): per-call parity over NCCL, bench with the
# fused rounds on and off
set -u
mkdir -p gpurun_out
S=gpurun_out/r2s_summary.txt
: > $S
nvidia-smi -L | tee -a $S
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider --timeout=300 -rfEs -k "oracle_cluster and 4-fused" > gpurun_out/r2s_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2s_pytest.log | tail -30 | tee -a $S
tail -40 gpurun_out/r2s_pytest.log > gpurun_out/r2s_pytest_tail.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
timeout 300 $T bench.py --gpus 4 --steps 10 --warmup 3 $B > gpurun_out/r2s_bench_n4_fused.json 2> gpurun_out/r2s_bench_n4_fused.err
echo "bench N=4 fused rc=$?" | tee -a $S
VPB_BOUNDARY_FUSED=0 timeout 300 $T bench.py --gpus 4 --steps 10 --warmup 3 $B > gpurun_out/r2s_bench_n4_exact.json 2> gpurun_out/r2s_bench_n4_exact.err
echo "bench N=4 exact rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2s_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "frac %.3f" % d["roofline"]["frac"],
              "breakdown", {k: round(v, 3) for k, v in d["breakdown_ms_per_step"].items()}, "sort", d["sort_p"]["ms_per_sort"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2s_bench_*.err | tail -30 | tee -a $S
