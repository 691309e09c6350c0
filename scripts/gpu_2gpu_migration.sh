#!/usr/bin/env bash
# Round 2, GPU call R (2 GPUs):   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_r2r.sh'
# fused migration rounds over NCCL: per-call parity against the oracle cluster (exact / fused / fused with second messages),
# decomposed runs against the one-domain oracle, the reference deck over NCCL; bench at N=2 with the fused rounds on and off
set -u
mkdir -p gpurun_out
S=gpurun_out/r2r_summary.txt
: > $S
nvidia-smi -L | tee -a $S
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_deck.py -q -m gpu -p no:cacheprovider --timeout=600 -rfEs -k "multi or nccl" > gpurun_out/r2r_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2r_pytest.log | tail -30 | tee -a $S
tail -40 gpurun_out/r2r_pytest.log > gpurun_out/r2r_pytest_tail.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 3 $B > gpurun_out/r2r_bench_n2_fused.json 2> gpurun_out/r2r_bench_n2_fused.err
echo "bench N=2 fused rc=$?" | tee -a $S
VPB_BOUNDARY_FUSED=0 timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 3 $B > gpurun_out/r2r_bench_n2_exact.json 2> gpurun_out/r2r_bench_n2_exact.err
echo "bench N=2 exact rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import glob, json
for f in sorted(glob.glob("gpurun_out/r2r_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "frac %.3f" % d["roofline"]["frac"],
              "breakdown", {k: round(v, 3) for k, v in d["breakdown_ms_per_step"].items()}, "sort", d["sort_p"]["ms_per_sort"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2r_bench_*.err | tail -30 | tee -a $S
