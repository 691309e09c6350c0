#!/usr/bin/env bash
# Round 2, GPU call T (2 GPUs): per-rank class times of the bench (is boundary_p the wait for the slower rank?), the driver's
# fused rounds (boundary.fused = 1) against the one-domain oracle
set -u
mkdir -p gpurun_out
S=gpurun_out/r2t_summary.txt
: > $S
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
B="--no-e2e --no-cpu-baseline --field-cells 0 --no-deck-e2e"
timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 3 $B > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err
echo "bench N=2 rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
d = json.loads([l for l in open("gpurun_out/r2t_bench_n2.json") if l.startswith("{")][-1])
print("ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "breakdown", {k: round(v, 3) for k, v in d["breakdown_ms_per_step"].items()})
print("ranks_ms_per_step", d["ranks_ms_per_step"])
PY
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider --timeout=300 -rfEs -k "single_domain and 2-fused" > gpurun_out/r2t_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2t_pytest.log | tail -5 | tee -a $S
