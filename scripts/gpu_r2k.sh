#!/usr/bin/env bash
# Round 2, GPU call K (2 GPUs):   gpurun --gpus 2 --timeout 700 -- 'bash scripts/gpu_r2k.sh'
# the multi-GPU workers (thermal and Harris-sheet decomposed runs against the oracle, per-call NCCL parity), each under
# its own short timeout
set -u
mkdir -p gpurun_out
S=gpurun_out/r2k_summary.txt
: > $S
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $T --master-port 29521 tests/dist_gpu_worker.py > gpurun_out/r2k_dist_thermal.log 2>&1
echo "dist thermal rc=$?" | tee -a $S
VPB_DIST_KIND=harris timeout 200 $T --master-port 29522 tests/dist_gpu_worker.py > gpurun_out/r2k_dist_harris.log 2>&1
echo "dist harris rc=$?" | tee -a $S
timeout 200 $T --master-port 29523 tests/dist_gpu_percall_worker.py > gpurun_out/r2k_dist_percall.log 2>&1
echo "dist percall rc=$?" | tee -a $S
grep -h "DIST_GPU_OK\|PERCALL_OK\|Error\|error\|assert" gpurun_out/r2k_dist_*.log | tail -20 | tee -a $S
timeout 200 python -m pytest tests/test_gpu_deck.py -q -m gpu -p no:cacheprovider --timeout=150 -rfEs -k "two_gpus" > gpurun_out/r2k_pytest_deck.log 2>&1
echo "deck on two gpus rc=$?" | tee -a $S
tail -3 gpurun_out/r2k_pytest_deck.log | tee -a $S
