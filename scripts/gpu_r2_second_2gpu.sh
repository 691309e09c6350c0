#!/usr/bin/env bash
# Second GPU call of round 2, on two GPUs:   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_r2_second_2gpu.sh'
# The multi-GPU pieces written in round 1 after the GPU budget was spent: per-call parity of halos and migration over
# NCCL against the oracle cluster, and reference host programs (decks) on two ranks with the NCCL bootstrap done
# through the host program's own mp layer.
set -u
mkdir -p gpurun_out
export VPB_RUN_UNVALIDATED=1
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2_multi_pytest.log 2>&1
echo "multi pytest rc=$?" | tee -a gpurun_out/r2_summary2.txt
timeout 1200 python -m pytest tests/test_gpu_deck.py -q -m gpu -k "ranks or grows" > gpurun_out/r2_deck2_pytest.log 2>&1
echo "deck (2 ranks, NCCL) pytest rc=$?" | tee -a gpurun_out/r2_summary2.txt
tail -n 8 gpurun_out/r2_multi_pytest.log gpurun_out/r2_deck2_pytest.log
