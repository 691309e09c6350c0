#!/usr/bin/env bash
# Round 2, GPU call B (2 GPUs):   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_r2b.sh'
# the un-gated suite after the layer-A / driver changes, the 2-GPU tests (NCCL per-call parity, decks on two GPUs,
# decomposed run against the oracle), and the deck trace with the new layer-A defaults
set -u
mkdir -p gpurun_out
S=gpurun_out/r2b_summary.txt
: > $S
nvidia-smi -L | tee -a $S
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=600 --durations=15 -rfEs -k "not trecon" > gpurun_out/r2b_pytest.log 2>&1
echo "gpu pytest rc=$?" | tee -a $S
grep -E "passed|failed|FAILED|ERROR|SKIPPED" gpurun_out/r2b_pytest.log | tail -40 | tee -a $S
for steps in 40 200; do
( mkdir -p /tmp/deck$steps && cd /tmp/deck$steps && VPB_TRACE=1 VPB_DECK_STEPS=$steps timeout 300 $GRAFT_REPO_ROOT/oracle/_ref/hybrid/thermal_c1.b200.op -tpp=1 \
    > $GRAFT_REPO_ROOT/gpurun_out/r2b_deck_trace_$steps.out 2> $GRAFT_REPO_ROOT/gpurun_out/r2b_deck_trace_$steps.err )
echo "deck trace $steps steps rc=$?" | tee -a $S
grep -hE "simulation time|vpb trace" gpurun_out/r2b_deck_trace_$steps.err | tee -a $S
done
