#!/usr/bin/env bash
# the last ~70 GPU-seconds of round 1: default pair kernel after the chunk_lo change, the FULL/LEAN variants' tails, one deck
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
timeout 25 python -m pytest tests/test_gpu_particles.py -x -q -p no:cacheprovider -k "pair_tails or pair_extreme" > $O/pytest_gpu51_default.log 2>&1; echo "default rc=$?" | tee $O/call51.txt
VPB_RUN_UNVALIDATED=1 timeout 25 python -m pytest tests/test_gpu_particles.py -x -q -p no:cacheprovider -k "variants_extreme_and_tails" > $O/pytest_gpu51_variants.log 2>&1; echo "variants rc=$?" | tee -a $O/call51.txt
timeout 20 python -m pytest tests/test_gpu_deck.py -x -q -p no:cacheprovider -k "runs_on_the_library" > $O/pytest_gpu51_deck.log 2>&1; echo "deck rc=$?" | tee -a $O/call51.txt
tail -n 3 $O/pytest_gpu51_default.log $O/pytest_gpu51_variants.log $O/pytest_gpu51_deck.log
