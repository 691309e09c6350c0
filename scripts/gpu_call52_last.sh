#!/usr/bin/env bash
# the last ~30 GPU-seconds of round 1: a reference deck on TWO ranks sharing the GPU (autoboot + host-staged transport)
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
export VPB_RUN_UNVALIDATED=1
timeout 18 python -m pytest tests/test_gpu_deck.py -x -q -p no:cacheprovider -k "deck_on_ranks and 2" > $O/pytest_gpu52_ranks.log 2>&1; echo "ranks rc=$?" | tee $O/call52.txt
timeout 9 python -m pytest tests/test_gpu_aniso.py -x -q -p no:cacheprovider -k "fields_aniso and n0" > $O/pytest_gpu52_aniso.log 2>&1; echo "aniso rc=$?" | tee -a $O/call52.txt
tail -n 12 $O/pytest_gpu52_ranks.log; tail -n 3 $O/pytest_gpu52_aniso.log
