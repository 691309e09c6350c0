#!/usr/bin/env bash
# the last GPU-seconds of round 1: boundary_p growing tight arrays on two ranks; the deck on four ranks sharing the GPU
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out; mkdir -p $O
export VPB_RUN_UNVALIDATED=1
timeout 7 python -m pytest tests/test_gpu_deck.py -x -q -p no:cacheprovider -k "grows_tight" > $O/pytest_gpu53_grow.log 2>&1; echo "grow rc=$?" | tee $O/call53.txt
timeout 7 python -m pytest tests/test_gpu_deck.py -x -q -p no:cacheprovider -k "deck_on_ranks and 4" > $O/pytest_gpu53_ranks4.log 2>&1; echo "ranks4 rc=$?" | tee -a $O/call53.txt
tail -n 6 $O/pytest_gpu53_grow.log; tail -n 3 $O/pytest_gpu53_ranks4.log
