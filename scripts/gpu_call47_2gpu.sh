#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/dist_gpu_worker.py > $O/dist2_47.log 2>&1; echo "exit $?" >> $O/dist2_47.log
tail -12 $O/dist2_47.log
