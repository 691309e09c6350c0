#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/dist_gpu_worker.py > $O/dist2_25.log 2>&1; echo "exit $?" >> $O/dist2_25.log
tail -5 $O/dist2_25.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > $O/b25_n2_256.json 2> $O/b25_n2_256.err; echo "exit $?" >> $O/b25_n2_256.err
tail -c 600 $O/b25_n2_256.json; tail -3 $O/b25_n2_256.err
