#!/usr/bin/env bash
# 2-GPU pass: multi-rank functional test, then the weak-scaling bench at N=2
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/gpus2.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_gpu_worker.py > $O/dist2_10.log 2>&1; echo "exit $?" >> $O/dist2_10.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --cells 128 --ppc 64 --steps 20 --warmup 3 > $O/b10_n2_128.json 2> $O/b10_n2_128.err; echo "exit $?" >> $O/b10_n2_128.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 3 > $O/b10_n2_256.json 2> $O/b10_n2_256.err; echo "exit $?" >> $O/b10_n2_256.err
tail -5 $O/dist2_10.log
