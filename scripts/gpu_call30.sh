#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
echo "--- small, e2e only"; timeout 300 python bench.py --cells 64 --ppc 8 --field-cells 0 --no-cpu-baseline --steps 3 --warmup 1 2>&1 | tail -c 600
echo; echo "--- small, cpu baseline only"; timeout 300 python bench.py --cells 64 --ppc 8 --field-cells 0 --no-e2e --steps 3 --warmup 1 2>&1 | tail -c 600
