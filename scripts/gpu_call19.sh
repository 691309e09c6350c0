#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_particles.py -m gpu -q -x -n 4 -p no:cacheprovider -k "pair or planes" > $O/pytest_gpu19_pair.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu19_pair.log
tail -15 $O/pytest_gpu19_pair.log
for c in 4 3; do
  VPB_ADVANCE_P_PAIR_CPS=$c timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b19_256_pair$c.json 2> $O/b19_256_pair$c.err
  tail -c 1200 $O/b19_256_pair$c.json; tail -3 $O/b19_256_pair$c.err
done
