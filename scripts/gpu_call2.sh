#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu2.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu2.log
for dep in 1 0; do
  VPB_ADVANCE_P_DEPOSIT=$dep timeout 600 python bench.py --cells 128 --ppc 64 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $O/b2_128_dep$dep.json 2> $O/b2_128_dep$dep.err
done
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/b2_256.json 2> $O/b2_256.err; echo "exit $?" >> $O/b2_256.err
if timeout 300 python bench.py --cells 128 --ppc 64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/plain2.log 2>&1; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:advance_p_kernel -s 2 -c 2 -o $O/prof_advance_p_r1b \
      python bench.py --cells 128 --ppc 64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full2.log 2>&1
fi
if timeout 300 python bench.py --cells 128 --ppc 64 --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/plain3.log 2>&1; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:advance_p_kernel -s 36 -c 1 -o $O/prof_advance_p_r1b_drift \
      python bench.py --cells 128 --ppc 64 --steps 19 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_full3.log 2>&1
fi
ls -la $O | tail -15
