#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu11.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu11.log
for s in 0 1; do
  VPB_ADVANCE_P_STREAM_STORE=$s timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b11_256_s$s.json 2> $O/b11_256_s$s.err
done
VPB_ADVANCE_P_STREAM_CTAS_PER_SM=3 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b11_256_s0c3.json 2> $O/b11_256_s0c3.err
if timeout 300 python bench.py --workload fields --field-cells 512 > $O/plain11.log 2>&1; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'advance_b_kernel|advance_e_kernel' -s 6 -c 3 -o $O/prof_fields_r1j_512 \
      python bench.py --workload fields --field-cells 512 > $O/ncu_full15.log 2>&1
fi
ls $O | tail -3
