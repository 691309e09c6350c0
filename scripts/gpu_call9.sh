#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > $O/pytest_gpu9.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu9.log
timeout 600 python bench.py --workload fields --field-cells 1024 > $O/b9_fields1024.json 2> $O/b9_fields1024.err
timeout 600 python bench.py --workload fields --field-cells 512 > $O/b9_fields512.json 2> $O/b9_fields512.err
for c in 3 2; do
  VPB_ADVANCE_P_STREAM_CTAS_PER_SM=$c timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b9_256_c$c.json 2> $O/b9_256_c$c.err
done
if timeout 300 python bench.py --steps 19 --warmup 1 --no-e2e --no-cpu-baseline --field-cells 0 > $O/plain9.log 2>&1; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advance_p_stream -s 8 -c 1 -o $O/prof_advance_p_r1i_256_step4 \
      python bench.py --steps 5 --warmup 1 --no-e2e --no-cpu-baseline --field-cells 0 > $O/ncu_full14.log 2>&1
fi
ls $O | tail -3
