#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 120 scripts/ubench/f32x2_bench > $O/ubench_f32x2.txt 2>&1; echo "ubench exit $?" >> $O/ubench_f32x2.txt
cat $O/ubench_f32x2.txt
for c in 5 4; do
  VPB_ADVANCE_P_STREAM_CPS=$c timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/b18_256_cps$c.json 2> $O/b18_256_cps$c.err
  tail -c 1500 $O/b18_256_cps$c.json
done
