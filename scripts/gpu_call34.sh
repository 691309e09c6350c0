#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
for cfg in "16 8" "4 2" "4 4" "2 1" "2 2" "8 4" "1 1"; do
  set -- $cfg
  export VPB_SORT_BY=$1 VPB_SORT_SWEEP_CTAS_PER_SM=$2
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:'sort_claim|sort_gather' -s 4 -c 4 --csv --log-file $O/l34.csv python bench.py --steps 19 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/ncu34.log 2>&1
  echo "by=$1 ctas=$2"; grep -E "sort_(claim|gather)" $O/l34.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -4
done
