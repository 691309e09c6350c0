#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'sort_|planes_to' -c 24 --csv --log-file $O/launches_r1q_sort.csv python bench.py --steps 19 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/ncu_launches32.log 2>&1
python - <<'PY'
import csv,re
rows=list(csv.reader(l for l in open('gpurun_out/launches_r1q_sort.csv') if l.startswith('"')))
hdr=rows[0]; H={h:i for i,h in enumerate(hdr)}
cur={}
for r in rows[1:]:
    k=(int(r[H["ID"]]),re.sub(r"\(.*","",r[H["Kernel Name"]])[:44])
    cur.setdefault(k,{})[r[H["Metric Name"]]]=(r[H["Metric Value"]],r[H["Metric Unit"]])
for (i,k),m in sorted(cur.items()):
    print(i,k," | ".join("%s=%s %s"%(a.split("__")[1][:22],v[0],v[1]) for a,v in m.items()))
PY
