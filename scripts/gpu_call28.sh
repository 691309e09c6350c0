#!/usr/bin/env bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sort_|particle_convert|scan_' -c 80 --csv --log-file $O/launches_r1q_sort.csv python bench.py --steps 19 --warmup 3 --no-cpu-baseline --no-e2e --field-cells 0 > $O/ncu_launches28.log 2>&1
python - <<'PY'
import csv,re
rows=list(csv.reader(l for l in open('gpurun_out/launches_r1q_sort.csv') if l.startswith('"')))
hdr=rows[0]; H={h:i for i,h in enumerate(hdr)}
cur={}
for r in rows[1:]:
    k=(r[H["ID"]],re.sub(r"\(.*","",r[H["Kernel Name"]])[:44])
    cur.setdefault(k,{})[r[H["Metric Name"]]]=(r[H["Metric Value"]],r[H["Metric Unit"]])
for (i,k),m in cur.items():
    t=m.get("gpu__time_duration.sum",("0",""))
    tv=float(t[0].replace(",","")); tv = tv/1e6 if t[1].startswith("n") else (tv/1e3 if t[1].startswith("u") else tv)
    if tv>0.5: print(i,k,"%.2f ms"%tv, m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum"))
PY
