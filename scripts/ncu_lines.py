#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares from an .ncu-rep captured with --import-source on.
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
per = []
for r in rows[3:]:
    if len(r) > 8 and r[0].isdigit() and r[7].replace('.', '').isdigit():
        per.append((int(r[7]), int(r[0]), r[1][:100], int(r[4]) if r[4].isdigit() else 0))
tot = sum(p[0] for p in per); ts = sum(p[3] for p in per)
print("# %s: %d warp-instructions, %d stall samples" % (rep, tot, ts))
print("# line  %inst  %samples  source")
for n, l, s, sm in sorted(per, key=lambda x: -(x[0] / tot + x[3] / ts))[:top]:
    print("%5d %6.2f %6.2f  %s" % (l, 100 * n / tot, 100 * sm / ts, s))
