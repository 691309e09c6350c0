#!/usr/bin/env python
"""Dynamic opcode mix of the top kernel of an .ncu-rep captured with --import-source on: executed warp instructions
and stall samples per SASS opcode, per unit of work.
usage: ncu_opcode_mix.py report.ncu-rep units_of_work [top_n]      (e.g. 64-particle chunks: np/64)"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
ix = {h: i for i, h in enumerate(hdr)}
inst, smp = collections.Counter(), collections.Counter()
for r in rows[rows.index(hdr) + 1:]:
    if len(r) < len(hdr):
        continue
    op = re.sub(r"^@!?U?P\w+\s+", "", r[ix["Source"]].strip()).split()[0].split(".")[0]
    inst[op] += int(r[ix["Instructions Executed"]])
    smp[op] += int(r[ix["# Samples"]])
ti, ts = sum(inst.values()), sum(smp.values())
print("# %s: %.4g warp instructions = %.1f per unit (%g units), %d stall samples" % (rep.split("/")[-1], ti, ti / units, units, ts))
print("# opcode      per unit   % instr  % samples")
for op, n in inst.most_common(top):
    print("%-10s %9.1f %8.1f %9.1f" % (op, n / units, 100.0 * n / ti, 100.0 * smp[op] / ts))
