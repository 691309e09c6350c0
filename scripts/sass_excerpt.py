#!/usr/bin/env python
"""Static SASS evidence for the kernels bench.py times: opcode histogram and the first occurrences (with their
neighbours) of the instructions DESIGN.md section 4 names.

    python scripts/sass_excerpt.py > profiles/r2_sass_excerpts.txt

Reads the objects old_vpic_b200/build.py leaves under old_vpic_b200/csrc/_obj (cuobjdump -sass, no GPU needed).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "old_vpic_b200", "csrc", "_obj")

# (object, substring of the mangled kernel name, what it is, opcodes to show)
KERNELS = [
    ("vpb_advance_p_pair.o", "advance_p_pair_kernelILi1ELi4ELi1ELi1ELi0EE",
     "advance_p_pair_kernel<WIDE=1, CPS=4, PIPE=1, FULL=1, LEAN=0>: the variant bench.py times (component planes, 96-byte interpolator)",
     ["FFMA2", "FADD2", "FMUL2", "RED.E.ADD.F32", "REDG", "MATCH.ANY", "CREDUX", "LDG.E.EF.64", "LDG.E.64", "LDG.E.ENL2.256", "LDG.E.256", "STG.E.EF.64", "MUFU.RSQ", "MUFU.RCP", "SHFL.BFLY", "ATOMG"]),
    ("vpb_advance_p.o", "advance_p_stream_kernelILi1ELi1ELi0ELi5EE",
     "advance_p_stream_kernel<DEPOSIT=1, WIDE=1, STORE=0, CPS=5>: layer A, 48-byte records through a ring of bulk async copies",
     ["UBLKCP", "SYNCS", "LDS.128", "STG.E.128", "RED.E.ADD.F32", "MATCH.ANY", "LDG.E.256", "LDG.E.ENL2.256"]),
    ("vpb_sort_group.o", "group_keys_kernel", "group_keys_kernel (sort pass 1)", ["MATCH.ANY", "ATOMG", "IMAD.HI", "LDG", "STG"]),
    ("vpb_sort_group.o", "group_invert_kernel", "group_invert_kernel (sort pass 2)", ["LDG", "STG"]),
    ("vpb_sort_group.o", "group_gather_kernel", "group_gather_kernel (sort pass 3)", ["LDG", "STG"]),
]

INSN = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;")


def functions(obj):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True, check=True).stdout
    out, name, cur = {}, None, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = cur
            name, cur = m.group(1), []
            continue
        m = INSN.match(line)
        if m and name:
            cur.append((m.group(1), m.group(2)))
    if name:
        out[name] = cur
    return out


def opcode(text):
    t = text.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0] if t else ""


def main():
    cache = {}
    print("SASS excerpts, sm_100a, nvcc", subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2])
    print("made by scripts/sass_excerpt.py from old_vpic_b200/csrc/_obj/*.o (flags: old_vpic_b200/build.py)\n")
    for obj, key, what, shows in KERNELS:
        if obj not in cache:
            cache[obj] = functions(obj)
        names = [n for n in cache[obj] if key in n]
        if not names:
            print("!! no kernel matching", key, "in", obj)
            continue
        name = names[0]
        ins = cache[obj][name]
        print("=" * 110)
        print(what)
        print(name, "--", len(ins), "instructions")
        hist = collections.Counter(opcode(t) for _, t in ins)
        print("opcode histogram (static):")
        row = []
        for op, n in hist.most_common():
            row.append("%s %d" % (op, n))
        for i in range(0, len(row), 6):
            print("   " + ",  ".join(row[i:i + 6]))
        for s in shows:
            hits = [i for i, (_, t) in enumerate(ins) if opcode(t).startswith(s)]
            if not hits:
                continue
            print("-- %s: %d static occurrences; first three in context" % (s, len(hits)))
            for h in hits[:3]:
                for j in range(max(0, h - 1), min(len(ins), h + 2)):
                    print("   %s /*%s*/  %s ;" % (">>" if j == h else "  ", ins[j][0], ins[j][1]))
        print()


if __name__ == "__main__":
    sys.exit(main())
