"""Builds old_vpic_b200/libvpic_b200.so in-tree with nvcc for sm_100a.

    python -m old_vpic_b200.build [--force]

Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false.  -fmad=false
is part of the product, not a debug switch: the kernels reproduce the reference's
scalar arithmetic operation by operation, and the x86-64 reference build has no
fused multiply-add (DESIGN.md "numerics").  Objects are cached per source file.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvpic_b200.so")
OBJ = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=default", "--threads", "2"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(d):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def _compile(src, force, hdr_m):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    s = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(s), hdr_m):
        return obj, False
    cmd = ["nvcc"] + NVCC_FLAGS + ["-c", s, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, True


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    hdr_m = _deps_mtime()
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hdr_m), sources()))
    objs = [o for o, _ in res]
    if force or any(c for _, c in res) or not os.path.exists(OUT):
        cmd = ["nvcc", "-shared", "-o", OUT] + objs + ["-Xlinker", "-Bsymbolic", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("built", OUT)
    elif verbose:
        print("up to date", OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
