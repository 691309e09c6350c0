"""The drop-in boundary: struct layouts agree with the reference compiled from source, and
libvpic_b200.so loads and exports every entry point include/vpic_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import ROOT, abi, loader
from old_vpic_b200 import lib


def test_layouts_match_reference(ref_scalar):
    out = (C.c_long * 128)()
    n = ref_scalar.refh_layout(out, 128)
    got = list(out[:n])
    P, Mv, I, F, A, Fd, G, S = (abi.particle_dtype, abi.mover_dtype, abi.injector_dtype, abi.interpolator_dtype,
                                abi.accumulator_dtype, abi.field_dtype, abi.GridStruct, abi.SpeciesStruct)
    off = lambda dt, name: dt.fields[name][1]
    want = [
        P.itemsize, off(P, "i"), off(P, "ux"), off(P, "q"), off(P, "tag"), off(P, "tag2"),
        Mv.itemsize, off(Mv, "i"),
        I.itemsize, off(I, "dispx"), off(I, "sp_id"),
        F.itemsize, off(F, "ey"), off(F, "cbx"), off(F, "dcbzdz"),
        A.itemsize, off(A, "jy"), off(A, "jz"),
        Fd.itemsize, off(Fd, "cbx"), off(Fd, "tcax"), off(Fd, "rhob"), off(Fd, "jfx"), off(Fd, "rhof"), off(Fd, "ematx"),
        off(Fd, "fmatx"), off(Fd, "cmat"),
        64,
        C.sizeof(G), G.dt.offset, G.damp.offset, G.x0.offset, G.dx.offset, G.rdx.offset, G.nx.offset, G.bc.offset,
        G.range.offset, G.neighbor.offset, G.rangel.offset, G.rangeh.offset, G.nb.offset, G.boundary.offset,
        C.sizeof(S), S.np.offset, S.max_np.offset, S.p.offset, S.nm.offset, S.max_nm.offset, S.pm.offset, S.q_m.offset,
        S.sort_interval.offset, S.sort_out_of_place.offset, S.partition.offset, S.next.offset, S.name.offset,
        C.sizeof(abi.FieldAdvanceMethods), abi.FieldAdvanceMethods.advance_b.offset, abi.FieldAdvanceMethods.energy_f.offset,
        abi.FieldAdvanceMethods.clean_div_b.offset,
    ]
    assert got[:len(want)] == want
    # field_advance_t {f,m,g,method[1]} and material_t (field_advance.h:307-312, material.h:42-50)
    assert got[len(want):len(want) + 2] == [184, 24]
    assert got[len(want) + 2:] == [72, 4, 56, 64]


def test_c_header_static_asserts_compile(tmp_path):
    """include/*.h carries the same numbers as static asserts; compile it as C and as C++."""
    src = tmp_path / "t.c"
    src.write_text('#include "vpic_b200.h"\nint main(void){return 0;}\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c11", "-I", inc, "-c", str(src), "-o", str(tmp_path / "t.o")])
    subprocess.check_call(["g++", "-std=c++17", "-x", "c++", "-I", inc, "-c", str(src), "-o", str(tmp_path / "t2.o")])


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "vpic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef[^;{]*\(\s*\*[^;]*;", "", text)         # function-pointer typedefs are not prototypes
    text = re.sub(r"\w+\s*\(\s*\*\s*\w+\s*\)\s*\([^)]*\)\s*;", ";", text)   # nor are function-pointer members
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text))
    names -= {"defined", "VPB_STATIC_ASSERT", "sizeof", "offsetof"}
    # function-pointer parameter names etc. do not occur in this header; keep only real prototypes
    return sorted(n for n in names if re.search(r"\b%s\s*\([^;{]*\)\s*;" % re.escape(n), text))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(lib.SO), "run python -m old_vpic_b200.build"
    L = C.CDLL(lib.SO)      # loading must not need a GPU
    missing = [n for n in _declared_functions() if not hasattr(L, n)]
    assert not missing, missing
    for name in lib.DATA_SYMBOLS:
        tab = abi.FieldAdvanceMethods.in_dll(L, name)
        assert all(getattr(tab, n) for n in abi.FieldAdvanceMethods.NAMES), name
    assert set(lib.SIGNATURES) >= set(_declared_functions()), sorted(set(_declared_functions()) - set(lib.SIGNATURES))


def _prototypes(text):
    """name -> (return type, [parameter types]) for every C prototype in `text`, types normalised for comparison."""
    import re
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"^\s*#.*$", " ", text, flags=re.M)
    text = re.sub(r"ALIGNED\(\d+\)|BEGIN_C_DECLS|END_C_DECLS", " ", text)
    out = {}

    def norm(t):
        t = re.sub(r"ALIGNED\(\d+\)|\bconst\b|\brestrict\b|\bstruct\b|\bextern\b", " ", t)
        t = re.sub(r"\bvpb_", "", t)
        t = t.replace("*", " * ")
        return " ".join(t.split())

    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b([a-z_][a-z0-9_]*)\s*\(([^;{}()]*)\)\s*;", text):
        ret, name, params = m.group(1), m.group(2), m.group(3)
        plist = []
        for prm in params.split(","):
            prm = norm(prm)
            if prm in ("void", ""):
                continue
            toks = prm.split()
            if len(toks) > 1 and toks[-1] != "*" and not toks[-1].endswith("_t"):
                toks = toks[:-1]                      # drop the parameter name
            plist.append(" ".join(toks))
        out[name] = (norm(ret), plist)
    return out


def test_reference_named_prototypes_match_the_reference_headers():
    """Same names, same argument lists: every layer-A function of include/vpic_b200.h against the reference header
    that declares it (spa.h, sf_interface.h, field_advance.h, util_base.h), struct names compared modulo the vpb_
    prefix.  (Layouts are covered by test_layouts_match_reference.)"""
    ref_root = "/root/reference/src"
    if not os.path.isdir(ref_root):
        pytest.skip("needs /root/reference")
    ref = {}
    for h in ("species_advance/standard/spa.h", "sf_interface/sf_interface.h", "field_advance/field_advance.h", "util/util_base.h"):
        ref.update(_prototypes(open(os.path.join(ref_root, h)).read()))
    mine_text = open(os.path.join(ROOT, "include", "vpic_b200.h")).read()
    mine_text = mine_text[:mine_text.index("(B) Device-resident layer")]
    mine = {k: v for k, v in _prototypes(mine_text).items() if not k.startswith("vpb_")}
    assert len(mine) >= 27
    checked = 0
    for name, (ret, params) in mine.items():
        assert name in ref, "%s is not declared by the reference headers" % name
        r_ret, r_params = ref[name]
        assert len(params) == len(r_params), (name, params, r_params)
        for a, b in zip(params, r_params):       # an opaque pointer (void *) may stand for any pointer of the reference
            assert a == b or (a == "void *" and b.endswith("*")), (name, params, r_params)
        assert ret == r_ret, (name, ret, r_ret)
        checked += 1
    assert checked >= 27


def test_product_does_not_use_the_oracle():
    """oracle/ is test infrastructure: nothing under old_vpic_b200/ imports, links or executes it (comments may name its
    files), the library does not depend on an oracle shared object, and bench.py reaches for it only inside its
    cpu_baseline / --impl reference legs."""
    import re
    import subprocess
    pkg = os.path.join(ROOT, "old_vpic_b200")
    for base, _, files in os.walk(pkg):
        if "_obj" in base:
            continue
        for name in files:
            if not name.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                continue
            for ln in open(os.path.join(base, name), errors="replace"):
                code = ln.split("//")[0].split("#")[0] if not name.endswith(".py") else ln.split("#")[0]
                assert not re.search(r"(import|from)\s+oracle|libvpic_oracle|orc_[a-z_]+\s*\(|oracle/_ref", code), (name, ln)
    so = os.path.join(pkg, "libvpic_b200.so")
    if os.path.exists(so):
        needed = subprocess.run(["objdump", "-p", so], capture_output=True, text=True).stdout
        assert "oracle" not in needed and "vpic_ref" not in needed
    src = open(os.path.join(ROOT, "bench.py")).read()
    for m in re.finditer(r"^(\s*)from oracle import|^(\s*)import oracle", src, re.M):
        indent = len(m.group(1) or m.group(2) or "")
        assert indent >= 4, "bench.py imports the oracle at module level"
