"""The drop-in link of INTEGRATION.md done for real (oracle/build_hybrid.sh, needs /root/reference at build time): the
reference's host objects minus the hot-path translation units + libvpic_b200.so must link into deck executables with
no symbol left over -- for the UNMODIFIED decks/trecon-part/turbulence.cxx and for this repository's small deck."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HYB = os.path.join(ROOT, "oracle", "_ref", "hybrid")
SO = os.path.join(ROOT, "old_vpic_b200", "libvpic_b200.so")

HOT = {"advance_p", "move_p", "sort_p", "boundary_p", "accumulate_rhob", "center_p", "uncenter_p", "energy_p", "accumulate_rho_p",
       "accumulate_hydro_p", "load_interpolator", "clear_accumulators", "reduce_accumulators", "unload_accumulator",
       "new_interpolator", "delete_interpolator", "new_accumulators", "delete_accumulators", "new_hydro", "delete_hydro",
       "clear_hydro", "synchronize_hydro", "_standard_v4_field_advance", "util_malloc_aligned", "util_free_aligned"}


def dynsyms(path, undefined):
    out = subprocess.run(["nm", "-D", "--undefined-only" if undefined else "--defined-only", path], capture_output=True, text=True,
                         check=True).stdout
    return {line.split()[-1].split("@")[0] for line in out.splitlines() if line.strip()}


@pytest.mark.parametrize("deck", ["turbulence", "thermal_small", "sheet_small", "absorb_small"])
def test_deck_links_against_the_library(deck):
    exe = os.path.join(HYB, deck + ".b200.op")
    if not os.path.exists(exe):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref/hybrid not built (needs /root/reference at build time)")
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_hybrid.sh")])
    provided = dynsyms(SO, undefined=False)
    wanted = {s for s in dynsyms(exe, undefined=True) if s in provided or s in HOT}
    # every hot-path symbol the reference's host code references comes from the library, and nothing is missing
    assert HOT <= wanted if deck == "turbulence" else HOT & wanted
    assert wanted <= provided, sorted(wanted - provided)
    # the reference's own hot-path objects are NOT in the executable
    defined = dynsyms(exe, undefined=False)
    assert not (HOT - {"util_malloc_aligned", "util_free_aligned"}) & defined
