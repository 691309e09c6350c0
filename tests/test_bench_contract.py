"""CPU: bench.py's reference arm (the reference's own V4/SSE pthreads advance_p from oracle/_ref on this host's cores)
prints ONE JSON line with the contract's keys, and the library's tuning getter does not pin unset knobs."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT, loader


def test_reference_arm_prints_one_contract_line():
    if not loader.ref_available("sse"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 1e6
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_tuning_getter_does_not_pin_defaults():
    """vpb_get_tuning(name) of a knob nobody set reports 0 without making 0 the knob's value (no device needed)."""
    from old_vpic_b200 import lib
    L = lib.load()
    assert L.vpb_get_tuning(b"test.never_set") == 0
    L.vpb_set_tuning(b"test.never_set", 7)
    assert L.vpb_get_tuning(b"test.never_set") == 7
    env = dict(os.environ, VPB_TEST_FROM_ENV="5")
    code = "from old_vpic_b200 import lib; L = lib.load(); print(L.vpb_get_tuning(b'test.from_env'))"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=120)
    assert r.stdout.strip() == "5", (r.stdout, r.stderr[-500:])
