"""`-m gpu` wrappers for the multi-GPU workers (torchrun, NCCL): skipped on boxes with fewer GPUs than ranks."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


def gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout
    except (OSError, subprocess.TimeoutExpired):
        return 0
    return sum(1 for line in out.splitlines() if line.startswith("GPU "))


def torchrun(world, worker, timeout=900, env=None):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, worker)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("world,migration", [(2, "exact"), (2, "fused")])
def test_decomposed_run_matches_single_domain(world, migration):
    """tests/dist_gpu_worker.py: thermal plasma split over the ranks, C++ driver, against a single-domain run of the same
    particles (ran at 2 and 8 GPUs in round 1 from scripts/gpu_call47_2gpu.sh / gpu_call43_8gpu.sh); `fused`: the
    driver's migration rounds as fixed-capacity messages (boundary.fused = 1)."""
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    r = torchrun(world, "dist_gpu_worker.py", env={"VPB_BOUNDARY_FUSED": "1"} if migration == "fused" else None)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]


@pytest.mark.parametrize("world", [2, 8])
def test_decomposed_harris_sheet_matches_single_domain(world):
    """BASELINE configs[4] scaled down (VPB_DIST_KIND=harris in tests/dist_gpu_worker.py): the trecon-part plasma and
    sheet field with conducting, particle-reflecting z walls, split 1x1x2 (2x2x2 at 8 GPUs) like bench.py --workload
    harris3d, against the CPU oracle stepping the same particles on one domain."""
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    r = torchrun(world, "dist_gpu_worker.py", env={"VPB_DIST_KIND": "harris"})
    assert r.returncode == 0 and "DIST_GPU_OK kind=harris" in r.stdout, (r.stdout + r.stderr)[-3000:]


@pytest.mark.parametrize("world,migration", [(2, "exact"), (4, "exact"), (2, "fused"), (2, "fused_overflow"), (4, "fused")])
def test_decomposed_calls_match_oracle_cluster(world, migration):
    """tests/dist_gpu_percall_worker.py: per-call, bit-level parity of halos and migration over NCCL; `fused*`: the
    second wave of movers through the fused fixed-capacity rounds (and their second message)."""
    if gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = {"exact": {}, "fused": {"VPB_BOUNDARY_FUSED": "2"}, "fused_overflow": {"VPB_BOUNDARY_FUSED": "2", "VPB_BOUNDARY_CAP_MAX": "16"}}[migration]
    r = torchrun(world, "dist_gpu_percall_worker.py", env=env)
    assert r.returncode == 0 and "PERCALL_OK world=%d" % world in r.stdout, (r.stdout + r.stderr)[-3000:]
