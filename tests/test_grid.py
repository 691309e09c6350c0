"""The host-side grid mirror (old_vpic_b200.grid) builds the same grid_t as the reference's
size_grid/join_grid/set_fbc/set_pbc/partition_* (src/grid/ops.c, partition.c)."""
import numpy as np
import pytest

from helpers import RefGrid, abi, host_grid


@pytest.mark.parametrize("kind", ["periodic", "metal", "absorbing"])
@pytest.mark.parametrize("n", [(6, 5, 4), (8, 1, 6), (1, 1, 16), (3, 3, 3)])
def test_single_rank_grid_matches_reference(ref_scalar, kind, n):
    r = RefGrid(ref_scalar, n, kind, Lbox=(2.0, 3.0, 5.0))
    h = host_grid(n, kind, L=(2.0, 3.0, 5.0), dt=r.struct.dt)
    for k in ("dt", "cvac", "eps0", "damp", "x0", "y0", "z0", "x1", "y1", "z1", "dx", "dy", "dz", "rdx", "rdy", "rdz",
              "nx", "ny", "nz", "rangel", "rangeh"):
        assert getattr(h.struct, k) == getattr(r.struct, k), k
    assert list(h.struct.bc) == list(r.struct.bc)
    assert np.array_equal(h.neighbor, r.neighbor)


def test_two_rank_decomposition_is_consistent():
    """What rank 0 thinks lies across its +x face is rank 1's first interior plane, and vice versa."""
    g0 = host_grid((8, 4, 4), "periodic", topo=(2, 1, 1), rank=0)
    g1 = host_grid((8, 4, 4), "periodic", topo=(2, 1, 1), rank=1)
    assert g0.n == (4, 4, 4) and g1.n == (4, 4, 4)
    assert g0.struct.bc[abi.boundary(1, 0, 0)] == 1 and g0.struct.bc[abi.boundary(-1, 0, 0)] == 1
    assert g1.struct.bc[abi.boundary(1, 0, 0)] == 0 and g0.struct.bc[abi.boundary(0, 1, 0)] == 0
    v = g0.voxel(4, 2, 3)
    nn = g0.neighbor[6 * v + 3]
    assert g1.struct.rangel <= nn <= g1.struct.rangeh
    assert nn - g1.struct.rangel == g1.voxel(1, 2, 3)
    v = g1.voxel(1, 2, 3)
    assert g1.neighbor[6 * v + 0] - g0.struct.rangel == g0.voxel(4, 2, 3)
    # y is periodic onto the same rank
    v = g0.voxel(2, 4, 1)
    assert g0.neighbor[6 * v + 4] - g0.struct.rangel == g0.voxel(2, 1, 1)
    assert abs(g1.struct.x0 - 4.0) < 1e-6 and abs(g0.struct.x1 - 4.0) < 1e-6
