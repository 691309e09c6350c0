"""Host side of the device-resident driver's sort (csrc/vpb_sort_group.cu): the group order key(x,y,z) = fx[x] + fy[y]
+ fz[z] is a bijection from the interior voxels into [0, keys), spatial neighbours are close in key space, and the key
space is at most ~2x the number of voxels.  Pure host code of the library: no GPU needed."""
import numpy as np
import pytest

from old_vpic_b200 import lib


def tables(L, n):
    fx, fy, fz = (np.zeros(m + 2, np.int32) for m in n)
    nkeys = L.vpb_sort_group_order(n[0], n[1], n[2], fx.ctypes.data, fy.ctypes.data, fz.ctypes.data)
    return fx, fy, fz, int(nkeys)


@pytest.mark.parametrize("n", [(1, 1, 1), (5, 4, 3), (16, 16, 16), (33, 1, 17), (64, 64, 64), (100, 3, 50), (2048, 1, 1024), (7, 130, 2)])
def test_group_order_is_a_bijection(n):
    L = lib.load()
    fx, fy, fz, nkeys = tables(L, n)
    key = (fz[1:-1, None, None].astype(np.int64) + fy[None, 1:-1, None] + fx[None, None, 1:-1]).reshape(-1)
    assert key.min() >= 0 and key.max() < nkeys
    assert len(np.unique(key)) == n[0] * n[1] * n[2]
    assert nkeys <= 2 * n[0] * n[1] * n[2] + 4096
    # ghost coordinates share the key of the nearest interior one (the look-ahead key is clamped anyway)
    assert fx[0] == fx[1] and fx[-1] == fx[-2] and fy[0] == fy[1] and fz[-1] == fz[-2]


def test_group_order_keeps_neighbours_close():
    """256^3 (BASELINE configs[3]): for most voxels the six face neighbours lie within 2^16 keys (4 M particles at 64 per
    cell, 200 MB: what the L2 holds), where x-fastest order puts every z neighbour 66 564 voxels away."""
    L = lib.load()
    n = (256, 256, 256)
    fx, fy, fz, nkeys = tables(L, n)
    assert nkeys == 256 ** 3
    rng = np.random.default_rng(0)
    c = rng.integers(2, 256, size=(200000, 3))
    k0 = fx[c[:, 0]].astype(np.int64) + fy[c[:, 1]] + fz[c[:, 2]]
    far = 0
    for a, f in enumerate((fx, fy, fz)):
        d = c.copy()
        d[:, a] += 1
        k1 = fx[d[:, 0]].astype(np.int64) + fy[d[:, 1]] + fz[d[:, 2]]
        far += np.mean(np.abs(k1 - k0) > 2 ** 16)
    assert far / 3 < 0.05, far / 3
