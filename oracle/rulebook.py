"""Oracle: spconv rulebooks (integer, bit-exact) -- numpy.

Restates spconv 2.3.6 indice-pair generation (third-party, absent; semantics
per SURVEY.md appendix A2) at the reference's call sites:

* ``SubMConv3d(k=3, padding=1)`` keys ``subm1..5`` (unidet3d/spconv_unet.py:43-56,
  unidet3d/unidet3d.py:96-103): out set == in set; for kernel offset
  kappa in {0,1,2}^3 (row-major, index k = kx*9+ky*3+kz, centre 13) the input of
  output voxel c is the voxel at c + kappa - 1 when present and inside
  [0, spatial_shape).
* ``SparseConv3d(k=2, stride=2)`` keys ``spconv1..4`` (spconv_unet.py:148-154):
  out_shape = (in_shape - 2)//2 + 1; input c feeds output c//2 through slot
  kappa = c % 2 (slot index sx*4+sy*2+sz) iff c//2 < out_shape; out set = unique
  parents, canonical order ascending (b,x,y,z).
* ``SparseInverseConv3d(k=2)`` (spconv_unet.py:178-183) reuses the same pairs
  reversed; fine rows without a pair produce zero rows.

Tables are stored offset-major, ``int32 [K, N_out]``, -1 = no input.
"""
import numpy as np

from .voxelize import linear_key


def _lookup(sorted_keys, order, q):
    pos = np.searchsorted(sorted_keys, q)
    pos = np.clip(pos, 0, len(sorted_keys) - 1)
    hit = sorted_keys[pos] == q
    return np.where(hit, order[pos], -1)


def subm3_table(coords, spatial_shape):
    """coords int32 [N,4] (any row order) -> int32 [27, N] input-row table."""
    coords = np.asarray(coords, np.int32)
    n = len(coords)
    shape = np.asarray(spatial_shape, np.int64)
    key = linear_key(coords)
    order = np.argsort(key, kind="stable")
    skey = key[order]
    table = np.full((27, n), -1, np.int32)
    c = coords.astype(np.int64)
    for kx in range(3):
        for ky in range(3):
            for kz in range(3):
                k = kx * 9 + ky * 3 + kz
                q = c.copy()
                q[:, 1] += kx - 1
                q[:, 2] += ky - 1
                q[:, 3] += kz - 1
                inb = np.all((q[:, 1:] >= 0) & (q[:, 1:] < shape[None, :]), 1)
                qk = linear_key(np.where(inb[:, None], q, 0))
                r = _lookup(skey, order, qk)
                table[k] = np.where(inb, r, -1)
    return table


def subm3_table_bruteforce(coords, spatial_shape):
    """Independent dict-based construction (known-answer check for small N)."""
    d = {tuple(int(v) for v in row): i for i, row in enumerate(coords)}
    n = len(coords)
    table = np.full((27, n), -1, np.int32)
    for i, (b, x, y, z) in enumerate(coords):
        for kx in range(3):
            for ky in range(3):
                for kz in range(3):
                    q = (int(b), int(x) + kx - 1, int(y) + ky - 1, int(z) + kz - 1)
                    if all(0 <= q[a + 1] < int(spatial_shape[a]) for a in range(3)):
                        table[kx * 9 + ky * 3 + kz, i] = d.get(q, -1)
    return table


def down2(coords, spatial_shape):
    """Strided k=2,s=2 rulebook.

    -> coarse coords int32 [Nc,4] (ascending), child table int32 [8, Nc] (fine row
    per slot, -1 = none), up table int32 [8, Nf] (coarse row of fine row i in
    row slot(i), -1 elsewhere / when dropped), out_shape int64 [3].
    """
    coords = np.asarray(coords, np.int32)
    shape = np.asarray(spatial_shape, np.int64)
    out_shape = (shape - 2) // 2 + 1
    c = coords.astype(np.int64)
    parent = c.copy()
    parent[:, 1:] = c[:, 1:] // 2
    slot = (c[:, 1] % 2) * 4 + (c[:, 2] % 2) * 2 + (c[:, 3] % 2)
    keep = np.all(parent[:, 1:] < out_shape[None, :], 1)
    pkey = linear_key(parent)
    uniq, inv = np.unique(pkey[keep], return_inverse=True)
    nc = len(uniq)
    ccoords = np.zeros((nc, 4), np.int32)
    ccoords[inv] = parent[keep].astype(np.int32)
    child = np.full((8, nc), -1, np.int32)
    rows = np.nonzero(keep)[0]
    child[slot[keep], inv] = rows.astype(np.int32)
    up = np.full((8, len(coords)), -1, np.int32)
    up[slot[keep], rows] = inv.astype(np.int32)
    return ccoords, child, up, out_shape
