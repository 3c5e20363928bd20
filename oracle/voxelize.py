"""Oracle: voxelisation (reference: unidet3d/unidet3d.py:136-176 ``collate``).

Restates ``ME.utils.batch_sparse_collate`` + ``ME.TensorField(...).sparse()`` +
``field.inverse_mapping`` (MinkowskiEngine fork @ce930ee, Dockerfile:8-11 --
third-party, absent; semantics per SURVEY.md appendix A1):

* coords = floor((xyz - per-scene min) / voxel_size) -> int32, batch index in
  column 0 (unidet3d.py:158-161);  division is IEEE fp32 division;
* feats  = hstack(colour, xyz - per-scene mean)          (unidet3d.py:160);
* unique voxel rows, feature = unweighted mean of member points (ME default
  ``UNWEIGHTED_AVERAGE``), inverse_mapping[p] = voxel row of point p (:174);
* spatial_shape = clip(max(coord)+1 over the batch, min=min_spatial_shape),
  from the PRE-unique coords (:168-169).

Voxel row order is implementation-defined in ME; the canonical order here (and
in the CUDA path) is ascending lexicographic (b, x, y, z).
"""
import numpy as np


def point_coords(points_list, voxel_size, elastic_list=None):
    """Per-point int32 (b,x,y,z) and fp32 6-ch features (before dedup).  ``elastic_list`` (unidet3d.py:162-166): per
    scene [n,3] coordinates already in voxel units, float64 when the augmentation was applied and float32 when its
    coin flip skipped it (transforms_3d.py:39-43); coords = floor(el - el.min(0)) IN THAT DTYPE, features unchanged."""
    coords, feats = [], []
    vs = np.float32(voxel_size)
    for b, p in enumerate(points_list):
        p = np.asarray(p, dtype=np.float32)
        xyz = p[:, :3]
        mn = xyz.min(0)
        if elastic_list is not None:
            el = np.asarray(elastic_list[b])
            assert el.dtype in (np.float32, np.float64)
            c = np.floor(el - el.min(0)).astype(np.int32)
        else:
            c = np.floor((xyz - mn) / vs).astype(np.int32)
        mean = (xyz.astype(np.float64).sum(0) / len(xyz)).astype(np.float32)
        f = np.concatenate([p[:, 3:], xyz - mean], 1).astype(np.float32)
        coords.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
        feats.append(f)
    return np.concatenate(coords, 0), np.concatenate(feats, 0)


def linear_key(coords, dims=None):
    """int64 key, ascending key order == ascending lexicographic (b,x,y,z)."""
    c = coords.astype(np.int64)
    return ((c[:, 0] * (1 << 16) + c[:, 1]) * (1 << 16) + c[:, 2]) * (1 << 16) + c[:, 3]


def voxelize(points_list, voxel_size, min_spatial_shape=128, elastic_list=None):
    """-> coords int32 [M,4], feats fp32 [M,6], inverse int64 [N], spatial_shape int[3]."""
    bcoords, feats = point_coords(points_list, voxel_size, elastic_list)
    spatial_shape = np.clip(bcoords[:, 1:].max(0) + 1, min_spatial_shape, None).astype(np.int64)
    key = linear_key(bcoords)
    uniq, first, inverse, counts = np.unique(key, return_index=True, return_inverse=True,
                                             return_counts=True)
    vox_coords = bcoords[first]
    acc = np.zeros((len(uniq), feats.shape[1]), np.float64)
    np.add.at(acc, inverse, feats.astype(np.float64))
    vox_feats = (acc / counts[:, None]).astype(np.float32)
    return vox_coords, vox_feats, inverse.astype(np.int64), spatial_shape
