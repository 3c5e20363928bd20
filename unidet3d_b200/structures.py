"""Plain containers mirroring the few mmengine / mmdet3d / spconv types that cross the hot-path
boundary (the real packages are not required):

* ``SparseConvTensor`` -- spconv.pytorch.SparseConvTensor surface used by the reference
  (unidet3d/unidet3d.py:353-354,456-457; unidet3d/spconv_unet.py:83-85,216-218): ``features``,
  ``indices`` int32 [N,4] (b,x,y,z), ``spatial_shape``, ``batch_size``, ``replace_feature``; the
  ``indice_dict`` rulebook cache becomes ``pyramid`` (see rulebook.py).
* ``DepthInstance3DBoxes`` -- tensor holder with mmdet3d's origin convention
  (unidet3d/unidet3d.py:529-533,591).
* ``InstanceData`` / ``PointData`` / ``Det3DDataSample`` -- attribute bags (unidet3d/structures.py:5-25).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch


class SparseConvTensor:
    def __init__(self, features: torch.Tensor, indices: torch.Tensor, spatial_shape: Sequence[int], batch_size: int,
                 pyramid=None, canonical: bool = False, extents: Optional[Sequence[int]] = None):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in (spatial_shape.tolist() if torch.is_tensor(spatial_shape) else spatial_shape)]
        self.batch_size = int(batch_size)
        self.pyramid = pyramid          # rulebook cache (plays the role of spconv's indice_dict)
        self.canonical = canonical      # rows already in ascending (b,x,y,z) order
        self.extents = None if extents is None else [int(e) for e in extents]   # max coord + 1 (<= spatial_shape)
        self.features_act = None        # optional operand-form copy of `features` for the first consumer conv

    def replace_feature(self, new_features: torch.Tensor) -> "SparseConvTensor":
        t = SparseConvTensor(new_features, self.indices, self.spatial_shape, self.batch_size, self.pyramid,
                             self.canonical, self.extents)
        return t


class DepthInstance3DBoxes:
    """Minimal stand-in: stores boxes bottom-centred like mmdet3d (z -= dz * (origin_z - 0))."""

    def __init__(self, tensor: torch.Tensor, box_dim: int = 7, with_yaw: bool = True, origin=(0.5, 0.5, 0.0)):
        t = tensor.clone().reshape(-1, box_dim).to(torch.float32)
        if tuple(origin) != (0.5, 0.5, 0.0):
            dst = t.new_tensor((0.5, 0.5, 0.0))
            src = t.new_tensor(origin)
            t[:, :3] += t[:, 3:6] * (dst - src)
        self.tensor, self.box_dim, self.with_yaw = t, box_dim, with_yaw

    @property
    def gravity_center(self):
        c = self.tensor[:, :3].clone()
        c[:, 2] += self.tensor[:, 5] * 0.5
        return c

    def __len__(self):
        return self.tensor.shape[0]


class _Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(self.__dict__)})"


class InstanceData(_Bag):
    pass


class PointData(_Bag):
    pass


class Det3DDataSample(_Bag):
    """Needs ``lidar_path`` (dataset is inferred from its path components, unidet3d.py:366-369) and
    ``gt_pts_seg.sp_pts_mask`` (int64 superpoint id per point)."""
    pass
