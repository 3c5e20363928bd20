"""GPU versions of the two training-pipeline transforms that run directly in front of the hot path,
under the reference's names and constructor arguments:

* ``ElasticTransfrom(gran, mag, voxel_size, p)`` -- reference unidet3d/transforms_3d.py:12-83 (the class name's spelling
  is the reference's).  ``transform`` adds ``elastic_coords`` (what ``UniDet3D.loss`` voxelises instead of the points,
  unidet3d.py:152-161).  Random numbers come from ``numpy.random`` in the reference's order (one ``rand()``, then three
  ``randn`` grids per elastic pass): with the same seed the result equals the reference's to double rounding.  The
  blur of the noise grids and the trilinear displacement of every point run on the GPU.
* ``PointSample_(num_points)`` -- reference transforms_3d.py:233-295: ``np.random.choice`` indices (with replacement,
  like the reference), row gathers and the re-indexing of instance / superpoint ids on the GPU.

Inputs are dicts of CUDA tensors with the reference's keys (``points`` fp32 [N, >=3], ``pts_instance_mask`` /
``pts_semantic_mask`` / ``sp_pts_mask`` int64 [N]).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .ops import _p, _stream, check


def _dims3(d):
    return (C.c_int32 * 3)(int(d[0]), int(d[1]), int(d[2]))


def elastic_blur(noise: torch.Tensor) -> torch.Tensor:
    """noise fp32 [3, X, Y, Z] (CUDA), blurred IN PLACE (six 3-tap passes, transforms_3d.py:60-74)."""
    if not (noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous() and noise.dim() == 4 and noise.shape[0] == 3):
        raise _lib.Ud3dError("elastic_blur: expected a contiguous CUDA fp32 tensor [3, X, Y, Z]")
    lib = _lib.load()
    d = _dims3(noise.shape[1:])
    wsb = int(lib.ud3d_elastic_workspace_bytes(d))
    ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=noise.device)
    check(lib.ud3d_elastic_blur(_p(noise), d, _p(ws), wsb, _stream()), "ud3d_elastic_blur")
    return noise


def elastic_apply(x: torch.Tensor, noise_blurred: torch.Tensor, gran: float, mag: float) -> torch.Tensor:
    """x fp64 [n, 3] voxel-unit coordinates -> x + interp(x) * mag."""
    if not (x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and x.dim() == 2 and x.shape[1] == 3):
        raise _lib.Ud3dError("elastic_apply: expected a contiguous CUDA fp64 tensor [n, 3]")
    out = torch.empty_like(x)
    check(_lib.load().ud3d_elastic_apply(_p(x), x.shape[0], _p(noise_blurred), _dims3(noise_blurred.shape[1:]), float(gran),
                                         float(mag), _p(out), _stream()), "ud3d_elastic_apply")
    return out


def voxel_units(points: torch.Tensor, voxel_size: float) -> torch.Tensor:
    if not (points.is_cuda and points.dtype == torch.float32 and points.stride(1) == 1 and points.shape[1] >= 3):
        raise _lib.Ud3dError("voxel_units: expected CUDA fp32 points [n, >=3] with unit column stride")
    out = torch.empty((points.shape[0], 3), dtype=torch.float64, device=points.device)
    check(_lib.load().ud3d_points_to_voxel_units(_p(points), points.stride(0), points.shape[0], float(voxel_size), _p(out),
                                                 _stream()), "ud3d_points_to_voxel_units")
    return out


def compact_ids(ids: torch.Tensor, max_id: int):
    """int64 ids -> (dense ranks of the present non-negative values, negative ids stay -1; number of distinct ids)."""
    if not (ids.is_cuda and ids.dtype == torch.int64 and ids.is_contiguous()):
        raise _lib.Ud3dError("compact_ids: expected a contiguous CUDA int64 tensor")
    lib = _lib.load()
    out = torch.empty_like(ids)
    n_unique = torch.zeros(1, dtype=torch.int32, device=ids.device)
    wsb = int(lib.ud3d_compact_ids_workspace_bytes(int(max_id)))
    ws = torch.empty(wsb, dtype=torch.uint8, device=ids.device)
    check(lib.ud3d_compact_ids(_p(ids), ids.numel(), int(max_id), _p(out), _p(n_unique), _p(ws), wsb, _stream()), "ud3d_compact_ids")
    return out, n_unique


def elastic_voxel_coords(elastic: torch.Tensor, scene_offsets: torch.Tensor, batch_size: int):
    """elastic fp64 [n, 3] (packed scenes) -> (coords int32 [n, 4] = (b, floor(el - per-scene min)), max_coord int32 [3])."""
    if not (elastic.is_cuda and elastic.dtype == torch.float64 and elastic.is_contiguous() and elastic.shape[1] == 3):
        raise _lib.Ud3dError("elastic_voxel_coords: expected a contiguous CUDA fp64 tensor [n, 3]")
    lib = _lib.load()
    n = elastic.shape[0]
    coords = torch.empty((n, 4), dtype=torch.int32, device=elastic.device)
    maxc = torch.empty(3, dtype=torch.int32, device=elastic.device)
    wsb = int(lib.ud3d_elastic_voxel_coords_workspace_bytes(batch_size))
    ws = torch.empty(wsb, dtype=torch.uint8, device=elastic.device)
    check(lib.ud3d_elastic_voxel_coords(_p(elastic), n, _p(scene_offsets), batch_size, _p(coords), _p(maxc), _p(ws), wsb,
                                        _stream()), "ud3d_elastic_voxel_coords")
    return coords, maxc


class ElasticTransfrom:
    def __init__(self, gran, mag, voxel_size, p=1.0):
        self.gran, self.mag, self.voxel_size, self.p = gran, mag, voxel_size, p

    def elastic(self, x: torch.Tensor, gran, mag) -> torch.Tensor:
        # the grid size needs |x|.max(0) on the host (a 3-value read-back), exactly as the reference computes it
        amax = x.abs().amax(0).cpu().numpy()
        noise_dim = amax.astype(np.int32) // gran + 3
        noise = np.stack([np.random.randn(noise_dim[0], noise_dim[1], noise_dim[2]).astype('float32') for _ in range(3)])
        noise = elastic_blur(torch.from_numpy(noise).to(x.device))
        return elastic_apply(x, noise, gran, mag)

    def transform(self, input_dict):
        coords = voxel_units(input_dict['points'], self.voxel_size)
        if np.random.rand() < self.p:
            coords = self.elastic(coords, self.gran[0], self.mag[0])
            coords = self.elastic(coords, self.gran[1], self.mag[1])
        input_dict['elastic_coords'] = coords
        return input_dict

    __call__ = transform


class PointSample_:
    def __init__(self, num_points):
        self.num_points = num_points

    def _choices(self, n):
        return np.random.choice(range(n), min(self.num_points, n))          # transforms_3d.py:249-251 (with replacement)

    def transform(self, input_dict):
        points = input_dict['points']
        choices = torch.from_numpy(np.asarray(self._choices(points.shape[0]), dtype=np.int64)).to(points.device)
        input_dict['points'] = points.index_select(0, choices)
        inst = input_dict.get('pts_instance_mask', None)
        sem = input_dict.get('pts_semantic_mask', None)
        sp = input_dict.get('sp_pts_mask', None)
        # one read-back for the id ranges (the bitmap of compact_ids is sized by the largest id)
        tops = [t.max() for t in (inst, sp) if t is not None]
        tops = torch.stack(tops).cpu().tolist() if tops else []
        if inst is not None:
            input_dict['pts_instance_mask'] = compact_ids(inst.index_select(0, choices), max(int(tops.pop(0)), 0))[0]
        if sem is not None:
            input_dict['pts_semantic_mask'] = sem.index_select(0, choices)
        if sp is not None:
            input_dict['sp_pts_mask'] = compact_ids(sp.index_select(0, choices), max(int(tops.pop(0)), 0))[0]
        input_dict['choices'] = choices
        return input_dict

    __call__ = transform
