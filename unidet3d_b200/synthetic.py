"""Deterministic synthetic indoor scenes (ScanNet-shaped) for tests and bench.

There is no dataset in the build/bench environment, so every measurement and
parity test runs on scenes from this generator:

* an axis-aligned "room": floor + 4 walls + ``n_boxes`` furniture boxes
  (5 faces each), points sampled uniformly by area on the surfaces with
  sigma = 4 mm normal jitter; total surface area is chosen so the scene hits a
  target number of occupied voxels (9.3 m^2 -> ~30k voxels @ 0.02 m / 100k pts);
* colours uniform in [0, 255] then ``(c - 127.5) / 127.5`` like the reference's
  ``NormalizePointsColor_`` (reference: unidet3d/loading.py:70-106,
  configs/unidet3d_1xb8_scannet.py:181-183);
* superpoints: points bucketed on a coarse grid, ids compacted to ``0..S-1``,
  int64 like the reference's loader (unidet3d/loading.py:23-52).

Only numpy is used so the generator runs identically on the build box and on
the GPU box.
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_scene", "make_batch", "SCENE_PRESETS"]

# name -> (n_points, voxel_size, surface area m^2, superpoint cell m)
SCENE_PRESETS = {
    # BASELINE.json configs[0]: plumbing-sized scene
    "small20k": (20_000, 0.05, 21.0, 0.22),
    # BASELINE.json configs[1..3]: ScanNet-shaped 100k-pt cloud, ~30k voxels
    "scannet100k": (100_000, 0.02, 9.3, 0.075),
    # BASELINE.json configs[4]: S3DIS large-scene stress, ~120k voxels
    "s3dis500k": (500_000, 0.02, 33.5, 0.105),
    # tiny scenes for unit tests
    "tiny": (3_000, 0.05, 2.5, 0.20),
}


def _rect_points(rng, n, origin, u, v):
    """n points uniform on the parallelogram origin + a*u + b*v."""
    a = rng.random((n, 1), dtype=np.float64)
    b = rng.random((n, 1), dtype=np.float64)
    return origin[None, :] + a * u[None, :] + b * v[None, :]


def make_scene(seed: int, n_points: int = 100_000, area: float = 12.0,
               sp_cell: float = 0.21, n_boxes: int = 6, max_superpoints: int = 4096):
    """Return (points float32 [N,6], superpoints int64 [N]).

    points[:, :3] are metres, points[:, 3:] normalised colours in [-1, 1].
    """
    rng = np.random.default_rng(1234 + int(seed))
    # unit-shape room, rescaled below so the total surface area == ``area``
    w, l, h = 1.0 + 0.4 * rng.random(), 1.0 + 0.4 * rng.random(), 0.55 + 0.1 * rng.random()
    rects = []  # (origin, u, v)
    z0 = np.zeros(3)
    rects.append((z0, np.array([w, 0, 0.0]), np.array([0, l, 0.0])))          # floor
    rects.append((z0, np.array([w, 0, 0.0]), np.array([0, 0, h])))            # wall y=0
    rects.append((np.array([0, l, 0.0]), np.array([w, 0, 0.0]), np.array([0, 0, h])))
    rects.append((z0, np.array([0, l, 0.0]), np.array([0, 0, h])))            # wall x=0
    rects.append((np.array([w, 0, 0.0]), np.array([0, l, 0.0]), np.array([0, 0, h])))
    for _ in range(n_boxes):
        sx, sy, sz = 0.1 + 0.25 * rng.random(3)
        sz = min(sz, 0.8 * h)
        ox, oy = rng.random() * (w - sx), rng.random() * (l - sy)
        o = np.array([ox, oy, 0.0])
        ex, ey, ez = np.array([sx, 0, 0.0]), np.array([0, sy, 0.0]), np.array([0, 0, sz])
        rects.append((o + ez, ex, ey))            # top
        rects.append((o, ex, ez))
        rects.append((o + ey, ex, ez))
        rects.append((o, ey, ez))
        rects.append((o + ex, ey, ez))
    areas = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v in rects])
    scale = np.sqrt(area / areas.sum())
    counts = rng.multinomial(n_points, areas / areas.sum())
    pts = []
    for (o, u, v), c in zip(rects, counts):
        if c == 0:
            continue
        p = _rect_points(rng, int(c), o * scale, u * scale, v * scale)
        nrm = np.cross(u, v)
        nrm = nrm / np.linalg.norm(nrm)
        p = p + rng.normal(0.0, 0.004, (int(c), 1)) * nrm[None, :]
        pts.append(p)
    xyz = np.concatenate(pts, 0)
    perm = rng.permutation(len(xyz))
    xyz = xyz[perm]
    # arbitrary world offset so per-scene min subtraction is exercised
    xyz = xyz + rng.uniform(-3.0, 3.0, (1, 3))
    rgb = (rng.integers(0, 256, (len(xyz), 3)).astype(np.float64) - 127.5) / 127.5
    points = np.concatenate([xyz, rgb], 1).astype(np.float32)
    # superpoints: coarse grid buckets, compacted, capped
    cell = sp_cell
    while True:
        g = np.floor((points[:, :3] - points[:, :3].min(0)) / np.float32(cell)).astype(np.int64)
        key = (g[:, 0] * 4096 + g[:, 1]) * 4096 + g[:, 2]
        uniq, sp = np.unique(key, return_inverse=True)
        if len(uniq) <= max_superpoints:
            break
        cell *= 1.15
    return points, sp.astype(np.int64)


def make_batch(preset: str = "scannet100k", batch_size: int = 8, seed0: int = 0):
    """List of (points, superpoints) for ``batch_size`` scenes, plus voxel size."""
    n_points, voxel, area, sp_cell = SCENE_PRESETS[preset]
    scenes = [make_scene(seed0 + i, n_points, area, sp_cell) for i in range(batch_size)]
    return scenes, voxel
