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

__all__ = ["make_scene", "make_batch", "make_scannet_gt", "SCENE_PRESETS", "make_unet_state_dict",
           "make_detector_backbone_state_dict", "make_encoder_state_dict", "make_model_state_dict"]

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


def make_scannet_gt(superpoints: np.ndarray, n_inst: int, seed: int, n_classes: int = 18):
    """ScanNet-style training annotations for a synthetic scene: every instance is a union of superpoints (that is how the
    reference's loader delivers them: ``sp_masks`` [G, n_sp], ``pts_instance_mask`` [N] with -1 = no instance).
    -> (labels int64 [G], sp_masks bool [G, n_sp], pts_instance_mask int64 [N])."""
    rng = np.random.default_rng(seed)
    n_sp = int(superpoints.max()) + 1
    sp_inst = rng.integers(-1, n_inst, n_sp)
    sp_inst[:n_inst] = np.arange(n_inst)              # every instance owns at least one superpoint
    labels = rng.integers(0, n_classes, n_inst).astype(np.int64)
    sp_masks = sp_inst[None, :] == np.arange(n_inst)[:, None]
    return labels, sp_masks, sp_inst[superpoints].astype(np.int64)


# ----------------------------------------------------------------------------------------------
# synthetic, reference-layout weights: there are no checkpoints offline.
# Keys / shapes follow the reference modules (unidet3d/spconv_unet.py, unidet3d/encoder.py,
# unidet3d/unidet3d.py:95-111); BatchNorm statistics are non-trivial so folding bugs show.
# ----------------------------------------------------------------------------------------------
import math  # noqa: E402

import torch  # noqa: E402

def make_unet_state_dict(num_planes, block_reps=2, gen=None, p=""):
    """Random non-trivial weights with the reference's key names / shapes."""
    g = gen or torch.Generator().manual_seed(0)
    sd = {}

    def bn(key, c):
        sd[key + ".weight"] = torch.rand(c, generator=g) + 0.5
        sd[key + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[key + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[key + ".running_var"] = torch.rand(c, generator=g) + 0.5
        sd[key + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def conv(key, co, k, ci, active):
        sd[key] = torch.randn(co, k, k, k, ci, generator=g) * (1.0 / (active * ci)) ** 0.5

    def block(key, ci, co):
        if ci != co:
            conv(key + ".i_branch.0.weight", co, 1, ci, 1)
        bn(key + ".conv_branch.0", ci)
        conv(key + ".conv_branch.2.weight", co, 3, ci, 11)
        bn(key + ".conv_branch.3", co)
        conv(key + ".conv_branch.5.weight", co, 3, co, 11)

    c = num_planes[0]
    for i in range(block_reps):
        block(p + f"blocks.block{i}", c, c)
    if len(num_planes) > 1:
        c1 = num_planes[1]
        bn(p + "conv.0", c)
        conv(p + "conv.2.weight", c1, 2, c, 4)
        sd.update(make_unet_state_dict(num_planes[1:], block_reps, g, p + "u."))
        bn(p + "deconv.0", c1)
        conv(p + "deconv.2.weight", c, 2, c1, 1)
        for i in range(block_reps):
            block(p + f"blocks_tail.block{i}", c * (2 - i), c)
    return sd


def make_detector_backbone_state_dict(in_channels=6, num_planes=(32, 64, 96, 128, 160), seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {"input_conv.0.weight":
          torch.randn(num_planes[0], 3, 3, 3, in_channels, generator=g) * (1.0 / (11 * in_channels)) ** 0.5}
    for k, v in make_unet_state_dict(list(num_planes), 2, g).items():
        sd["unet." + k] = v
    c = num_planes[0]
    sd["output_layer.0.weight"] = torch.rand(c, generator=g) + 0.5
    sd["output_layer.0.bias"] = torch.randn(c, generator=g) * 0.1
    sd["output_layer.0.running_mean"] = torch.randn(c, generator=g) * 0.1
    sd["output_layer.0.running_var"] = torch.rand(c, generator=g) + 0.5
    sd["output_layer.0.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return sd


def make_encoder_state_dict(num_layers, in_channels, d_model, hidden_dim, n_cls_out, seed=0):
    """Random weights with the reference's key names, torch-default-like init scales."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(key, o, i, wscale=None):
        s = wscale if wscale is not None else 1.0 / math.sqrt(i)
        sd[key + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * s
        sd[key + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * s

    def ln(key, c):
        sd[key + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[key + ".bias"] = 0.1 * torch.randn(c, generator=g)

    lin("input_proj.0", d_model, in_channels)
    lin("input_proj.2", d_model, d_model)
    for i in range(num_layers):
        p = f"self_attn_layers.{i}"
        sd[p + ".attn.in_proj_weight"] = (torch.rand(3 * d_model, d_model, generator=g) * 2 - 1) * math.sqrt(6.0 / (4 * d_model))
        sd[p + ".attn.in_proj_bias"] = 0.02 * torch.randn(3 * d_model, generator=g)
        lin(p + ".attn.out_proj", d_model, d_model)
        ln(p + ".norm", d_model)
        lin(f"ffn_layers.{i}.net.0", hidden_dim, d_model)
        lin(f"ffn_layers.{i}.net.3", d_model, hidden_dim)
        ln(f"ffn_layers.{i}.norm", d_model)
    ln("out_norm", d_model)
    lin("outs_cls.0", d_model, d_model)
    lin("outs_cls.2", n_cls_out, d_model)
    lin("out_bboxes.linear", 8, d_model)
    return sd


def make_model_state_dict(cfg, seed=0):
    """Full detector state_dict (``input_conv``, ``unet.*``, ``output_layer``, ``decoder.*``) for a model cfg
    from unidet3d_b200.configs.model_cfg."""
    d = cfg["decoder"]
    sd = make_detector_backbone_state_dict(cfg["in_channels"], cfg["backbone"]["num_planes"], seed)
    n_union = len(set(sum(d["datasets_classes"], []))) + 1
    enc = make_encoder_state_dict(d["num_layers"], d["in_channels"], d["d_model"], d["hidden_dim"], n_union, seed)
    sd.update({"decoder." + k: v for k, v in enc.items()})
    return sd
