"""CPU: pin the oracle against fixtures produced by the reference's own Python
(tests/golden/make_golden.py) and against independent known-answer constructions."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import rulebook, unet as ounet, encoder as oenc, postprocess as opost, voxelize as ovox
from oracle.spconv import sparse_conv, weight_to_koc
from oracle.pool import scatter_mean


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _relerr(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def test_encoder_matches_reference_module(golden_dir):
    g = _load(golden_dir, "encoder_ref.npz")
    sd = {k[3:]: torch.as_tensor(g[k]) for k in g.files if k.startswith("sd.")}
    classes = [["chair", "table", "sofa"], ["table", "board"], ["bed", "chair", "oven", "sink"]]
    cfg = dict(num_layers=2, num_heads=2, activation_fn="gelu", datasets=["scannet", "s3dis", "arkitscenes"],
               datasets_classes=classes, angles=[False, False, True])
    x = [torch.as_tensor(g[f"x{i}"]) for i in range(3)]
    c = [torch.as_tensor(g[f"c{i}"]) for i in range(3)]
    out = oenc.encoder_forward(sd, cfg, x, c, [str(n) for n in g["names"]])
    for i in range(3):
        assert out["cls_preds"][i].shape == g[f"cls{i}"].shape
        assert _relerr(out["cls_preds"][i], g[f"cls{i}"]) < 1e-5
        assert _relerr(out["bboxes"][i], g[f"box{i}"]) < 1e-5
        for l in range(2):
            assert _relerr(out["aux_outputs"][l]["cls_preds"][i], g[f"aux{l}_cls{i}"]) < 1e-5
            assert _relerr(out["aux_outputs"][l]["bboxes"][i], g[f"aux{l}_box{i}"]) < 1e-5
    assert out["bboxes"][1].shape[1] == 7 and out["bboxes"][0].shape[1] == 6


def test_unet_matches_reference_module_over_dense_conv(golden_dir):
    g = _load(golden_dir, "unet_ref.npz")
    sd = {k[3:]: torch.as_tensor(g[k]) for k in g.files if k.startswith("sd.")}
    levels = ounet.build_pyramid(g["coords"], g["shape"], 5)
    y = ounet.unet_forward(sd, torch.as_tensor(g["feats"]), levels)
    assert _relerr(y, g["out"]) < 2e-5


def test_sparse_conv_vs_dense_conv3d():
    rng = np.random.default_rng(0)
    shape = np.array([9, 7, 6])
    cc = np.unique(rng.integers(0, shape, (150, 3)), axis=0)
    coords = np.concatenate([np.zeros((len(cc), 1), np.int64), cc], 1).astype(np.int32)
    coords = coords[rng.permutation(len(coords))]
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(len(coords), 5, generator=g)
    # submanifold 3x3x3
    w = torch.randn(4, 3, 3, 3, 5, generator=g)
    table = rulebook.subm3_table(coords, shape)
    assert np.array_equal(table, rulebook.subm3_table_bruteforce(coords, shape))
    out = sparse_conv(feats, table, weight_to_koc(w))
    dense = torch.zeros(1, 5, *shape)
    i = torch.as_tensor(coords).long()
    dense[0, :, i[:, 1], i[:, 2], i[:, 3]] = feats.t()
    ref = F.conv3d(dense, w.permute(0, 4, 1, 2, 3), padding=1)[0, :, i[:, 1], i[:, 2], i[:, 3]].t()
    assert _relerr(out, ref) < 1e-5
    # strided k2 s2 (odd extents 9,7 drop the last index) and its inverse
    w2 = torch.randn(3, 2, 2, 2, 5, generator=g)
    cco, child, up, oshape = rulebook.down2(coords, shape)
    assert list(oshape) == [4, 3, 3]
    d = sparse_conv(feats, child, weight_to_koc(w2))
    refd = F.conv3d(dense, w2.permute(0, 4, 1, 2, 3), stride=2)
    j = torch.as_tensor(cco).long()
    assert _relerr(d, refd[0, :, j[:, 1], j[:, 2], j[:, 3]].t()) < 1e-5
    occ = F.max_pool3d((dense.abs().sum(1, keepdim=True) > 0).float(), 2, 2)
    assert int(occ.sum()) == len(cco)
    w3 = torch.randn(5, 2, 2, 2, 3, generator=g)
    u = sparse_conv(d, up, weight_to_koc(w3))
    dd = torch.zeros(1, 3, *oshape.tolist())
    dd[0, :, j[:, 1], j[:, 2], j[:, 3]] = d.t()
    refu = F.conv_transpose3d(dd, w3.permute(4, 0, 1, 2, 3), stride=2)
    full = torch.zeros(1, 5, *shape.tolist())
    full[:, :, :8, :6, :6] = refu
    assert _relerr(u, full[0, :, i[:, 1], i[:, 2], i[:, 3]].t()) < 1e-5
    assert (u[(up < 0).all(0)] == 0).all()


def test_postprocess_matches_reference_logic(golden_dir):
    g = _load(golden_dir, "post_ref.npz")
    pts, sp = torch.as_tensor(g["points"]), torch.as_tensor(g["sp"])
    for tag, fast, use_sp, thr in [("scannet", True, True, 0.5), ("s3dis", False, True, 0.55),
                                   ("arkit", None, False, 0.55)]:
        b, l, s = opost.predict_by_feat(torch.as_tensor(g[f"{tag}_cls"]), torch.as_tensor(g[f"{tag}_box"]), sp, pts,
                                        topk_insts=300, fast_nms=fast, iou_thr=thr, use_superpoints=use_sp,
                                        low_sp_thr=0.18, up_sp_thr=0.81)
        assert np.array_equal(l.numpy(), g[f"{tag}_out_labels"])
        assert np.allclose(s.numpy(), g[f"{tag}_out_scores"], rtol=1e-6, atol=0)
        ref = g[f"{tag}_out_boxes"]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(b.numpy()), fin)
        assert np.allclose(b.numpy()[fin], ref[fin], rtol=1e-5, atol=1e-6)
    # bbox decode, encoder.py:241-283
    dec = oenc.bbox_pred_to_bbox(torch.as_tensor(g["centers"]), torch.as_tensor(g["arkit_raw"]))
    assert np.allclose(dec.numpy(), g["arkit_box"], rtol=1e-6, atol=1e-6)
    # face distances > 0  <=> inside, unidet3d.py:652-677
    inside = opost.face_distances_inside(pts[:500], torch.as_tensor(g["fd_boxes"]))
    assert np.array_equal(inside.numpy(), (g["fd_out"].min(-1) > 0).T)


def test_voxelize_and_scatter_mean_semantics():
    rng = np.random.default_rng(1)
    pts = [rng.uniform(-2, 2, (500, 6)).astype(np.float32), rng.uniform(0, 1, (300, 6)).astype(np.float32)]
    coords, feats, inv, shape = ovox.voxelize(pts, 0.25, min_spatial_shape=4)
    bc, bf = ovox.point_coords(pts, 0.25)
    assert np.array_equal(coords[inv], bc)
    key = ovox.linear_key(coords)
    assert np.all(np.diff(key) > 0)                       # canonical ascending, unique
    v = 17
    assert np.allclose(feats[v], bf[inv == v].mean(0), atol=1e-6)
    assert np.array_equal(shape, np.maximum(bc[:, 1:].max(0) + 1, 4))
    src = torch.arange(12.0).view(6, 2)
    out = scatter_mean(src, torch.tensor([0, 0, 3, 3, 3, 1]))
    assert out.shape == (4, 2) and torch.equal(out[2], torch.zeros(2)) and torch.equal(out[0], torch.tensor([1.0, 2.0]))
