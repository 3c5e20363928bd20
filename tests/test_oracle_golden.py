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
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(b).detach().double()
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


def _criterion_case(golden_dir):
    g = np.load(os.path.join(golden_dir, "criterion_ref.npz"))
    names = [str(n) for n in g["names"]]
    cfg = dict(datasets=["scannet", "s3dis", "arkitscenes"], datasets_weights=[1.0, 0.7, 1.3], topk=[6, 4, 5],
               loss_weight=[0.5, 1.0], non_object_weight=0.1, w_cls=0.5, w_box=2.0, iter_matcher=True)
    gts = [dict(labels=torch.as_tensor(g[f"gt_labels{i}"]), boxes=torch.as_tensor(g[f"gt_boxes{i}"]),
                query_masks=torch.as_tensor(g[f"qmask{i}"])) for i in range(len(names))]
    layers = [dict(cls_preds=[torch.as_tensor(g[f"l{l}_cls{i}"]) for i in range(len(names))],
                   bboxes=[torch.as_tensor(g[f"l{l}_box{i}"]) for i in range(len(names))]) for l in range(3)]
    return g, names, cfg, gts, layers


def test_criterion_matches_reference_criterion(golden_dir):
    """oracle/criterion.py vs the reference's own criterion.py + IoU losses (tests/golden/criterion_ref.npz):
    matched (query, gt) pairs bit-exact, per-layer losses and det_loss within 1e-5."""
    from oracle import criterion as oc
    g, names, cfg, gts, layers = _criterion_case(golden_dir)
    total = 0.0
    for l, lay in enumerate(layers):
        for i in range(len(names)):
            if len(gts[i]["labels"]) == 0:
                continue
            iq, ig = oc.uni_matcher(lay["cls_preds"][i], lay["bboxes"][i], gts[i]["labels"], gts[i]["boxes"],
                                    gts[i]["query_masks"], cfg["topk"][cfg["datasets"].index(names[i])])
            assert np.array_equal(iq.numpy(), g[f"l{l}_iq{i}"]) and np.array_equal(ig.numpy(), g[f"l{l}_ig{i}"]), (l, i)
        loss, _ = oc.layer_loss(lay["cls_preds"], lay["bboxes"], gts, names, cfg)
        assert abs(float(loss) - float(g[f"layer_loss{l}"])) < 1e-5 * max(1.0, abs(float(g[f"layer_loss{l}"])))
        total += float(loss)
    pred = dict(layers[0], aux_outputs=layers[1:])
    det = float(oc.criterion(pred, gts, names, cfg))
    assert abs(det - float(g["det_loss"])) < 1e-5 * abs(float(g["det_loss"])) and abs(det - total) < 1e-4


def test_gt_target_helpers_match_reference(golden_dir):
    """get_bboxes_by_masks / get_targets (unidet3d.py:220-275, 371-409) restatements vs the reference functions."""
    from oracle import criterion as oc
    g = np.load(os.path.join(golden_dir, "criterion_ref.npz"))
    bb = oc.bboxes_by_masks(g["bm_inst"], g["bm_points"])
    assert np.array_equal(bb.numpy(), g["bm_boxes"])
    tg = oc.targets_by_distance(g["tg_centers"], g["tg_boxes"], 6)
    assert np.array_equal(tg.numpy(), g["tg_masks"])


@pytest.mark.parametrize("tag", ["one_gt", "t_eq_k1", "ties", "masked"])
def test_criterion_edge_cases_match_reference(golden_dir, tag):
    """single GT / T == topk + 1 / duplicated predictions (tied costs) / a GT whose queries are all masked out:
    matched pairs bit-exact and layer loss within 1e-5 of the reference's own criterion.py."""
    from oracle import criterion as oc
    g = np.load(os.path.join(golden_dir, "criterion_ref.npz"))
    cfg = dict(datasets=["scannet", "s3dis", "arkitscenes"], datasets_weights=[1.0, 0.7, 1.3], topk=[6, 4, 5],
               loss_weight=[0.5, 1.0], non_object_weight=0.1, w_cls=0.5, w_box=2.0, iter_matcher=True)
    cls, box = torch.as_tensor(g[f"e_{tag}_cls"]), torch.as_tensor(g[f"e_{tag}_box"])
    gt = dict(labels=torch.as_tensor(g[f"e_{tag}_labels"]), boxes=torch.as_tensor(g[f"e_{tag}_gt"]),
              query_masks=torch.as_tensor(g[f"e_{tag}_qm"]))
    iq, ig = oc.uni_matcher(cls, box, gt["labels"], gt["boxes"], gt["query_masks"], 6)
    assert np.array_equal(iq.numpy(), g[f"e_{tag}_iq"]) and np.array_equal(ig.numpy(), g[f"e_{tag}_ig"])
    loss, _ = oc.layer_loss([cls], [box], [gt], ["scannet"], cfg)
    assert abs(float(loss) - float(g[f"e_{tag}_loss"])) < 1e-5 * max(1.0, abs(float(g[f"e_{tag}_loss"])))


def test_encoder_parameter_gradients_match_reference_module_under_autograd(golden_dir):
    """backward_ref.npz: the reference's UniDet3DEncoder in train mode (all three heads), loss = randomly weighted logits
    and boxes, torch.autograd -> the oracle under autograd gives the same gradient for every parameter and input."""
    g = _load(golden_dir, "backward_ref.npz")
    sd = {k[len("enc_sd."):]: torch.as_tensor(g[k]).clone().requires_grad_(torch.as_tensor(g[k]).is_floating_point())
          for k in g.files if k.startswith("enc_sd.")}
    classes = [["chair", "table", "sofa"], ["table", "board"], ["bed", "chair", "oven", "sink"]]
    cfg = dict(num_layers=2, num_heads=2, activation_fn="gelu", datasets=["scannet", "s3dis", "arkitscenes"],
               datasets_classes=classes, angles=[False, False, True])
    x = [torch.as_tensor(g[f"enc_x{i}"]).clone().requires_grad_(True) for i in range(3)]
    c = [torch.as_tensor(g[f"enc_c{i}"]) for i in range(3)]
    out = oenc.encoder_forward(sd, cfg, x, c, [str(n) for n in g["enc_names"]], all_heads=True)
    heads = out["aux_outputs"] + [dict(cls_preds=out["cls_preds"], bboxes=out["bboxes"])]
    loss = 0.0
    for h, hd in enumerate(heads):
        for i in range(3):
            loss = loss + (hd["cls_preds"][i] * torch.as_tensor(g[f"enc_gc{h}_{i}"])).sum() + (hd["bboxes"][i] * torch.as_tensor(g[f"enc_gb{h}_{i}"])).sum()
    loss.backward()
    names = [k[len("enc_grad."):] for k in g.files if k.startswith("enc_grad.")]
    assert len(names) == 36
    for k in names:
        assert sd[k].grad is not None, k
        assert _relerr(sd[k].grad, g["enc_grad." + k]) < 2e-4, (k, _relerr(sd[k].grad, g["enc_grad." + k]))
    for i in range(3):
        assert _relerr(x[i].grad, g[f"enc_dx{i}"]) < 2e-4


def test_unet_parameter_gradients_match_reference_module_under_autograd(golden_dir):
    """backward_ref.npz: the reference's SpConvUNet in TRAIN mode (BatchNorm batch statistics) over the dense conv3d stand-in,
    loss = randomly weighted output features -> the oracle (gather / mm / index_add convs, F.batch_norm(training=True))
    under autograd gives the same output, the same gradient for all 74 parameters and for the input features."""
    from oracle import spconv as ospconv
    g = _load(golden_dir, "backward_ref.npz")
    sd = {}
    for k in g.files:
        if k.startswith("unet_sd."):
            t = torch.as_tensor(g[k]).clone()
            track = t.is_floating_point() and not k.endswith(("running_mean", "running_var"))
            sd[k[len("unet_sd."):]] = t.requires_grad_(True) if track else t
    feats = torch.as_tensor(g["unet_feats"]).clone().requires_grad_(True)
    levels = ounet.build_pyramid(g["unet_coords"], g["unet_shape"], 3)
    ospconv.TRAIN_MODE = True
    try:
        y = ounet.unet_forward(sd, feats, levels)
    finally:
        ospconv.TRAIN_MODE = False
    assert _relerr(y, g["unet_out"]) < 2e-5
    (y * torch.as_tensor(g["unet_R"])).sum().backward()
    names = [k[len("unet_grad."):] for k in g.files if k.startswith("unet_grad.")]
    assert len(names) == 74
    worst = max((_relerr(sd[k].grad, g["unet_grad." + k]), k) for k in names)
    assert worst[0] < 1e-3, worst
    assert _relerr(feats.grad, g["unet_dfeats"]) < 1e-3


def test_training_loss_and_all_parameter_gradients_match_the_reference_loss_end_to_end(golden_dir):
    """train_step_ref.npz: the reference's own ``UniDet3D.loss`` (GT boxes by masks / shifted boxes + distance targets,
    collate, SpConvUNet in train mode, pooling, encoder with all heads, criterion) under torch.autograd.  The oracle
    composition (oracle/train_step.py) gives the same derived GT, the same loss and the same gradient for all 113 parameters."""
    from oracle import train_step as ots
    g = _load(golden_dir, "train_step_ref.npz")
    names = [str(n) for n in g["names"]]
    sd = {}
    for k in g.files:
        if k.startswith("sd."):
            t = torch.as_tensor(g[k]).clone()
            track = t.is_floating_point() and not k.endswith(("running_mean", "running_var"))
            sd[k[3:]] = t.requires_grad_(True) if track else t
    classes = [["chair", "table", "sofa", "bed", "sink"], ["table", "board", "bed", "oven"]]
    cfg = dict(voxel_size=float(g["voxel_size"]), min_spatial_shape=32,
               encoder=dict(num_layers=2, num_heads=2, activation_fn="gelu", datasets=["scannet", "3rscan"], datasets_classes=classes,
                            angles=[False, False]))
    crit_cfg = dict(datasets=["scannet", "3rscan"], datasets_weights=[1.0, 0.7], topk=[3, 2], loss_weight=[0.5, 1.0],
                    non_object_weight=0.1, w_cls=0.5, w_box=2.0, iter_matcher=True)
    points, sps = [g["points0"], g["points1"]], [g["sp0"], g["sp1"]]
    specs = [dict(labels=g["labels0"], inst=g["inst0"], sp_masks=g["sp_masks0"]), dict(labels=g["labels1"], boxes=g["gt_boxes1"])]
    loss, gts = ots.training_loss(sd, cfg, crit_cfg, points, sps, names, specs, target_topk=4)
    for i in range(2):                                    # the GT the reference derived on the way
        assert _relerr(gts[i]["boxes"], g[f"used_boxes{i}"]) < 1e-6
        assert np.array_equal(gts[i]["query_masks"].numpy(), g[f"used_sp_masks{i}"])
    ref = float(g["det_loss"])
    assert abs(float(loss.detach()) - ref) < 1e-4 * abs(ref), (float(loss.detach()), ref)
    loss.backward()
    gnames = [k[len("grad."):] for k in g.files if k.startswith("grad.")]
    assert len(gnames) == 113
    errs = {k: _relerr(sd[k].grad, g["grad." + k]) for k in gnames}
    worst = max(errs.values())
    # measured: loss identical, gradients median 1.2e-6, max 3.7e-6
    assert worst < 1e-4, sorted(((e, k) for k, e in errs.items()), reverse=True)[:5]
    assert sorted(errs.values())[len(errs) // 2] < 1e-5


def test_forward_scenes_matches_the_reference_predict_end_to_end(golden_dir):
    """predict_ref.npz: the reference's own ``UniDet3D.predict`` (collate, backbone, pooling, encoder, predict_by_feat) on
    one scene per dataset flavour -- scannet (fast NMS + superpoint trim), s3dis (aligned 3-D NMS + trim), 3rscan (fast NMS,
    no superpoints, [n, 7] boxes) -- against oracle/detector.py::forward_scenes, the function the GPU detector is compared
    with: same detections (labels exact, scores / boxes to fp32 round-off)."""
    from oracle import detector as odet
    g = _load(golden_dir, "predict_ref.npz")
    names = [str(n) for n in g["names"]]
    sd = {k[3:]: torch.as_tensor(g[k]) for k in g.files if k.startswith("sd.")}
    det_sd = {k: t for k, t in sd.items() if not k.startswith("decoder.")}
    enc_sd = {k[len("decoder."):]: t for k, t in sd.items() if k.startswith("decoder.")}
    classes = [["chair", "table", "sofa", "bed", "sink"], ["table", "board", "bed", "oven"], ["chair", "sofa", "lamp"]]
    cfg = dict(voxel_size=float(g["voxel_size"]), min_spatial_shape=32,
               encoder=dict(num_layers=2, num_heads=2, activation_fn="gelu", datasets=names, datasets_classes=classes,
                            angles=[False, False, False]),
               test_cfg=dict(topk_insts=120, score_thr=0.0, iou_thr=[0.5, 0.55, 0.55], low_sp_thr=0.18, up_sp_thr=0.81),
               fast_nms=[True, False, True], use_superpoints=[True, True, False])
    for i, name in enumerate(names):
        (b, l, s), = odet.forward_scenes(det_sd, enc_sd, cfg, [g[f"points{i}"]], [g[f"sp{i}"]], [name])
        assert b.shape == g[f"boxes{i}"].shape, (name, b.shape, g[f"boxes{i}"].shape)
        assert np.array_equal(np.asarray(l), g[f"labels{i}"]), name
        assert np.allclose(np.asarray(s), g[f"scores{i}"], rtol=1e-4, atol=1e-6), name
        ref = g[f"boxes{i}"]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(np.asarray(b)), fin)
        assert np.allclose(np.asarray(b)[fin], ref[fin], rtol=1e-4, atol=1e-5), (name, np.abs(np.asarray(b)[fin] - ref[fin]).max())
    assert g["boxes2"].shape[1] == 7 and not g["boxes2"][:, 6].any()           # the yaw-0 padding quirk of the fast-NMS path


def test_voxelize_matches_the_reference_collate(golden_dir):
    """collate_ref.npz: the reference's own ``UniDet3D.collate`` (over a numpy MinkowskiEngine stand-in), plain and with
    elastic coordinates of both dtypes, against oracle.voxelize: coordinates, inverse map and spatial shape bit-exact,
    voxel features to fp32 round-off of the mean."""
    g = _load(golden_dir, "collate_ref.npz")
    pts = [g["points0"], g["points1"]]
    for tag, el in (("plain", None), ("elastic", [g["elastic0"], g["elastic1"]])):
        coords, feats, inverse, shape = ovox.voxelize(pts, float(g["voxel_size"]), 16, el)
        assert np.array_equal(coords, g[f"{tag}_coords"]) and coords.dtype == np.int32
        assert np.array_equal(inverse, g[f"{tag}_inverse"])
        assert np.array_equal(shape, g[f"{tag}_shape"])
        assert np.allclose(feats, g[f"{tag}_feats"], rtol=1e-6, atol=1e-6)
    assert not np.array_equal(g["plain_coords"], g["elastic_coords"])
