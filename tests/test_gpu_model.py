"""GPU parity of the registered modules against (a) fixtures produced by the reference's own Python
and (b) the CPU oracle on synthetic scenes.  Tolerance: max-abs error / max-abs reference <= 1e-3
(north_star), bit-exact for indices."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import detector as odet, encoder as oenc, unet as ounet
from unidet3d_b200.synthetic import make_scene, SCENE_PRESETS

DEV = "cuda"


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def test_unet_module_vs_reference_fixture(golden_dir):
    import unidet3d_b200 as u
    g = np.load(os.path.join(golden_dir, "unet_ref.npz"))
    sd = {k[3:]: torch.as_tensor(g[k]) for k in g.files if k.startswith("sd.")}
    m = u.MODELS.build(dict(type="SpConvUNet", num_planes=[8, 16, 24, 32, 40], return_blocks=True)).eval()
    assert not m.load_state_dict(sd).missing_keys
    m.to(DEV)
    x = u.SparseConvTensor(torch.as_tensor(g["feats"]).to(DEV), torch.as_tensor(g["coords"]).to(DEV), g["shape"].tolist(), 2)
    y, blocks = m(x)
    assert torch.equal(y.indices.cpu(), torch.as_tensor(g["coords"]))       # rows stay in input order
    assert len(blocks) == 5
    assert relerr(y.features, g["out"]) < 1e-3, relerr(y.features, g["out"])


@pytest.mark.parametrize("planes,reps", [([32, 64, 96, 128, 160], 2), ([32, 64], 1), ([64], 2), ([32, 32, 64], 2)])
def test_unet_stage_plan_matches_per_conv_executor(planes, reps):
    """ud3d_unet_forward (one C call for the whole recursion) == the Python recursion, bit for bit, incl. the per-level
    outputs of return_blocks=True."""
    import unidet3d_b200 as u
    torch.manual_seed(3)
    m = u.MODELS.build(dict(type="SpConvUNet", num_planes=planes, block_reps=reps, return_blocks=True)).eval()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.normal_(0, 0.2); mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5); mod.bias.data.normal_(0, 0.2)
    m.to(DEV)
    g = torch.Generator().manual_seed(0)
    coords = torch.unique(torch.cat([torch.zeros(4000, 1, dtype=torch.int32),
                                     torch.randint(0, 40, (4000, 3), generator=g, dtype=torch.int32)], 1), dim=0)
    feats = torch.randn(coords.shape[0], planes[0], generator=g)
    outs = {}
    for plan in (False, True):
        m.use_stage_plan = plan
        x = u.SparseConvTensor(feats.to(DEV), coords.to(DEV), [40, 40, 40], 1)
        y, blocks = m(x)
        outs[plan] = (y.features.clone(), [b.features.clone() for b in blocks])
    assert torch.equal(outs[True][0], outs[False][0])
    assert len(outs[True][1]) == len(outs[False][1]) == len(planes)
    for a, b in zip(outs[True][1], outs[False][1]):
        assert a.shape == b.shape and torch.equal(a, b)


@pytest.mark.parametrize("d,heads,hidden,layers,act", [(256, 8, 1024, 3, "gelu"), (128, 4, 256, 1, "relu")])
def test_encoder_stage_plan_matches_per_op_executor(d, heads, hidden, layers, act):
    """ud3d_encoder_forward (one C call) == the per-op Python executor, bit for bit."""
    import unidet3d_b200 as u
    torch.manual_seed(11)
    classes = [["chair", "table", "sofa"], ["table", "board"], ["bed", "chair", "oven", "sink"]]
    m = u.MODELS.build(dict(type="UniDet3DEncoder", num_layers=layers, datasets_classes=classes, in_channels=32, d_model=d,
                            num_heads=heads, hidden_dim=hidden, dropout=0.0, activation_fn=act,
                            datasets=["scannet", "s3dis", "arkitscenes"], angles=[False, False, True])).eval().to(DEV)
    lens = [700, 1, 333, 1290]
    names = ["scannet", "arkitscenes", "s3dis", "arkitscenes"]
    x = [torch.randn(t, 32, device=DEV) for t in lens]
    c = [torch.randn(t, 3, device=DEV) for t in lens]
    outs = {}
    for plan in (False, True):
        m.use_stage_plan = plan
        o = m(x, c, names)
        outs[plan] = o
        assert o["aux_outputs"] == []
    for k in ("cls_preds", "bboxes"):
        for a, b in zip(outs[True][k], outs[False][k]):
            assert a.shape == b.shape and torch.equal(a, b)


def test_encoder_module_vs_reference_fixture(golden_dir):
    import unidet3d_b200 as u
    g = np.load(os.path.join(golden_dir, "encoder_ref.npz"))
    sd = {k[3:]: torch.as_tensor(g[k]) for k in g.files if k.startswith("sd.")}
    classes = [["chair", "table", "sofa"], ["table", "board"], ["bed", "chair", "oven", "sink"]]
    m = u.MODELS.build(dict(type="UniDet3DEncoder", num_layers=2, datasets_classes=classes, in_channels=8, d_model=64,
                            num_heads=2, hidden_dim=128, dropout=0.0, activation_fn="gelu",
                            datasets=["scannet", "s3dis", "arkitscenes"], angles=[False, False, True])).eval()
    assert not m.load_state_dict(sd).missing_keys
    m.to(DEV)
    m.eval_aux_outputs = True
    x = [torch.as_tensor(g[f"x{i}"]).to(DEV) for i in range(3)]
    c = [torch.as_tensor(g[f"c{i}"]).to(DEV) for i in range(3)]
    out = m(x, c, [str(n) for n in g["names"]])
    assert len(out["aux_outputs"]) == 2
    for i in range(3):
        assert relerr(out["cls_preds"][i], g[f"cls{i}"]) < 1e-3
        assert relerr(out["bboxes"][i], g[f"box{i}"]) < 1e-3
        for l in range(2):
            assert relerr(out["aux_outputs"][l]["cls_preds"][i], g[f"aux{l}_cls{i}"]) < 1e-3
            assert relerr(out["aux_outputs"][l]["bboxes"][i], g[f"aux{l}_box{i}"]) < 1e-3
    m.eval_aux_outputs = False
    out2 = m(x, c, [str(n) for n in g["names"]])
    assert out2["aux_outputs"] == [] and relerr(out2["cls_preds"][0], g["cls0"]) < 1e-3


def _build(cfg, seed=0):
    import unidet3d_b200 as u
    model = u.MODELS.build(cfg).eval()
    det_sd = ounet.make_detector_backbone_state_dict(6, cfg["backbone"]["num_planes"], seed)
    d = cfg["decoder"]
    n_union = len(set(sum(d["datasets_classes"], []))) + 1
    enc_sd = oenc.make_encoder_state_dict(d["num_layers"], d["in_channels"], d["d_model"], d["hidden_dim"], n_union, seed)
    sd = dict(det_sd)
    sd.update({"decoder." + k: v for k, v in enc_sd.items()})
    res = model.load_state_dict(sd, strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    return model.to(DEV), det_sd, enc_sd


@pytest.mark.parametrize("datasets,preset,B", [(("scannet",), "tiny", 2), (("scannet", "s3dis", "arkitscenes"), "small20k", 3)])
def test_detector_end_to_end_vs_oracle(datasets, preset, B):
    from unidet3d_b200 import configs, ops
    cfg = configs.model_cfg(datasets, topk_insts=200 if preset == "tiny" else 1000)
    model, det_sd, enc_sd = _build(cfg)
    n, v, a, c = SCENE_PRESETS[preset]
    cfg["voxel_size"] = model.voxel_size = v
    scenes = [make_scene(i, n, a, c) for i in range(B)]
    pts, sps = [s[0] for s in scenes], [s[1] for s in scenes]
    names = [datasets[i % len(datasets)] for i in range(B)]
    stages = {}
    ref = odet.forward_scenes(det_sd, enc_sd, configs.oracle_cfg(cfg), pts, sps, names, stages)
    # stage-wise parity
    dev_pts = torch.as_tensor(np.concatenate(pts)).to(DEV)
    offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device=DEV)
    x, inverse = model.collate(dev_pts, offs, B)
    assert np.array_equal(x.indices.cpu().numpy(), stages["coords"])
    assert np.array_equal(inverse.cpu().numpy().astype(np.int64), stages["inverse"])
    assert x.spatial_shape == list(stages["shape"])
    for l, lv in enumerate(x.pyramid.levels):
        assert np.array_equal(lv.subm.cpu().numpy(), stages["levels"][l]["subm"])
    n_sps = [int(s.max()) + 1 for s in sps]
    sp_off = np.concatenate([[0], np.cumsum(n_sps)])
    sp_b = torch.as_tensor(np.concatenate([s + o for s, o in zip(sps, sp_off[:-1])])).to(DEV)
    pooled = model.extract_feat(x, sp_b, inverse, sp_off)
    assert relerr(pooled, stages["pooled"]) < 1e-3, relerr(pooled, stages["pooled"])
    # full path
    launches0 = ops.launch_count(reset=True)
    res = model.forward_scenes(pts, sps, names)
    assert ops.launch_count() > 100
    for i in range(B):
        (b, l, s), (rb, rl, rs) = res[i], ref[i]
        assert b.shape[1] == rb.shape[1]
        # detections are a discrete function of the logits; compare the matched prefix robustly
        assert abs(len(s) - len(rs)) <= max(2, len(rs) // 50)
        m = min(len(s), len(rs))
        same = (l[:m] == rl[:m]).float().mean()
        assert same > 0.9, same
        assert np.allclose(np.sort(s.numpy())[::-1][:m // 2], np.sort(rs.numpy())[::-1][:m // 2], rtol=2e-3, atol=1e-5)


def test_encoder_logits_and_boxes_full_size():
    """fp parity of the final logits / boxes at the real model size (d=256, 6 layers, 19-way head)."""
    from unidet3d_b200 import configs
    cfg = configs.model_cfg(("scannet",))
    model, det_sd, enc_sd = _build(cfg, seed=1)
    g = torch.Generator().manual_seed(0)
    T = [700, 333]
    x = [torch.randn(t, 32, generator=g) for t in T]
    c = [torch.randn(t, 3, generator=g) for t in T]
    ref = oenc.encoder_forward(enc_sd, configs.oracle_cfg(cfg)["encoder"], x, c, ["scannet"] * 2, all_heads=False)
    out = model.decoder([t.to(DEV) for t in x], [t.to(DEV) for t in c], ["scannet"] * 2)
    for i in range(2):
        assert relerr(out["cls_preds"][i], ref["cls_preds"][i]) < 1e-3
        assert relerr(out["bboxes"][i], ref["bboxes"][i]) < 1e-3


def test_pinned_batch_stager_matches_direct_call():
    """Double-buffered H2D staging (io.PinnedBatchStager) feeds the detector the same data as a direct call."""
    from unidet3d_b200 import configs, io
    cfg = configs.model_cfg(("scannet",), topk_insts=200)
    model, det_sd, enc_sd = _build(cfg)
    n, v, a, c = SCENE_PRESETS["tiny"]
    cfg["voxel_size"] = model.voxel_size = v
    batches = [([make_scene(10 * b + i, n, a, c)[0] for i in range(2)], [make_scene(10 * b + i, n, a, c)[1] for i in range(2)])
               for b in range(3)]
    got = [model.forward_scenes(p, s, ["scannet"] * 2, ns) for p, s, ns in io.PinnedBatchStager(batches, DEV)]
    for (pts, sps), res in zip(batches, got):
        ref = model.forward_scenes(pts, sps, ["scannet"] * 2)
        for (b, l, s), (rb, rl, rs) in zip(res, ref):
            assert torch.equal(l, rl) and torch.allclose(s, rs, rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_loss_value_vs_oracle():
    """UniDet3D.loss (unidet3d.py:277-364, forward value): GT boxes by instance masks (scannet) / shifted GT boxes with
    distance targets (arkitscenes, rotated), all seven heads, matcher + criterion -- against the CPU oracle end to end."""
    import unidet3d_b200 as u
    from unidet3d_b200 import configs
    from unidet3d_b200.structures import DepthInstance3DBoxes, Det3DDataSample, InstanceData, PointData
    from unidet3d_b200.synthetic import make_scene, make_model_state_dict, SCENE_PRESETS
    from oracle import criterion as oc, encoder as oenc, unet as ounet, voxelize as ovox
    from oracle.pool import scatter_mean, superpoint_pool
    datasets = ("scannet", "arkitscenes")
    cfg = configs.model_cfg(datasets, topk_insts=100)
    n, v, a, c = SCENE_PRESETS["tiny"]
    cfg["voxel_size"] = v
    model = u.MODELS.build(cfg).eval()
    sd = make_model_state_dict(cfg, 0)
    model.load_state_dict(sd, strict=False)
    model.cuda()
    rng = np.random.default_rng(4)
    scenes = [make_scene(20 + i, n, a, c) for i in range(2)]
    names = ["scannet", "arkitscenes"]
    samples, gts_ref = [], []
    for i, (pts, sp) in enumerate(scenes):
        xyz = pts[:, :3] - pts[:, :3].min(0)
        n_sp = int(sp.max()) + 1
        if names[i] == "scannet":
            # instances = unions of superpoints; GT boxes come from the instance masks, sp_masks from the loader
            sp_inst = rng.integers(-1, 6, n_sp)
            sp_inst[:6] = np.arange(6)
            inst = sp_inst[sp]
            labels = rng.integers(0, 18, 6)
            sp_masks = np.stack([sp_inst == k for k in range(6)])
            gi = InstanceData(labels_3d=torch.as_tensor(labels), sp_masks=torch.as_tensor(sp_masks))
            seg = PointData(sp_pts_mask=torch.as_tensor(sp), pts_instance_mask=torch.as_tensor(inst))
            boxes = oc.bboxes_by_masks(inst, xyz)
            qm = torch.as_tensor(sp_masks)
        else:
            G = 5
            ctr = rng.uniform(xyz.min(0), xyz.max(0), (G, 3))
            t = np.concatenate([ctr + pts[:, :3].min(0), rng.uniform(0.3, 1.2, (G, 3)), rng.uniform(-3, 3, (G, 1))], 1).astype(np.float32)
            labels = rng.integers(0, 17, G)
            gi = InstanceData(labels_3d=torch.as_tensor(labels),
                              bboxes_3d=DepthInstance3DBoxes(torch.as_tensor(t), box_dim=7, with_yaw=True, origin=(0.5, 0.5, 0.5)))
            seg = PointData(sp_pts_mask=torch.as_tensor(sp))
            boxes = torch.as_tensor(t.copy())
            boxes[:, :3] -= torch.as_tensor(pts[:, :3].min(0))
            centers = scatter_mean(torch.as_tensor(xyz), torch.as_tensor(sp))
            qm = oc.targets_by_distance(centers, boxes, 6)
        samples.append(Det3DDataSample(lidar_path=f"data/{names[i]}/points/x.bin", gt_pts_seg=seg, gt_instances_3d=gi))
        gts_ref.append(dict(labels=torch.as_tensor(labels), boxes=boxes, query_masks=qm))
    out = model.loss(dict(points=[torch.as_tensor(s[0]) for s in scenes]), samples)
    loss = float(out["det_loss"])
    # oracle pipeline
    det_sd = {k: t for k, t in sd.items() if not k.startswith("decoder.")}
    enc_sd = {k[len("decoder."):]: t for k, t in sd.items() if k.startswith("decoder.")}
    ocfg = configs.oracle_cfg(cfg)
    pts = [s[0] for s in scenes]
    sps = [s[1] for s in scenes]
    n_sps = [int(s.max()) + 1 for s in sps]
    sp_off = np.concatenate([[0], np.cumsum(n_sps)])
    coords, feats, inverse, shape = ovox.voxelize(pts, ocfg["voxel_size"], ocfg["min_spatial_shape"])
    x, _ = ounet.backbone_forward(det_sd, coords, torch.as_tensor(feats), shape)
    pooled = superpoint_pool(x, inverse, np.concatenate([s + o for s, o in zip(sps, sp_off[:-1])]), int(sp_off[-1]))
    xs = [pooled[sp_off[i]:sp_off[i + 1]] for i in range(2)]
    ctrs = [scatter_mean(torch.as_tensor(p[:, :3] - p[:, :3].min(0)), torch.as_tensor(s)) for p, s in zip(pts, sps)]
    pred = oenc.encoder_forward(enc_sd, ocfg["encoder"], xs, ctrs, names, all_heads=True)
    cc = cfg["criterion"]
    ref = float(oc.criterion(pred, gts_ref, names, dict(datasets=list(datasets), datasets_weights=cc["datasets_weights"],
                                                         topk=cc["topk"], loss_weight=cc["loss_weight"],
                                                         non_object_weight=cc["non_object_weight"], iter_matcher=True)))
    assert abs(loss - ref) < 2e-3 * abs(ref), (loss, ref)


@pytest.mark.gpu
def test_pipelined_batches_match_one_at_a_time():
    """forward_pipelined (several batches in flight on their own streams, pinned host inputs) returns, in order, what
    forward_scenes returns for each batch on its own."""
    import unidet3d_b200 as u
    from unidet3d_b200 import configs
    from unidet3d_b200.synthetic import make_scene, make_model_state_dict, SCENE_PRESETS
    cfg = configs.model_cfg(("scannet", "arkitscenes"), topk_insts=100)
    n, v, a, c = SCENE_PRESETS["tiny"]
    cfg["voxel_size"] = v
    model = u.MODELS.build(cfg).eval()
    model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
    model.cuda()
    batches = []
    for j in range(5):
        scenes = [make_scene(100 + 3 * j + i, n + 100 * i, a, c) for i in range(2 + j % 2)]
        names = [("scannet", "arkitscenes")[(i + j) % 2] for i in range(len(scenes))]
        batches.append(([torch.as_tensor(s[0]).pin_memory() for s in scenes], [torch.as_tensor(s[1]).pin_memory() for s in scenes], names))
    # the threaded pipeline runs FIRST, on the cold model: with one host thread per batch the first launch of every kernel
    # (function attributes in the per-device context) happens concurrently from two threads
    cold = list(model.forward_pipelined(iter(batches), depth=2, threaded=True))
    ref = [model.forward_scenes(*b) for b in batches]
    runs = [cold] + [list(model.forward_pipelined(iter(batches), depth=depth, threaded=threaded))
                     for depth in (1, 2, 3) for threaded in (False, True)]
    for got in runs:
        assert len(got) == len(ref)
        for rb, gb in zip(ref, got):
            assert len(rb) == len(gb)
            for (b0, l0, s0), (b1, l1, s1) in zip(rb, gb):
                assert b0.shape == b1.shape and torch.equal(l0, l1)
                assert torch.allclose(s0, s1, atol=1e-5)
                # boxes: the superpoint pooling sums with float atomics (1e-7 run-to-run noise), and a trimmed box is the
                # AABB of a voted point set -- a point on a box face can flip in or out, moving one face by a point spacing
                d = (b0 - b1).abs()
                assert float(d.max()) < 0.05 and float((d <= 1e-4).float().mean()) > 0.9, d.max()
