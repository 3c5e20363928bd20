"""Training-side kernels of the backbone (SURVEY.md 8a R6 second half, 8f rank 2) against torch (CPU) autograd / the
oracle: train-mode (Sync)BatchNorm statistics + running-stat update, sparse-conv weight gradient, sparse-conv input
gradient (= the forward gather-GEMM over the transposed rulebook), train-mode backbone forward."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import spconv as ospconv, unet as ounet, rulebook as orb
from oracle.spconv import sparse_conv
from unidet3d_b200.synthetic import make_scene, SCENE_PRESETS

DEV = "cuda"


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


@pytest.mark.parametrize("n,c", [(1, 32), (777, 32), (20000, 96), (5001, 320)])
def test_bn_train_statistics(n, c):
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(n + c)
    x = torch.randn(n, c, generator=g) * 2 + torch.randn(c, generator=g)
    bn = torch.nn.BatchNorm1d(c, eps=1e-4, momentum=0.1)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(c, generator=g) + 0.5); bn.bias.copy_(torch.randn(c, generator=g))
        bn.running_mean.copy_(torch.randn(c, generator=g)); bn.running_var.copy_(torch.rand(c, generator=g) + 0.5)
    import copy
    ref_bn = copy.deepcopy(bn).train()
    if n > 1:
        ref = ref_bn(x)
    bn_d = copy.deepcopy(bn).to(DEV)
    sc, sh, mean, invstd = ops.bn_train(x.to(DEV), bn_d)
    y = x.to(DEV) * sc + sh
    if n > 1:
        assert relerr(y, ref) < 1e-5
        assert relerr(bn_d.running_mean, ref_bn.running_mean) < 1e-5
        assert relerr(bn_d.running_var, ref_bn.running_var) < 1e-5
        assert int(bn_d.num_batches_tracked) == 1
    assert relerr(mean, x.double().mean(0)) < 1e-5


@pytest.mark.parametrize("c_in,c_out,K,n", [(32, 32, 27, 3000), (64, 32, 27, 900), (6, 32, 27, 5000), (96, 64, 8, 2000),
                                            (256, 100, 1, 1500)])
def test_conv_weight_and_input_gradients_vs_autograd(c_in, c_out, K, n):
    """dW and dX of  y = sum_k x[table[k]] @ W_k  against torch autograd on the oracle's gather -> mm -> index_add_."""
    from unidet3d_b200 import ops
    rng = np.random.default_rng(n + c_in)
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, c_in, generator=g, requires_grad=True)
    w = (torch.randn(c_out, K, c_in, generator=g) / (c_in * 3) ** 0.5).requires_grad_(True)
    if K > 1:
        table = np.where(rng.random((K, n)) < 0.4, rng.integers(0, n, (K, n)), -1).astype(np.int32)
        # a proper rulebook has at most one output per (offset, input): make every offset's map injective
        for k in range(K):
            idx = table[k]
            seen = {}
            for o in np.nonzero(idx >= 0)[0]:
                if idx[o] in seen:
                    idx[o] = -1
                else:
                    seen[idx[o]] = o
        y = sparse_conv(x, table, w.permute(1, 2, 0))
    else:
        table = None
        y = x @ w[:, 0].t()
    dy = torch.randn(n, c_out, generator=g)
    y.backward(dy)
    tb = torch.as_tensor(table).to(DEV) if table is not None else None
    dw = ops.conv_wgrad(x.detach().to(DEV), dy.to(DEV), K, tb)
    assert relerr(dw, w.grad) < 1e-4, relerr(dw, w.grad)
    # accumulate flag
    dw2 = ops.conv_wgrad(x.detach().to(DEV), dy.to(DEV), K, tb, out=dw.clone(), accumulate=True)
    assert relerr(dw2, 2 * w.grad) < 1e-4
    # input gradient through the transposed rulebook: table_t[k][i] = o  <=>  table[k][o] = i
    if table is not None:
        tt = np.full((K, n), -1, np.int32)
        for k in range(K):
            o = np.nonzero(table[k] >= 0)[0]
            tt[k, table[k, o]] = o
        dx = ops.conv_dgrad(dy.to(DEV), w.detach().to(DEV), torch.as_tensor(tt).to(DEV), n, reverse_offsets=False)
    else:
        dx = ops.conv_dgrad(dy.to(DEV), w.detach().to(DEV), None, n, reverse_offsets=False)
    assert relerr(dx, x.grad) < 1e-4, relerr(dx, x.grad)


def test_subm3_input_gradient_uses_the_same_table_reversed():
    """For a submanifold conv the transposed rulebook is the SAME table with the 27 kernel offsets reversed
    (table[k][o] = i  <=>  table[26 - k][i] = o): dX = conv(dY, W^T with offsets flipped) on the level's own table."""
    from unidet3d_b200 import ops
    rng = np.random.default_rng(0)
    coords = np.unique(rng.integers(0, 12, (1500, 3)), axis=0).astype(np.int32)
    coords = np.concatenate([np.zeros((len(coords), 1), np.int32), coords], 1)
    table = orb.subm3_table(coords, np.array([12, 12, 12]))
    n = len(coords)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, 32, generator=g, requires_grad=True)
    w = (torch.randn(32, 27, 32, generator=g) * 0.1).requires_grad_(True)
    y = sparse_conv(x, table, w.permute(1, 2, 0))
    dy = torch.randn(n, 32, generator=g)
    y.backward(dy)
    dx = ops.conv_dgrad(dy.to(DEV), w.detach().to(DEV), torch.as_tensor(table).to(DEV), n, reverse_offsets=True)
    assert relerr(dx, x.grad) < 1e-4, relerr(dx, x.grad)


def test_backbone_train_mode_forward_vs_oracle():
    """module.train(): batch-statistics BatchNorm through the whole U-Net and the output layer, running statistics
    updated like torch -- against the oracle with F.batch_norm(training=True)."""
    import copy
    import unidet3d_b200 as u
    from unidet3d_b200 import configs
    from unidet3d_b200.synthetic import make_model_state_dict
    cfg = configs.model_cfg(("scannet",), topk_insts=100)
    n, v, a, c = SCENE_PRESETS["tiny"]
    cfg["voxel_size"] = v
    model = u.MODELS.build(cfg)
    sd = make_model_state_dict(cfg, 0)
    model.load_state_dict(sd, strict=False)
    model.to(DEV).train()
    scenes = [make_scene(60 + i, n, a, c) for i in range(2)]
    pts, sps = [s[0] for s in scenes], [s[1] for s in scenes]
    P = torch.as_tensor(np.concatenate(pts)).to(DEV)
    offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device=DEV)
    with torch.no_grad():
        x, inv = model.collate(P, offs, 2)
        n_sps = [int(s.max()) + 1 for s in sps]
        sp_off = np.concatenate([[0], np.cumsum(n_sps)])
        sp_b = torch.as_tensor(np.concatenate([s + o for s, o in zip(sps, sp_off[:-1])])).to(DEV)
        pooled = model.extract_feat(x, sp_b, inv, sp_off)
    # oracle, train mode (its BatchNorm updates the running statistics of det_sd in place)
    from oracle import voxelize as ovox
    from oracle.pool import superpoint_pool
    det_sd = {k: t.clone() for k, t in sd.items() if not k.startswith("decoder.")}
    coords, feats, inverse, shape = ovox.voxelize(pts, v, 128)
    ospconv.TRAIN_MODE = True
    try:
        xo, _ = ounet.backbone_forward(det_sd, coords, torch.as_tensor(feats), shape)
    finally:
        ospconv.TRAIN_MODE = False
    ref = superpoint_pool(xo, inverse, np.concatenate([s + o for s, o in zip(sps, sp_off[:-1])]), int(sp_off[-1]))
    assert relerr(pooled, ref) < 1e-3, relerr(pooled, ref)
    msd = model.state_dict()
    for k in ("unet.blocks.block0.conv_branch.0.running_mean", "unet.u.u.blocks.block1.conv_branch.3.running_var",
              "unet.blocks_tail.block0.conv_branch.0.running_var", "output_layer.0.running_mean"):
        assert relerr(msd[k], det_sd[k]) < 1e-3, (k, relerr(msd[k], det_sd[k]))
    # eval mode is unaffected by the mode switch itself (but sees the updated running statistics)
    model.eval()


def test_bn_relu_and_pool_backward_vs_autograd():
    """ud3d_bn_backward_* and ud3d_segmented_mean_backward against torch.autograd."""
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(4)
    n, c = 3001, 48
    x = (torch.randn(n, c, generator=g) * 1.5 + 0.3).requires_grad_(True)
    gamma = (torch.rand(c, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(c, generator=g) * 0.3).requires_grad_(True)
    a = torch.relu(torch.nn.functional.batch_norm(x, None, None, gamma, beta, True, 0.1, 1e-4))
    da = torch.randn(n, c, generator=g)
    a.backward(da)
    bn = torch.nn.BatchNorm1d(c, eps=1e-4, momentum=0.1).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(gamma), bn.bias.copy_(beta)
    xd = x.detach().to(DEV)
    s, h, mean, invstd = ops.bn_train(xd, bn)
    assert relerr(ops.bn_relu_apply(xd, s, h), a) < 1e-5
    dx, dgamma, dbeta = ops.bn_relu_backward(xd, da.to(DEV), s, h, mean, invstd)
    assert relerr(dx, x.grad) < 1e-4, relerr(dx, x.grad)
    assert relerr(dgamma, gamma.grad) < 1e-4 and relerr(dbeta, beta.grad) < 1e-4
    # accumulate into an existing gradient, strided views
    buf = torch.ones(n, 2 * c, device=DEV)
    ops.bn_relu_backward(xd, da.to(DEV), s, h, mean, invstd, dx=buf[:, c:], accumulate=True)
    assert relerr(buf[:, c:] - 1.0, x.grad) < 1e-4 and float((buf[:, :c] - 1.0).abs().max()) == 0.0
    # superpoint mean-pool backward
    from oracle.pool import superpoint_pool
    n_pts, n_sp = 20000, 300
    inv = torch.randint(0, n, (n_pts,), generator=g)
    sp = torch.randint(0, n_sp - 2, (n_pts,), generator=g)
    v = torch.randn(n, c, generator=g, requires_grad=True)
    pooled = superpoint_pool(v, inv, sp, n_sp)
    dp = torch.randn(n_sp, c, generator=g)
    pooled.backward(dp)
    dv = ops.segmented_mean_backward(dp.to(DEV), sp.to(DEV), n, gather=inv.int().to(DEV))
    assert relerr(dv, v.grad) < 1e-5, relerr(dv, v.grad)
    assert torch.equal(dv, ops.segmented_mean_backward(dp.to(DEV), sp.to(DEV), n, gather=inv.int().to(DEV)))     # deterministic


@pytest.mark.parametrize("smooth", [True, False])
def test_backbone_backward_vs_autograd(smooth):
    """Gradients of EVERY backbone parameter (input conv, 48 U-Net convs, 44 + 1 train-mode BatchNorms) from the gradient of
    the pooled superpoint features: the tape of unidet3d_b200/train.py on the library's kernels against torch.autograd
    through the oracle (train-mode BatchNorm = batch statistics).

    ``smooth``: every BatchNorm bias is raised so that no ReLU is ever inactive -- the network is then a smooth function
    and the comparison is tight (1e-3): it checks the whole composition (tape order, residual / concat accumulation,
    transposed rulebooks, BatchNorm backward through the batch statistics).  With the synthetic weights as they are a
    gradient is a heavily cancelling sum over thousands of voxels and a handful of ReLU masks that flip on last-bit
    forward differences (measured: 3e-5 of the activations) move it by percents -- the oracle's OWN gradients move by
    0.1 % .. 12 % (median 3 %) when its input features are perturbed by 3e-5 -- so that variant only bounds the error (the masks themselves are checked
    exactly by test_bn_relu_and_pool_backward_vs_autograd)."""
    import unidet3d_b200 as u
    from unidet3d_b200 import configs, train
    from unidet3d_b200.synthetic import make_model_state_dict
    from oracle import voxelize as ovox
    from oracle.pool import superpoint_pool
    cfg = configs.model_cfg(("scannet",), topk_insts=100)
    n, v, a, c = SCENE_PRESETS["tiny"]
    cfg["voxel_size"] = v
    model = u.MODELS.build(cfg)
    sd = make_model_state_dict(cfg, 0)
    if smooth:
        for k in sd:
            if (".conv_branch.0.bias" in k or ".conv_branch.3.bias" in k or k.endswith(("conv.0.bias", "deconv.0.bias", "output_layer.0.bias"))):
                sd[k] = sd[k] + 9.0
    model.load_state_dict(sd, strict=False)
    model.to(DEV).train()
    scenes = [make_scene(70 + i, n, a, c) for i in range(2)]
    pts, sps = [s[0] for s in scenes], [s[1] for s in scenes]
    P = torch.as_tensor(np.concatenate(pts)).to(DEV)
    offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device=DEV)
    n_sps = [int(s.max()) + 1 for s in sps]
    sp_off = np.concatenate([[0], np.cumsum(n_sps)])
    sp_all = np.concatenate([s + o for s, o in zip(sps, sp_off[:-1])])
    g = torch.Generator().manual_seed(9)
    d_pooled = torch.randn(int(sp_off[-1]), 32, generator=g)
    with torch.no_grad():
        x, inv = model.collate(P, offs, 2)
        pooled, tape = train.backbone_forward(model, x, torch.as_tensor(sp_all).to(DEV), inv, int(sp_off[-1]))
        train.backbone_backward(tape, pooled, d_pooled.to(DEV))
    # oracle with autograd
    det_sd = {k: t.clone().float() for k, t in sd.items() if not k.startswith("decoder.")}
    params = {k: t.requires_grad_(True) for k, t in det_sd.items()
              if t.is_floating_point() and not k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
    coords, feats, inverse, shape = ovox.voxelize(pts, v, 128)
    ospconv.TRAIN_MODE = True
    try:
        xo, _ = ounet.backbone_forward(det_sd, coords, torch.as_tensor(feats), shape)
    finally:
        ospconv.TRAIN_MODE = False
    ref = superpoint_pool(xo, inverse, sp_all, int(sp_off[-1]))
    assert relerr(pooled, ref) < 1e-3
    (ref * d_pooled).sum().backward()
    got = {k: p.grad for k, p in model.named_parameters() if not k.startswith("decoder.")}
    errs = {}
    for k, t in params.items():
        assert k in got and got[k] is not None, k
        errs[k] = relerr(got[k], t.grad)
    assert len(errs) == 49 + 2 * 45        # 48 U-Net convs + input conv, 44 + 1 BatchNorms (gamma, beta)
    tol = 1e-3 if smooth else 0.2
    bad = sorted(((e, k) for k, e in errs.items() if not e < tol), reverse=True)
    assert not bad, (len(bad), bad[:12], sorted((e, k) for k, e in errs.items())[:3])
    if smooth:
        assert float((xo.detach() > 0).float().mean()) == 1.0          # the premise: no inactive ReLU at the output layer


def test_encoder_backward_pieces_vs_autograd():
    """LayerNorm, GELU / ReLU and attention-core backward kernels against torch.autograd."""
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(12)
    rows, c = 1333, 256
    x = torch.randn(rows, c, generator=g, requires_grad=True)
    gamma = (torch.rand(c, generator=g) + 0.5).requires_grad_(True)
    beta = torch.randn(c, generator=g).requires_grad_(True)
    y = torch.nn.functional.layer_norm(x, (c,), gamma, beta, 1e-5)
    dy = torch.randn(rows, c, generator=g)
    y.backward(dy)
    dx, dgamma, dbeta = ops.layernorm_backward(x.detach().to(DEV), dy.to(DEV), gamma.detach().to(DEV), 1e-5)
    assert relerr(dx, x.grad) < 1e-4 and relerr(dgamma, gamma.grad) < 1e-4 and relerr(dbeta, beta.grad) < 1e-4
    again = ops.layernorm_backward(x.detach().to(DEV), dy.to(DEV), gamma.detach().to(DEV), 1e-5)
    assert torch.equal(again[1], dgamma) and torch.equal(again[0], dx)          # deterministic
    for act, fn in (("gelu", torch.nn.functional.gelu), ("relu", torch.relu)):
        z = (torch.randn(rows, c, generator=g) * 2).requires_grad_(True)
        fn(z).backward(dy)
        assert relerr(ops.activation_backward(z.detach().to(DEV), dy.to(DEV), act), z.grad) < 1e-5
    # attention core: 3 scenes of different lengths (one shorter than a warp), 8 heads x 32
    lens = [700, 17, 333]
    H, d = 8, 256
    T = sum(lens)
    qkv = (torch.randn(T, 3 * d, generator=g) * 0.7).requires_grad_(True)
    outs, a0 = [], 0
    for n_ in lens:
        q, k, v = [qkv[a0:a0 + n_, i * d:(i + 1) * d].reshape(n_, H, 32).transpose(0, 1) for i in range(3)]
        o = torch.softmax(q @ k.transpose(1, 2) / 32 ** 0.5, -1) @ v
        outs.append(o.transpose(0, 1).reshape(n_, d))
        a0 += n_
    out = torch.cat(outs)
    d_out = torch.randn(T, d, generator=g)
    out.backward(d_out)
    cu = torch.tensor(np.cumsum([0] + lens), dtype=torch.int32, device=DEV)
    for variant in ("reg", "warp"):             # thread-per-token with register-resident rows (product) / warp-per-token
        dqkv = ops.attention_backward(qkv.detach().to(DEV), cu, H, out.detach().to(DEV), d_out.to(DEV), variant=variant)
        assert relerr(dqkv[:, :d], qkv.grad[:, :d]) < 1e-4, (variant, relerr(dqkv[:, :d], qkv.grad[:, :d]))
        assert relerr(dqkv[:, d:2 * d], qkv.grad[:, d:2 * d]) < 1e-4 and relerr(dqkv[:, 2 * d:], qkv.grad[:, 2 * d:]) < 1e-4, variant
        again = ops.attention_backward(qkv.detach().to(DEV), cu, H, out.detach().to(DEV), d_out.to(DEV), variant=variant)
        assert torch.equal(again, dqkv)                                                   # deterministic


def test_encoder_backward_vs_autograd():
    """Gradients of every encoder parameter and of the pooled input features, from random gradients on the class logits
    and boxes of all seven heads (train mode evaluates them all, encoder.py:219-229): the tape of unidet3d_b200/train.py
    against torch.autograd through the oracle encoder (pinned to the reference's own encoder.py)."""
    import unidet3d_b200 as u
    from unidet3d_b200 import configs, train
    from oracle import encoder as oenc
    cfg = configs.model_cfg(("scannet", "s3dis", "arkitscenes"), topk_insts=100)
    d = cfg["decoder"]
    n_union = len(set(sum(d["datasets_classes"], []))) + 1
    enc_sd = oenc.make_encoder_state_dict(d["num_layers"], d["in_channels"], d["d_model"], d["hidden_dim"], n_union, 3)
    enc = u.MODELS.build(d)
    assert not enc.load_state_dict(enc_sd, strict=False).missing_keys
    enc.to(DEV).train()
    g = torch.Generator().manual_seed(21)
    lens = [410, 37, 655]
    names = ["scannet", "arkitscenes", "s3dis"]
    xs = [torch.randn(t, d["in_channels"], generator=g) for t in lens]
    cs = [torch.randn(t, 3, generator=g) for t in lens]
    bounds = [0] + list(np.cumsum(lens))
    X = torch.cat(xs).to(DEV)
    Cn = torch.cat(cs).to(DEV)
    with torch.no_grad():
        out, tape = train.encoder_forward(enc, X, Cn, [int(b) for b in bounds], names)
    # oracle
    sd = {k: t.clone().float().requires_grad_(True) for k, t in enc_sd.items()}
    xo = [x.clone().requires_grad_(True) for x in xs]
    ref = oenc.encoder_forward(sd, configs.oracle_cfg(cfg)["encoder"], xo, cs, names, all_heads=True)
    heads_ref = ref["aux_outputs"] + [dict(cls_preds=ref["cls_preds"], bboxes=ref["bboxes"])]
    heads_out = out["aux_outputs"] + [dict(cls_preds=out["cls_preds"], bboxes=out["bboxes"])]
    assert len(heads_ref) == len(heads_out) == d["num_layers"] + 1
    loss = 0.0
    d_cls, d_box = [], []
    for hr, ho in zip(heads_ref, heads_out):
        dc, db = [], []
        for i in range(len(lens)):
            assert relerr(ho["cls_preds"][i], hr["cls_preds"][i]) < 1e-3 and relerr(ho["bboxes"][i], hr["bboxes"][i]) < 1e-3
            gc = torch.randn(hr["cls_preds"][i].shape, generator=g)
            gb = torch.randn(hr["bboxes"][i].shape, generator=g) * 0.1
            loss = loss + (hr["cls_preds"][i] * gc).sum() + (hr["bboxes"][i] * gb).sum()
            dc.append(gc.to(DEV)), db.append(gb.to(DEV))
        d_cls.append(dc), d_box.append(db)
    loss.backward()
    train.encoder_backward(tape, out, d_cls, d_box)
    errs = {"dX": relerr(tape.grad(X), torch.cat([x.grad for x in xo]))}
    for k, p in enc.named_parameters():
        assert p.grad is not None, k
        errs[k] = relerr(p.grad, sd[k].grad)
    # measured: 3e-6 .. 2.3e-3 on 82 of the 85 tensors.  The tensors right behind a ReLU (input_proj.0 -> dX, the class
    # MLP's outs_cls.0, evaluated by seven heads) see a few masks flip on last-bit forward differences, like the backbone's
    # (test_backbone_backward_vs_autograd): percent-level on those sums, bounded here by 2e-2; 5e-3 everywhere else
    def tol(k):
        return 2e-2 if (k == "dX" or k.startswith(("input_proj.0", "outs_cls.0"))) else 5e-3
    bad = sorted(((e, k) for k, e in errs.items() if not e < tol(k)), reverse=True)
    assert not bad, (len(bad), bad[:10], sorted((e, k) for k, e in errs.items())[:3])
    assert len(errs) == 1 + len(enc_sd)


def _scannet_training_batch(n_scenes, seed0, preset="tiny", n_inst=6):
    """Synthetic ScanNet-style training samples: instances are unions of superpoints (GT boxes come from the instance
    masks, unidet3d.py:220-275), sp_masks from the loader.  -> (scenes, samples, oracle GT dicts)."""
    from unidet3d_b200.structures import Det3DDataSample, InstanceData, PointData
    from oracle import criterion as oc
    n, v, a, c = SCENE_PRESETS[preset]
    rng = np.random.default_rng(seed0)
    scenes = [make_scene(seed0 + i, n, a, c) for i in range(n_scenes)]
    samples, gts_ref = [], []
    for pts, sp in scenes:
        xyz = pts[:, :3] - pts[:, :3].min(0)
        n_sp = int(sp.max()) + 1
        sp_inst = rng.integers(-1, n_inst, n_sp)
        sp_inst[:n_inst] = np.arange(n_inst)
        inst = sp_inst[sp]
        labels = rng.integers(0, 18, n_inst)
        sp_masks = np.stack([sp_inst == k for k in range(n_inst)])
        gi = InstanceData(labels_3d=torch.as_tensor(labels), sp_masks=torch.as_tensor(sp_masks))
        seg = PointData(sp_pts_mask=torch.as_tensor(sp), pts_instance_mask=torch.as_tensor(inst))
        samples.append(Det3DDataSample(lidar_path="data/scannet/points/x.bin", gt_pts_seg=seg, gt_instances_3d=gi))
        gts_ref.append(dict(labels=torch.as_tensor(labels), boxes=oc.bboxes_by_masks(inst, xyz), query_masks=torch.as_tensor(sp_masks)))
    return scenes, samples, gts_ref


def test_training_step_loss_and_all_gradients_vs_oracle_autograd():
    """The whole training step of configs[3] on two small scenes: ``train.loss_backward`` (collate -> train-mode backbone
    -> pooling -> encoder with seven heads -> GPU matcher -> criterion -> backward through the criterion, the encoder
    and the backbone) against torch.autograd through the oracle pipeline (pinned to the reference modules): the loss
    value, the matched pairs of every head (vs the oracle's own matcher) and the gradient of EVERY parameter of the
    detector.  The oracle's loss uses OUR matches (a flipped near-tie in the discrete matching would make the gradients
    incomparable); the matches themselves are compared separately.  BatchNorm biases are raised like in the smooth variant
    of test_backbone_backward_vs_autograd so that the backbone's ReLU masks cannot flip."""
    import unidet3d_b200 as u
    from unidet3d_b200 import configs, train
    from unidet3d_b200.synthetic import make_model_state_dict
    from oracle import criterion as oc, encoder as oenc, voxelize as ovox
    from oracle.pool import scatter_mean, superpoint_pool
    cfg = configs.model_cfg(("scannet",), topk_insts=100)
    n, v, a, c = SCENE_PRESETS["tiny"]
    cfg["voxel_size"] = v
    model = u.MODELS.build(cfg)
    sd = make_model_state_dict(cfg, 0)
    for k in sd:
        if (".conv_branch.0.bias" in k or ".conv_branch.3.bias" in k or k.endswith(("conv.0.bias", "deconv.0.bias", "output_layer.0.bias"))):
            sd[k] = sd[k] + 9.0
    model.load_state_dict(sd, strict=False)
    model.to(DEV).train()
    scenes, samples, gts_ref = _scannet_training_batch(2, 90)
    names = ["scannet", "scannet"]
    dbg = {}
    out = train.loss_backward(model, dict(points=[torch.as_tensor(s[0]) for s in scenes]), samples, debug=dbg)
    loss = float(out["det_loss"])
    # ---- oracle pipeline under autograd
    full = {k: t.clone().float() for k, t in sd.items()}
    params = {k: t.requires_grad_(True) for k, t in full.items()
              if t.is_floating_point() and not k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
    det_sd = {k: t for k, t in full.items() if not k.startswith("decoder.")}
    enc_sd = {k[len("decoder."):]: t for k, t in full.items() if k.startswith("decoder.")}
    ocfg = configs.oracle_cfg(cfg)
    pts, sps = [s[0] for s in scenes], [s[1] for s in scenes]
    n_sps = [int(s.max()) + 1 for s in sps]
    sp_off = np.concatenate([[0], np.cumsum(n_sps)])
    coords, feats, inverse, shape = ovox.voxelize(pts, ocfg["voxel_size"], ocfg["min_spatial_shape"])
    ospconv.TRAIN_MODE = True
    try:
        xo, _ = ounet.backbone_forward(det_sd, coords, torch.as_tensor(feats), shape)
    finally:
        ospconv.TRAIN_MODE = False
    pooled = superpoint_pool(xo, inverse, np.concatenate([s + o for s, o in zip(sps, sp_off[:-1])]), int(sp_off[-1]))
    assert relerr(dbg["pooled"], pooled) < 1e-3
    xs = [pooled[sp_off[i]:sp_off[i + 1]] for i in range(2)]
    ctrs = [scatter_mean(torch.as_tensor(p[:, :3] - p[:, :3].min(0)), torch.as_tensor(s)) for p, s in zip(pts, sps)]
    pred = oenc.encoder_forward(enc_sd, ocfg["encoder"], xs, ctrs, names, all_heads=True)
    cc = cfg["criterion"]
    ccfg = dict(datasets=["scannet"], datasets_weights=cc["datasets_weights"], topk=cc["topk"], loss_weight=cc["loss_weight"],
                non_object_weight=cc["non_object_weight"], iter_matcher=True)
    heads = pred["aux_outputs"] + [dict(cls_preds=pred["cls_preds"], bboxes=pred["bboxes"])]
    assert len(heads) == len(dbg["matches"]) == 7
    ref, same, total = 0.0, 0, 0
    for hd, ms in zip(heads, dbg["matches"]):
        idx = [tuple(t.cpu() for t in m.nonzero(as_tuple=True)) for m in ms]
        l, _ = oc.layer_loss(hd["cls_preds"], hd["bboxes"], gts_ref, names, ccfg, indices=idx)
        ref = ref + l
        for i, m in enumerate(ms):                     # the oracle's own matcher on its own predictions
            iq, ig = oc.uni_matcher(hd["cls_preds"][i].detach(), hd["bboxes"][i].detach(), gts_ref[i]["labels"], gts_ref[i]["boxes"],
                                    gts_ref[i]["query_masks"], cc["topk"][0])
            mo = torch.zeros(m.shape, dtype=torch.bool)
            mo[iq, ig] = True
            same += int((mo & m.cpu()).sum())
            total += int(max(mo.sum(), m.sum()))
    assert total > 0 and same >= 0.95 * total, (same, total)
    assert abs(loss - float(ref)) < 2e-3 * abs(float(ref)), (loss, float(ref))
    ref.backward()
    errs = {}
    for k, p in model.named_parameters():
        assert k in params, k
        assert p.grad is not None and params[k].grad is not None, k
        errs[k] = relerr(p.grad, params[k].grad)
    assert len(errs) == len(params)
    vals = sorted(errs.values())
    print("gradient rel. errors: median %.2e, p90 %.2e, max %.2e (%s)" % (vals[len(vals) // 2], vals[int(len(vals) * 0.9)], vals[-1],
                                                                          max(errs, key=errs.get)))
    bad = sorted(((e, k) for k, e in errs.items() if not e < 0.1), reverse=True)
    assert not bad, (len(bad), bad[:12])
    assert vals[len(vals) // 2] < 2e-2, vals[len(vals) // 2]


def test_train_step_updates_parameters_and_reduces_the_loss():
    """``train.train_step`` (zero_grad -> loss_backward -> clip -> AdamW, configs/unidet3d_1xb8_scannet.py optim_wrapper):
    a few steps on one fixed batch lower its loss; eval-mode inference afterwards uses the updated weights (the packed
    weight images are rebuilt)."""
    import unidet3d_b200 as u
    from unidet3d_b200 import configs, train
    from unidet3d_b200.synthetic import make_model_state_dict
    cfg = configs.model_cfg(("scannet",), topk_insts=100)
    n, v, a, c = SCENE_PRESETS["tiny"]
    cfg["voxel_size"] = v
    model = u.MODELS.build(cfg)
    model.load_state_dict(make_model_state_dict(cfg, 0), strict=False)
    model.to(DEV).train()
    scenes, samples, _ = _scannet_training_batch(2, 95)
    inputs = dict(points=[torch.as_tensor(s[0]) for s in scenes])
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.05)
    w0 = model.decoder.out_bboxes.linear.weight.detach().clone()
    c0 = model.unet.blocks.block0.conv_branch[2].weight.detach().clone()
    losses = [float(train.train_step(model, opt, inputs, samples)["det_loss"]) for _ in range(8)]
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0], losses
    assert not torch.equal(w0, model.decoder.out_bboxes.linear.weight) and not torch.equal(c0, model.unet.blocks.block0.conv_branch[2].weight)
    model.eval()
    res = model.forward_scenes([torch.as_tensor(s[0]).to(DEV) for s in scenes], [torch.as_tensor(s[1]).to(DEV) for s in scenes],
                               ["scannet"] * 2)
    assert len(res) == 2 and all(torch.isfinite(r[2]).all() for r in res)


def _criterion_case(datasets, names, Ts, Gs, dims, seed, masked_scene=None):
    """Random heads + GT for the criterion tests -> (crit, outputs on CPU, insts on DEV, oracle gts, oracle cfg)."""
    import types
    from unidet3d_b200.criterion import UniDet3DCriterion
    g = torch.Generator().manual_seed(seed)
    topk = [3, 2, 2][:len(datasets)]
    wts = [1.0, 0.7, 1.3][:len(datasets)]
    crit = UniDet3DCriterion(matcher=dict(costs=[dict(type="QueryClassificationCost", weight=0.5), dict(type="BboxCostJointTraining", weight=2.0)]),
                             loss_weight=[0.5, 1.0], non_object_weight=0.1, iter_matcher=True, bbox_loss_simple=dict(mode="diou"),
                             bbox_loss_rotated=dict(mode="diou"), datasets=datasets, datasets_weights=wts, topk=topk)
    Cs = [6 if n == "scannet" else 4 for n in names]
    gts, insts = [], []
    for i, (T, G, C, dim) in enumerate(zip(Ts, Gs, Cs, dims)):
        labels = torch.randint(0, C, (G,), generator=g)
        boxes = torch.cat((torch.rand(G, 3, generator=g) * 4, torch.rand(G, 3, generator=g) + 0.3), 1)
        if dim == 7:
            boxes = torch.cat((boxes, torch.rand(G, 1, generator=g) * 6 - 3), 1)
        qm = torch.rand(G, T, generator=g) < 0.7
        if masked_scene == i:
            qm[:] = False
        gts.append(dict(labels=labels, boxes=boxes, query_masks=qm))
        insts.append(types.SimpleNamespace(labels_3d=labels.to(DEV), query_masks=qm.to(DEV),
                                           bboxes_3d=types.SimpleNamespace(gravity_center=boxes[:, :3].to(DEV), tensor=boxes.to(DEV), with_yaw=dim == 7)))

    def head():
        cps = [torch.randn(T, C + 1, generator=g) for T, C in zip(Ts, Cs)]
        bbs = []
        for T, dim in zip(Ts, dims):
            b = torch.cat((torch.rand(T, 3, generator=g) * 4, torch.rand(T, 3, generator=g) + 0.2), 1)
            bbs.append(torch.cat((b, torch.rand(T, 1, generator=g) * 6 - 3), 1) if dim == 7 else b)
        return dict(cls_preds=cps, bboxes=bbs)

    final, aux = head(), [head(), head()]
    out = dict(cls_preds=final["cls_preds"], bboxes=final["bboxes"], aux_outputs=aux)
    cfg = dict(datasets=datasets, datasets_weights=wts, topk=topk, loss_weight=[0.5, 1.0], non_object_weight=0.1, w_cls=0.5, w_box=2.0,
               iter_matcher=True)
    return crit, out, insts, gts, cfg


def _to_dev(out):
    mv = lambda hd: dict(cls_preds=[t.to(DEV) for t in hd["cls_preds"]], bboxes=[t.to(DEV) for t in hd["bboxes"]])
    d = mv(out)
    d["aux_outputs"] = [mv(a) for a in out["aux_outputs"]]
    return d


@pytest.mark.parametrize("masked_scene", [None, 0])
def test_criterion_gradients_vs_oracle_autograd(masked_scene):
    """ud3d_criterion_layer_grad through train.criterion_backward (three heads; scenes with 5, 0 and 3 ground truths, two
    datasets with different weights; optionally a scene whose GT is fully masked = no matched pair): loss value and the
    gradients w.r.t. every head's logits and boxes against torch.autograd through the oracle criterion (pinned to the
    reference's criterion.py fixtures), on the matches of the GPU matcher (itself checked in test_gpu_ops.py)."""
    from unidet3d_b200 import train
    from oracle import criterion as oc
    names = ["scannet", "s3dis", "scannet"]
    crit, out, insts, gts, cfg = _criterion_case(["scannet", "s3dis"], names, [60, 45, 30], [5, 0, 3], [6, 6, 6], 5, masked_scene)
    dbg = {}
    loss, d_cls, d_box = train.criterion_backward(crit, _to_dev(out), insts, names, debug=dbg)
    heads = out["aux_outputs"] + [dict(cls_preds=out["cls_preds"], bboxes=out["bboxes"])]
    ref, leaves = 0.0, []
    for hd, ms in zip(heads, dbg["matches"]):
        r = dict(cls_preds=[t.clone().requires_grad_(True) for t in hd["cls_preds"]], bboxes=[t.clone().requires_grad_(True) for t in hd["bboxes"]])
        leaves.append(r)
        idx = [tuple(t.cpu() for t in m.nonzero(as_tuple=True)) if m.numel() else (torch.zeros(0, dtype=torch.long),) * 2 for m in ms]
        l, _ = oc.layer_loss(r["cls_preds"], r["bboxes"], gts, names, cfg, indices=idx)
        ref = ref + l
    ref.backward()
    assert abs(float(loss) - float(ref.detach())) < 1e-5 * max(1.0, abs(float(ref.detach())))
    n_box = 0
    for hd, dc, db in zip(leaves, d_cls, d_box):
        for t, g_ in zip(hd["cls_preds"], dc):
            assert torch.allclose(g_.cpu(), t.grad, atol=1e-6, rtol=1e-4)
        for t, g_ in zip(hd["bboxes"], db):
            want = t.grad if t.grad is not None else torch.zeros_like(t)
            assert torch.allclose(g_.cpu(), want, atol=1e-6, rtol=2e-4), float((g_.cpu() - want).abs().max())
            n_box += int((want != 0).any(1).sum())
    assert n_box > 5


def test_criterion_gradients_rotated_boxes(tmp_path):
    """7-parameter (yaw) boxes: the loss value against the oracle criterion, the logit gradients against torch.autograd
    through it, and the box gradients against the host build of the same dual-number templates (tests/harness; they are
    checked against finite differences in tests/test_box_loss_host.py -- mmcv's differentiable rotated IoU is not
    installable here)."""
    import ctypes as C
    import os
    import subprocess
    from unidet3d_b200 import train
    from oracle import criterion as oc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "libbl_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(root, "tests", "harness", "box_loss_host.cpp")], check=True)
    lib = C.CDLL(so)
    names = ["arkitscenes", "scannet"]
    crit, out, insts, gts, cfg = _criterion_case(["scannet", "arkitscenes"], names, [50, 40], [4, 3], [7, 6], 8)
    dbg = {}
    loss, d_cls, d_box = train.criterion_backward(crit, _to_dev(out), insts, names, debug=dbg)
    heads = out["aux_outputs"] + [dict(cls_preds=out["cls_preds"], bboxes=out["bboxes"])]
    ref = 0.0
    for hd, ms, dc, db in zip(heads, dbg["matches"], d_cls, d_box):
        cps = [t.clone().requires_grad_(True) for t in hd["cls_preds"]]
        idx = [tuple(t.cpu() for t in m.nonzero(as_tuple=True)) for m in ms]
        l, _ = oc.layer_loss(cps, hd["bboxes"], gts, names, cfg, indices=idx)
        ref = ref + l.detach()
        l.backward()
        for t, g_ in zip(cps, dc):
            assert torch.allclose(g_.cpu(), t.grad, atol=1e-6, rtol=1e-4)
        # rotated scene (0): expected box gradient = lw_box * w_ds / n_scenes_with_pairs / n_pairs * sum of pair gradients
        iq, ig = idx[0]
        assert len(iq) > 0
        p = np.ascontiguousarray(hd["bboxes"][0][iq].numpy(), np.float32)
        t = np.ascontiguousarray(gts[0]["boxes"][ig].numpy(), np.float32)
        lo, gr = np.zeros(len(iq), np.float32), np.zeros((len(iq), 7), np.float32)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        lib.bl_pair_loss_grad_f32(ptr(p), ptr(t), C.c_int(len(iq)), C.c_int(7), ptr(lo), ptr(gr))
        want = torch.zeros_like(hd["bboxes"][0]).index_add_(0, iq, torch.as_tensor(gr))
        n_has = sum(1 for a, _ in idx if len(a))
        want *= 1.0 * cfg["datasets_weights"][1] / n_has / len(iq)
        assert torch.allclose(db[0].cpu(), want, atol=2e-6, rtol=2e-3), float((db[0].cpu() - want).abs().max())
        assert float(want[:, 6].abs().max()) > 0
    assert abs(float(loss) - float(ref)) < 1e-4 * max(1.0, abs(float(ref)))


def test_criterion_gradients_vs_reference_fixture():
    """The library's criterion gradients (GPU matcher + ud3d_criterion_layer_grad) against the gradients of the
    REFERENCE's own criterion.py under torch.autograd (tests/golden/criterion_grad_ref.npz, generated in the build
    container): det_loss and d det_loss / d (logits, boxes) of three layers x four scenes (two datasets, a scene without
    GT, a GT no query may match, queries matched to several GTs)."""
    import types
    import criterion_grad_case as case
    from unidet3d_b200 import train
    from unidet3d_b200.criterion import UniDet3DCriterion

    def run(names, layers, gts, cfg):
        crit = UniDet3DCriterion(matcher=dict(costs=[dict(type="QueryClassificationCost", weight=cfg["w_cls"]),
                                                     dict(type="BboxCostJointTraining", weight=cfg["w_box"])]),
                                 loss_weight=cfg["loss_weight"], non_object_weight=cfg["non_object_weight"], iter_matcher=True,
                                 bbox_loss_simple=dict(mode="diou"), bbox_loss_rotated=dict(mode="diou"), datasets=cfg["datasets"],
                                 datasets_weights=cfg["datasets_weights"], topk=cfg["topk"])
        insts = [types.SimpleNamespace(labels_3d=g["labels"].to(DEV), query_masks=g["query_masks"].to(DEV),
                                       bboxes_3d=types.SimpleNamespace(gravity_center=g["boxes"][:, :3].to(DEV), tensor=g["boxes"].to(DEV),
                                                                       with_yaw=False)) for g in gts]
        dbg = {}
        loss, d_cls, d_box = train.criterion_backward(crit, _to_dev(dict(layers[0], aux_outputs=layers[1:])), insts, names, debug=dbg)
        # criterion_backward lists the heads aux first, final last; the fixture (and the checker) final first
        order = [len(layers) - 1] + list(range(len(layers) - 1))
        return loss, [d_cls[k] for k in order], [d_box[k] for k in order], [dbg["matches"][k] for k in order]

    case.check(run)
