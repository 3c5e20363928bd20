"""GPU parity tests of every C-ABI op against the CPU oracle (bit-exact for integer work,
<= 1e-3 relative (max-abs error over max-abs reference) for floating point; most ops land ~1e-5)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import voxelize as ovox, rulebook as orb, nms as onms, postprocess as opost, encoder as oenc
from oracle.spconv import sparse_conv, weight_to_koc
from oracle.pool import scatter_mean
from unidet3d_b200.synthetic import make_scene, SCENE_PRESETS

DEV = "cuda"


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def _scenes(preset, B, seed0=0):
    n, v, a, c = SCENE_PRESETS[preset]
    return [make_scene(seed0 + i, n, a, c) for i in range(B)], v


# ------------------------------------------------------------------ voxelise + grid + rulebooks
@pytest.mark.parametrize("preset,B", [("tiny", 3), ("small20k", 2)])
def test_voxelize_and_rulebooks_bit_exact(preset, B):
    from unidet3d_b200 import ops
    scenes, vs = _scenes(preset, B)
    pts = [s[0] for s in scenes]
    ref_coords, ref_feats, ref_inv, ref_shape = ovox.voxelize(pts, vs)
    ref_bc, ref_bf = ovox.point_coords(pts, vs)
    offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device=DEV)
    P = torch.as_tensor(np.concatenate(pts)).to(DEV)
    coords, feats, stats, maxc = ops.point_coords(P, offs, vs)
    assert np.array_equal(coords.cpu().numpy(), ref_bc)
    assert relerr(feats, ref_bf) < 1e-6
    shape = np.maximum(maxc.cpu().numpy() + 1, 128)
    assert np.array_equal(shape, ref_shape)
    dims = [B] + [int(x) + 1 for x in maxc.cpu().numpy()]
    grid = ops.Grid(dims, DEV)
    n1 = int(grid.build(coords).item())
    assert n1 == len(ref_coords)
    inv = grid.rank(coords)
    assert np.array_equal(inv.cpu().numpy().astype(np.int64), ref_inv)
    vox = grid.coords(n1)
    assert np.array_equal(vox.cpu().numpy(), ref_coords)
    vf = ops.voxel_mean(feats, inv, n1)
    assert relerr(vf, ref_feats) < 1e-5
    # SubM table (canonical order) + tile mask
    table, mask = ops.rulebook_subm3(vox, grid, canonical=True)
    ref_table = orb.subm3_table(ref_coords, ref_shape)
    assert np.array_equal(table.cpu().numpy(), ref_table)
    m = mask.cpu().numpy().view(np.uint32)
    for t in range(len(m)):
        act = (ref_table[:, t * 128:(t + 1) * 128] >= 0).any(1)
        assert m[t] == sum(1 << k for k in range(27) if act[k])
    # non-canonical row order
    perm = torch.randperm(n1, device=DEV)
    table_p, _ = ops.rulebook_subm3(vox[perm].contiguous(), grid, canonical=False)
    assert np.array_equal(table_p.cpu().numpy(), orb.subm3_table(ref_coords[perm.cpu().numpy()], ref_shape))
    # strided pyramid, 4 levels down
    c, s, rc, rs = vox, [int(x) for x in shape], ref_coords, ref_shape
    for lvl in range(4):
        parents = ops.down2_parents(c, s)
        out_shape = [(x - 2) // 2 + 1 for x in s]
        cg = ops.Grid([B] + [max(1, min(o, (d + 1) // 2)) for o, d in zip(out_shape, dims[1:])], DEV)
        dims = cg.dims
        nc = int(cg.build(parents).item())
        ref_cc, ref_child, ref_up, ref_os = orb.down2(rc, rs)
        assert nc == len(ref_cc) and list(ref_os) == out_shape
        cc = cg.coords(nc)
        assert np.array_equal(cc.cpu().numpy(), ref_cc)
        child, up, cm, um = ops.rulebook_down2(c, parents, nc, cg)
        assert np.array_equal(child.cpu().numpy(), ref_child)
        assert np.array_equal(up.cpu().numpy(), ref_up)
        um_np = um.cpu().numpy().view(np.uint32)
        for t in range(len(um_np)):
            act = (ref_up[:, t * 128:(t + 1) * 128] >= 0).any(1)
            assert um_np[t] == sum(1 << k for k in range(8) if act[k])
        c, s, rc, rs = cc, out_shape, ref_cc, ref_os


def test_down2_drops_odd_boundary():
    from unidet3d_b200 import ops
    rng = np.random.default_rng(0)
    shape = [9, 7, 6]
    cc = np.unique(rng.integers(0, shape, (200, 3)), axis=0)
    coords = np.concatenate([np.zeros((len(cc), 1), np.int64), cc], 1).astype(np.int32)
    coords = coords[rng.permutation(len(coords))]
    c = torch.as_tensor(coords).to(DEV)
    parents = ops.down2_parents(c, shape)
    cg = ops.Grid([1, 4, 3, 3], DEV)
    nc = int(cg.build(parents).item())
    ref_cc, ref_child, ref_up, _ = orb.down2(coords, shape)
    assert nc == len(ref_cc)
    child, up, _, _ = ops.rulebook_down2(c, parents, nc, cg)
    assert np.array_equal(child.cpu().numpy(), ref_child) and np.array_equal(up.cpu().numpy(), ref_up)
    assert (ref_up < 0).all(0).any()          # some fine rows really are dropped


# ------------------------------------------------------------------ gather-GEMM
def _rand_table(rng, K, n_out, n_in, density=0.4):
    t = rng.integers(0, n_in, (K, n_out)).astype(np.int32)
    t[rng.random((K, n_out)) > density] = -1
    return t


@pytest.mark.parametrize("c_in,c_out,K,n_in,n_out", [
    (32, 32, 27, 700, 700), (64, 64, 27, 300, 300), (6, 32, 27, 500, 500), (96, 96, 27, 260, 260),
    (128, 128, 27, 130, 130), (160, 160, 27, 200, 200), (64, 32, 8, 400, 900), (32, 64, 8, 900, 300),
    (320, 160, 27, 150, 150), (8, 40, 27, 333, 333), (192, 96, 1, 257, 257)])
def test_gemm_sparse_conv(c_in, c_out, K, n_in, n_out):
    from unidet3d_b200 import ops
    rng = np.random.default_rng(c_in * 1000 + c_out)
    g = torch.Generator().manual_seed(c_in + c_out)
    x = torch.randn(n_in, c_in, generator=g)
    w = torch.randn(c_out, K, c_in, generator=g) / (c_in * 3) ** 0.5
    table = _rand_table(rng, K, n_out, n_in)
    scale, shift = torch.rand(c_in, generator=g) + 0.5, torch.randn(c_in, generator=g) * 0.2
    res = torch.randn(n_out, c_out, generator=g)
    w_koc = w.permute(1, 2, 0).contiguous()
    ref = sparse_conv(torch.relu(x * scale + shift), table, w_koc) + res
    xd, td = x.to(DEV), torch.as_tensor(table).to(DEV)
    pw = ops.PackedWeight(w.to(DEV))
    kw = dict(table=td, in_scale=scale.to(DEV), in_shift=shift.to(DEV), in_relu=True, residual=res.to(DEV))
    out = ops.gemm(xd, pw, **kw)
    out_simt = ops.gemm_simt(xd, w.to(DEV), **kw)
    assert relerr(out_simt, ref) < 1e-5
    assert relerr(out, ref) < 2e-4, relerr(out, ref)
    # no prologue, no residual
    ref2 = sparse_conv(x, table, w_koc)
    assert relerr(ops.gemm(xd, pw, table=td), ref2) < 2e-4


def test_gemm_strided_views_and_tile_mask():
    """concat-free U-Net plumbing: read/write column slices of wider buffers; per-tile offset skipping."""
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(3)
    rng = np.random.default_rng(3)
    n, c = 515, 32
    buf = torch.randn(n, 2 * c, generator=g).to(DEV)
    w = (torch.randn(c, 27, c, generator=g) / 10).to(DEV)
    table = _rand_table(rng, 27, n, n, 0.3)
    table[5] = -1
    table[20, 128:256] = -1
    mask = np.zeros((n + 127) // 128, np.uint32)
    for t in range(len(mask)):
        act = (table[:, t * 128:(t + 1) * 128] >= 0).any(1)
        mask[t] = sum(1 << k for k in range(27) if act[k])
    td, md = torch.as_tensor(table).to(DEV), torch.as_tensor(mask.view(np.int32)).to(DEV)
    pw = ops.PackedWeight(w)
    ref = sparse_conv(buf[:, c:].cpu(), table, w.cpu().permute(1, 2, 0).contiguous())
    outbuf = torch.zeros(n, 2 * c, device=DEV)
    ops.gemm(buf[:, c:], pw, table=td, tile_mask=md, out=outbuf[:, :c])
    assert relerr(outbuf[:, :c], ref) < 2e-4
    assert float(outbuf[:, c:].abs().max()) == 0.0


@pytest.mark.parametrize("T,c_in,c_out,act", [(300, 256, 768, None), (1000, 256, 1024, "gelu"), (257, 1024, 256, None),
                                              (129, 32, 256, "relu"), (64, 256, 19, None), (500, 256, 100, None),
                                              (77, 256, 8, None)])
def test_gemm_dense_linear(T, c_in, c_out, act):
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(T)
    x = torch.randn(T, c_in, generator=g)
    w = torch.randn(c_out, c_in, generator=g) / c_in ** 0.5
    b = torch.randn(c_out, generator=g)
    res = torch.randn(T, c_out, generator=g)
    y = x @ w.t() + b
    y = torch.relu(y) if act == "relu" else torch.nn.functional.gelu(y) if act == "gelu" else y
    y = y + res
    out = ops.gemm(x.to(DEV), ops.PackedWeight(w.to(DEV)), bias=b.to(DEV), act=act, residual=res.to(DEV))
    assert relerr(out, y) < 2e-4, relerr(out, y)


# ------------------------------------------------------------------ pooling / LN / attention / heads
def test_segmented_mean():
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(0)
    n_vox, n_pts, n_seg = 5000, 40000, 700
    feats = torch.randn(n_vox, 32, generator=g)
    inv = torch.randint(0, n_vox, (n_pts,), generator=g)
    seg = torch.randint(0, n_seg - 3, (n_pts,), generator=g)       # last ids never occur -> zero rows
    seg = torch.sort(seg)[0][torch.randperm(n_pts, generator=g)] if False else seg
    scale, shift = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.3
    ref = scatter_mean(torch.relu(feats * scale + shift)[inv], seg, n_seg)
    out = ops.segmented_mean(feats.to(DEV), seg.to(DEV), n_seg, gather=inv.int().to(DEV), scale=scale.to(DEV),
                             shift=shift.to(DEV), relu=True)
    assert relerr(out, ref) < 1e-5
    assert float(out[-3:].abs().max()) == 0.0
    # deterministic: the fixed-point atomics make repeated calls bit-identical (random ids: ~every point is its own run)
    for _ in range(5):
        again = ops.segmented_mean(feats.to(DEV), seg.to(DEV), n_seg, gather=inv.int().to(DEV), scale=scale.to(DEV),
                                   shift=shift.to(DEV), relu=True)
        assert torch.equal(again, out)
    # superpoint centres: C=3 of a [n,6] matrix, runs of equal ids
    pts = torch.randn(n_pts, 6, generator=g)
    seg2 = torch.sort(seg)[0]
    ref2 = scatter_mean(pts[:, :3], seg2, n_seg)
    out2 = ops.segmented_mean(pts.to(DEV), seg2.to(DEV), n_seg, channels=3)
    assert relerr(out2, ref2) < 1e-5


def test_layernorm():
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(1)
    for C in (256, 128, 64):
        x, r = torch.randn(333, C, generator=g) * 3, torch.randn(333, C, generator=g)
        gam, bet = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
        ref = torch.nn.functional.layer_norm(x + r, (C,), gam, bet, 1e-5)
        out = ops.layernorm(x.to(DEV), gam.to(DEV), bet.to(DEV), residual=r.to(DEV))
        assert relerr(out, ref) < 1e-5


@pytest.mark.parametrize("lens", [[37, 21, 50], [300, 1], [1000, 64, 65, 129]])
def test_attention(lens):
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(sum(lens))
    H, d = 8, 256
    Tt = sum(lens)
    qkv = torch.randn(Tt, 3 * d, generator=g) * 1.5
    cu = torch.tensor(np.cumsum([0] + lens), dtype=torch.int32)
    ref = []
    for i, T in enumerate(lens):
        s = qkv[cu[i]:cu[i + 1]]
        q, k, v = [s[:, j * d:(j + 1) * d].view(T, H, 32).transpose(0, 1) for j in range(3)]
        a = torch.softmax(q @ k.transpose(1, 2) / 32 ** 0.5, -1) @ v
        ref.append(a.transpose(0, 1).reshape(T, d))
    ref = torch.cat(ref)
    out = ops.attention(qkv.to(DEV), cu.to(DEV), max(lens), H)
    assert relerr(out, ref) < 1e-4, relerr(out, ref)


def test_bbox_decode_and_gather_columns():
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(2)
    raw = torch.randn(100, 8, generator=g) * 0.5
    cen = torch.randn(100, 3, generator=g)
    for ang in (False, True):
        b = torch.hstack((torch.exp(raw[:, :6]), raw[:, 6:]))
        ref = oenc.bbox_pred_to_bbox(cen, b if ang else b[:, :6])
        out = ops.bbox_decode(raw.to(DEV), cen.to(DEV), ang)
        assert relerr(out, ref) < 1e-5
    src = torch.randn(50, 100, generator=g)
    cols = torch.tensor([5, 99, 0, 17], dtype=torch.int32)
    assert torch.equal(ops.gather_columns(src.to(DEV), cols.to(DEV)).cpu(), src[:, cols.long()])


# ------------------------------------------------------------------ post-processing
@pytest.mark.parametrize("T,C,k", [(90, 5, 300), (2000, 18, 1000), (4096, 84, 1000), (60, 17, 1000)])
def test_topk(T, C, k):
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(T)
    logits = torch.randn(T, C + 1, generator=g) * 2
    _, rs, rl, rq = opost.topk_candidates(logits, torch.zeros(T, 6), k)
    s, l, q = ops.topk_scores(logits.to(DEV), k)
    assert np.allclose(s.cpu().numpy(), rs.numpy(), rtol=1e-5, atol=0)
    # identical selection wherever scores are not within rounding of each other
    same = (l.cpu().long() == rl) & (q.cpu().long() == rq)
    assert same.float().mean() > 0.995
    assert torch.all(s[:-1] >= s[1:])


def _rand_boxes(rng, n, yaw):
    b = np.concatenate([rng.uniform(0, 5, (n, 3)), rng.uniform(0.3, 1.6, (n, 3))], 1)
    if yaw:
        b = np.concatenate([b, rng.uniform(-3.2, 3.2, (n, 1))], 1)
    return b.astype(np.float32)


@pytest.mark.parametrize("mode,yaw", [(0, True), (1, False), (2, False)])
def test_nms_multiclass(mode, yaw):
    from unidet3d_b200 import ops
    rng = np.random.default_rng(mode)
    n = 1000
    boxes = _rand_boxes(rng, n, yaw)
    scores = np.sort(rng.random(n).astype(np.float32))[::-1].copy()
    scores[-5:] = 0.0                                   # score_thr=0 filters these
    labels = rng.integers(0, 7, n).astype(np.int32)
    thr = 0.3
    rb, rs, rl, ri = opost.multiclass_nms(torch.as_tensor(boxes), torch.as_tensor(scores), torch.as_tensor(labels).long(),
                                          fast_nms=(mode == 1), iou_thr=thr)
    keep, nk = ops.nms_multiclass(torch.as_tensor(boxes).to(DEV), torch.as_tensor(scores).to(DEV),
                                  torch.as_tensor(labels).to(DEV), mode, thr)
    k = keep[: int(nk.item())].cpu().long()
    assert len(k) == len(ri) and torch.equal(k, ri)


def test_trim_boxes():
    from unidet3d_b200 import ops
    rng = np.random.default_rng(4)
    pts, sp = make_scene(0, 20000, 21.0, 0.5)
    xyz = torch.as_tensor(pts[:, :3] - pts[:, :3].min(0))
    n_sp = int(sp.max()) + 1
    M = 40
    ext = xyz.max(0).values.numpy()
    boxes = np.concatenate([rng.uniform(0, 1, (M, 3)) * ext, rng.uniform(0.3, 1.5, (M, 3)), rng.uniform(-3, 3, (M, 1))], 1).astype(np.float32)
    for bd in (7, 6):
        b = torch.as_tensor(boxes[:, :bd].copy())
        ref = opost.trim_bboxes_by_superpoints(torch.as_tensor(sp), xyz, b, 0.18, 0.81)
        P = torch.as_tensor(pts).to(DEV).clone()
        P[:, :3] = xyz.to(DEV)
        out = ops.trim_boxes(P, torch.as_tensor(sp).to(DEV), n_sp, b.to(DEV), 0.18, 0.81).cpu()
        fin = torch.isfinite(ref)
        assert torch.equal(torch.isfinite(out), fin)
        close = torch.isclose(out[fin], ref[fin], rtol=1e-5, atol=1e-6)
        assert close.float().mean() > 0.98     # a borderline point (|face distance| ~ 1 ulp) may flip a vote


# ------------------------------------------------------------------ operand-form (pre-split) feature maps
from opform import split_encode, split_decode  # noqa: E402


def test_act_split_roundtrip():
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(777, 96, generator=g) * 3
    sc, sh = torch.rand(96, generator=g) + 0.5, torch.randn(96, generator=g)
    out = ops.act_split(x.to(DEV), sc.to(DEV), sh.to(DEV), relu=True)
    ref = torch.relu(x * sc + sh)
    assert torch.equal(out.cpu().view(torch.int32), split_encode(ref).view(torch.int32)) or relerr(split_decode(out.cpu()), ref) < 1e-5
    assert relerr(split_decode(out.cpu()), ref) < 2e-5


@pytest.mark.parametrize("c_in,c_out,K,n", [(32, 32, 27, 1500), (64, 64, 27, 700), (64, 32, 27, 900), (128, 128, 27, 300),
                                            (160, 160, 27, 140), (32, 64, 8, 2000), (96, 64, 8, 500)])
def test_gemm_operand_form_in_and_out(c_in, c_out, K, n):
    """pre-activated split input (cp.async gather) + operand-form outputs for two consumer BatchNorms,
    with and without split-K (small n -> gridDim.z > 1)."""
    from unidet3d_b200 import ops
    rng = np.random.default_rng(n + c_in)
    g = torch.Generator().manual_seed(n + c_out)
    x = torch.relu(torch.randn(n, c_in, generator=g))
    w = torch.randn(c_out, K, c_in, generator=g) / (c_in * 3) ** 0.5
    table = _rand_table(rng, K, n, n, 0.35)
    res = torch.randn(n, c_out, generator=g)
    ref = sparse_conv(x, table, w.permute(1, 2, 0).contiguous()) + res
    s1, h1 = torch.rand(c_out, generator=g) + 0.5, torch.randn(c_out, generator=g) * 0.3
    s2, h2 = torch.rand(c_out, generator=g) + 0.5, torch.randn(c_out, generator=g) * 0.3
    xs = split_encode(x).to(DEV)
    a1 = torch.zeros(n, c_out, device=DEV)
    wide = torch.zeros(n, 2 * c_out, device=DEV)
    out = ops.gemm(xs, ops.PackedWeight(w.to(DEV)), table=torch.as_tensor(table).to(DEV), in_split=True,
                   residual=res.to(DEV), acts=[(a1, s1.to(DEV), h1.to(DEV)), (wide[:, c_out:], s2.to(DEV), h2.to(DEV))])
    assert relerr(out, ref) < 2e-4, relerr(out, ref)
    assert relerr(split_decode(a1.cpu()), torch.relu(ref * s1 + h1)) < 2e-4
    assert relerr(split_decode(wide[:, c_out:].cpu().contiguous()), torch.relu(ref * s2 + h2)) < 2e-4
    assert float(wide[:, :c_out].abs().max()) == 0.0
    # no_raw: only the operand-form output is produced
    a3 = torch.zeros(n, c_out, device=DEV)
    ops.gemm(xs, ops.PackedWeight(w.to(DEV)), table=torch.as_tensor(table).to(DEV), in_split=True, no_raw=True,
             acts=[(a3, s1.to(DEV), h1.to(DEV))])
    assert relerr(split_decode(a3.cpu()), torch.relu((ref - res) * s1 + h1)) < 2e-4


@pytest.mark.parametrize("lens", [[37, 21, 50], [300, 1], [1000, 64, 65, 129, 128], [1700, 1650]])
def test_attention_operand_form(lens):
    """q|k|v and the output in operand form (what the encoder uses): cp.async K/V ring + ldmatrix.trans path."""
    from unidet3d_b200 import ops
    g = torch.Generator().manual_seed(sum(lens) + 1)
    H, d = 8, 256
    Tt = sum(lens)
    qkv = torch.randn(Tt, 3 * d, generator=g) * 1.5
    cu = torch.tensor(np.cumsum([0] + lens), dtype=torch.int32)
    ref = []
    for i, T in enumerate(lens):
        s = qkv[cu[i]:cu[i + 1]]
        q, k, v = [s[:, j * d:(j + 1) * d].view(T, H, 32).transpose(0, 1) for j in range(3)]
        a = torch.softmax(q @ k.transpose(1, 2) / 32 ** 0.5, -1) @ v
        ref.append(a.transpose(0, 1).reshape(T, d))
    ref = torch.cat(ref)
    out_s = ops.attention(split_encode(qkv).to(DEV), cu.to(DEV), max(lens), H, split_in=True)
    assert relerr(split_decode(out_s.cpu()), ref) < 1e-4, relerr(split_decode(out_s.cpu()), ref)
    out_s2 = ops.attention(qkv.to(DEV), cu.to(DEV), max(lens), H, split_out=True)
    assert relerr(split_decode(out_s2.cpu()), ref) < 1e-4
    # tcgen05 kernel (the product path: S / O' / row sums in TMEM, P TMEM-resident, TMA tensor-map loads)
    out_s3 = ops.attention(split_encode(qkv).to(DEV), cu.to(DEV), max(lens), H, split_in=True, tcgen05=True)
    assert relerr(split_decode(out_s3.cpu()), ref) < 1e-4, relerr(split_decode(out_s3.cpu()), ref)


@pytest.mark.gpu
def test_subm3_tile_order_is_a_pure_regrouping():
    """ud3d_subm3_tile_order: perm is a permutation, table_p / tile_mask_p describe the same rulebook in the regrouped
    order, tiles see fewer active offsets, and a convolution through (table_p, tile_mask_p, row_perm) is bit-identical
    to the canonical-order call (fp32 and operand-form paths, with residual and operand-form outputs)."""
    from unidet3d_b200 import ops
    from unidet3d_b200.rulebook import build_pyramid
    scenes, _ = _scenes("small20k", 2, seed0=5)
    pts = torch.as_tensor(np.concatenate([s[0] for s in scenes])).cuda()
    offs = torch.tensor(np.cumsum([0] + [len(s[0]) for s in scenes]), dtype=torch.int32, device="cuda")
    coords_pt, _, _, maxc = ops.point_coords(pts, offs, 0.02)          # 2 cm voxels: ~20k voxels, 150+ tiles
    ext = (maxc.cpu().numpy() + 1).tolist()
    pyr = build_pyramid(None, [max(e, 128) for e in ext], 2, 1, extents=ext, seed_coords=coords_pt)
    lv = pyr.levels[0]
    perm, table_p, mask_p = ops.subm3_tile_order(lv.subm)
    torch.cuda.synchronize()
    n = lv.n
    perm_h, tab_h = perm.cpu().numpy(), lv.subm.cpu().numpy()
    assert np.array_equal(np.sort(perm_h), np.arange(n))
    assert np.array_equal(table_p.cpu().numpy(), tab_h[:, perm_h])

    def tile_masks(t):
        v = t >= 0
        pad = (-n) % 128
        v = np.concatenate([v, np.zeros((27, pad), bool)], 1).reshape(27, -1, 128).any(2)
        return (v * (1 << np.arange(27))[:, None]).sum(0).astype(np.uint32)
    ref_mask_p = tile_masks(tab_h[:, perm_h])
    assert np.array_equal(mask_p.cpu().numpy().view(np.uint32), ref_mask_p)
    pop = lambda m: np.array([bin(int(x)).count("1") for x in m]).mean()
    assert pop(ref_mask_p) < 0.85 * pop(tile_masks(tab_h)), (pop(ref_mask_p), pop(tile_masks(tab_h)))
    # bit-identical convolutions
    g = torch.Generator(device="cuda").manual_seed(0)
    c = 32
    x = torch.randn(n, c, device="cuda", generator=g)
    w = ops.PackedWeight(torch.randn(c, 27, c, device="cuda", generator=g) * 0.1)
    res = torch.randn(n, c, device="cuda", generator=g)
    sc = torch.rand(c, device="cuda", generator=g) + 0.5
    sh = torch.randn(c, device="cuda", generator=g) * 0.1
    y0 = ops.gemm(x, w, table=lv.subm, tile_mask=lv.subm_mask, in_scale=sc, in_shift=sh, in_relu=True, residual=res)
    y1 = ops.gemm(x, w, table=table_p, tile_mask=mask_p, in_scale=sc, in_shift=sh, in_relu=True, residual=res, row_perm=perm)
    assert torch.equal(y0, y1)
    xs = ops.act_split(x, sc, sh, relu=True)
    a0, a1 = torch.empty_like(x), torch.empty_like(x)
    z0 = ops.gemm(xs, w, table=lv.subm, tile_mask=lv.subm_mask, in_split=True, residual=res, acts=[(a0, sc, sh)])
    z1 = ops.gemm(xs, w, table=table_p, tile_mask=mask_p, in_split=True, residual=res, acts=[(a1, sc, sh)], row_perm=perm)
    assert torch.equal(z0, z1) and torch.equal(a0, a1)
    rel = float((z0 - y0).abs().max() / y0.abs().max())
    assert rel < 1e-3, rel


# ------------------------------------------------------------------ training-side targets / matcher / loss values (R14)
def _criterion_case(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "criterion_ref.npz"))
    names = [str(n) for n in g["names"]]
    cfg = dict(datasets=["scannet", "s3dis", "arkitscenes"], datasets_weights=[1.0, 0.7, 1.3], topk=[6, 4, 5],
               loss_weight=[0.5, 1.0], non_object_weight=0.1, w_cls=0.5, w_box=2.0, iter_matcher=True)
    return g, names, cfg


def test_criterion_layer_vs_reference_fixture(golden_dir):
    """ud3d_criterion_layer against the reference's own criterion.py (tests/golden/criterion_ref.npz): matched
    (query, gt) pairs bit-exact, per-layer loss and det_loss within 1e-4 relative; axis-aligned and rotated boxes,
    a GT with fewer than topk+1 candidate queries, a scene without GT."""
    import unidet3d_b200 as u
    from unidet3d_b200.structures import DepthInstance3DBoxes, InstanceData
    g, names, cfg = _criterion_case(golden_dir)
    diou = lambda t: dict(type=t, mode="diou", reduction="none")
    crit = u.MODELS.build(dict(
        type="UniDet3DCriterion",
        matcher=dict(type="UniMatcher", costs=[dict(type="QueryClassificationCost", weight=0.5),
                                               dict(type="BboxCostJointTraining", weight=2.0)]),
        loss_weight=cfg["loss_weight"], non_object_weight=0.1, iter_matcher=True,
        bbox_loss_simple=diou("UniDet3DAxisAlignedIoULoss"), bbox_loss_rotated=diou("UniDet3DRotatedIoU3DLoss"),
        datasets=cfg["datasets"], datasets_weights=cfg["datasets_weights"], topk=cfg["topk"]))
    insts = []
    for i in range(len(names)):
        gb = torch.as_tensor(g[f"gt_boxes{i}"]).to(DEV)
        dim = gb.shape[1] if gb.numel() else (7 if names[i] == "arkitscenes" else 6)
        insts.append(InstanceData(labels_3d=torch.as_tensor(g[f"gt_labels{i}"]).to(DEV),
                                  bboxes_3d=DepthInstance3DBoxes(gb, box_dim=dim, with_yaw=dim == 7, origin=(0.5, 0.5, 0.5)),
                                  query_masks=torch.as_tensor(g[f"qmask{i}"]).to(DEV)))
    layers = [dict(cls_preds=[torch.as_tensor(g[f"l{l}_cls{i}"]).to(DEV) for i in range(len(names))],
                   bboxes=[torch.as_tensor(g[f"l{l}_box{i}"]).to(DEV) for i in range(len(names))]) for l in range(3)]
    for l, lay in enumerate(layers):
        terms = crit.layer_terms(lay, insts, names)
        for i, (match, sums) in enumerate(terms):
            if f"l{l}_iq{i}" in g.files:
                ids = torch.argwhere(match).cpu().numpy()
                assert np.array_equal(ids[:, 0], g[f"l{l}_iq{i}"]) and np.array_equal(ids[:, 1], g[f"l{l}_ig{i}"]), (l, i)
            else:
                assert match.numel() == 0 and float(sums[3]) == 0
        loss = float(crit.get_layer_loss(lay, insts, names))
        ref = float(g[f"layer_loss{l}"])
        assert abs(loss - ref) < 1e-4 * abs(ref), (l, loss, ref)
    det = float(crit(dict(layers[0], aux_outputs=layers[1:]), insts, names)["det_loss"])
    assert abs(det - float(g["det_loss"])) < 1e-4 * abs(float(g["det_loss"]))


def test_criterion_layer_vs_oracle_large():
    """full-size (T = 3000 queries = query_thr, G = 40) scene against the CPU oracle: same matches (up to costs within
    1e-4 of a column threshold), same loss sums."""
    from unidet3d_b200 import ops
    from oracle import criterion as oc
    rng = np.random.default_rng(3)
    for dim, C, topk in [(6, 18, 6), (7, 17, 6)]:
        T, G = 3000, 40
        gt = np.concatenate([rng.uniform(0.5, 7.5, (G, 3)), rng.uniform(0.3, 2.0, (G, 3))] +
                            ([rng.uniform(-3, 3, (G, 1))] if dim == 7 else []), 1).astype(np.float32)
        pb = np.concatenate([rng.uniform(0.5, 7.5, (T, 3)), rng.uniform(0.3, 2.0, (T, 3))] +
                            ([rng.uniform(-3, 3, (T, 1))] if dim == 7 else []), 1).astype(np.float32)
        pb[:G * 8] = np.repeat(gt, 8, 0) + 0.08 * rng.standard_normal((G * 8, dim)).astype(np.float32)
        pb[:, 3:6] = np.abs(pb[:, 3:6]) + 0.05
        cls = (rng.standard_normal((T, C + 1)) * 2).astype(np.float32)
        labels = rng.integers(0, C, G)
        qm = rng.random((G, T)) < 0.3
        match, sums = ops.criterion_layer(torch.as_tensor(cls).to(DEV), torch.as_tensor(pb).to(DEV), torch.as_tensor(gt).to(DEV),
                                          torch.as_tensor(labels).to(DEV), torch.as_tensor(qm).to(DEV), topk, 0.5, 2.0, 0.1)
        # matched set: identical to the oracle's except where a cost lies within fp noise of its column's threshold
        # (the rotated BEV intersection is evaluated with different float operations on the two sides)
        cost = oc.match_cost(torch.as_tensor(cls), torch.as_tensor(pb), torch.as_tensor(labels), torch.as_tensor(gt))
        cost = torch.where(torch.as_tensor(qm).T, cost, torch.tensor(1e8))
        kth = torch.topk(cost, topk + 1, dim=0, largest=False).values[-1:]
        ref = cost < kth
        m = match.cpu()
        borderline = (cost - kth).abs() < 1e-4 * kth.abs().clamp(min=1.0)
        assert bool(((m == ref) | borderline).all()), torch.argwhere((m != ref) & ~borderline)[:8]
        assert int((m != ref).sum()) <= 2
        # loss terms: recomputed by the oracle for the matched set the kernel produced
        ids = torch.argwhere(m)
        iq, ig = ids[:, 0], ids[:, 1]
        tgt = torch.full((T,), C, dtype=torch.long)
        tgt[iq] = torch.as_tensor(labels)[ig]
        w = torch.tensor([1.0] * C + [0.1])
        ce_num = float((w[tgt] * torch.nn.functional.cross_entropy(torch.as_tensor(cls), tgt, reduction="none")).sum())
        box = float(oc.box_loss(torch.as_tensor(pb)[iq], torch.as_tensor(gt)[ig]).sum())
        s = sums.cpu().numpy()
        assert abs(s[0] - ce_num) < 1e-4 * ce_num and abs(s[1] - float(w[tgt].sum())) < 1e-3
        assert abs(s[2] - box) < 1e-4 * abs(box) + 1e-4 and int(s[3]) == len(iq)


def test_gt_target_helpers_vs_reference_fixture(golden_dir):
    """ud3d_boxes_by_instance / ud3d_targets_by_distance against the reference's get_bboxes_by_masks / get_targets."""
    from unidet3d_b200 import ops
    g, _, _ = _criterion_case(golden_dir)
    bb = ops.boxes_by_instance(torch.as_tensor(g["bm_points"]).to(DEV), torch.as_tensor(g["bm_inst"]).to(DEV), len(g["bm_boxes"]))
    assert np.array_equal(bb.cpu().numpy(), g["bm_boxes"])
    tg = ops.targets_by_distance(torch.as_tensor(g["tg_centers"]).to(DEV), torch.as_tensor(g["tg_boxes"]).to(DEV), 6)
    assert np.array_equal(tg.cpu().numpy(), g["tg_masks"])


@pytest.mark.parametrize("c_in,c_out,K,n", [(32, 32, 27, 40000), (64, 64, 27, 24000), (64, 32, 27, 24000), (128, 64, 8, 30000),
                                            (256, 768, 1, 20000), (1024, 256, 1, 20000), (32, 32, 27, 19000 + 77)])
def test_gemm_tmem_operand_kernel(c_in, c_out, K, n):
    """The experimental kernel variant whose A operand goes global -> registers -> TMEM (gemm_ts.cu; input in the
    interleaved operand form, in_split=2): same contract and tolerances as the shared-memory-operand kernel, with a
    rulebook + tile mask, a ragged last tile, and as a dense GEMM."""
    from unidet3d_b200 import ops
    rng = np.random.default_rng(n + c_in)
    g = torch.Generator().manual_seed(n + c_out)
    x = torch.relu(torch.randn(n, c_in, generator=g))
    w = torch.randn(c_out, K, c_in, generator=g) / (c_in * 3) ** 0.5
    res = torch.randn(n, c_out, generator=g)
    s1, h1 = torch.rand(c_out, generator=g) + 0.5, torch.randn(c_out, generator=g) * 0.3
    xs = split_encode(x).to(DEV)
    xs_il = split_encode(x, interleaved=True).to(DEV)
    assert torch.equal(ops.operand_form_interleave(xs).view(torch.int32), xs_il.view(torch.int32))
    if K > 1:
        table = _rand_table(rng, K, n, n, 0.35)
        # make some (tile, offset) pairs empty so that the tile masks differ between tiles
        for t in range(0, (n + 127) // 128, 3):
            table[rng.integers(0, K, 5), t * 128:(t + 1) * 128] = -1
        v = table >= 0
        pad = (-n) % 128
        v = np.concatenate([v, np.zeros((K, pad), bool)], 1).reshape(K, -1, 128).any(2)
        mask = torch.as_tensor((v * (1 << np.arange(K))[:, None]).sum(0).astype(np.int64).astype(np.int32)).to(DEV)
        ref = sparse_conv(x, table, w.permute(1, 2, 0).contiguous()) + res
        tb = torch.as_tensor(table).to(DEV)
    else:
        tb, mask = None, None
        ref = x @ w[:, 0].t() + res
    outs = []
    for fl, xin in ((1, xs), (2, xs_il)):
        a1 = torch.zeros(n, c_out, device=DEV)
        out = ops.gemm(xin, ops.PackedWeight(w.to(DEV)), table=tb, tile_mask=mask, in_split=fl, residual=res.to(DEV),
                       acts=[(a1, s1.to(DEV), h1.to(DEV))])
        torch.cuda.synchronize()
        assert relerr(out, ref) < 2e-4, (fl, relerr(out, ref))
        assert relerr(split_decode(a1.cpu()), torch.relu(ref * s1 + h1)) < 2e-4
        outs.append(out)
    assert relerr(outs[1], outs[0]) < 1e-5


def test_collate_with_elastic_coords():
    """UniDet3D.collate with the ElasticTransfrom side input (unidet3d.py:162-166): voxel coordinates from the distorted
    coordinates (already in voxel units), features from the original points -- bit-exact vs the oracle."""
    import unidet3d_b200 as u
    from unidet3d_b200 import configs
    cfg = configs.model_cfg(("scannet",), topk_insts=100)
    n, v, a, c = SCENE_PRESETS["tiny"]
    cfg["voxel_size"] = v
    model = u.MODELS.build(cfg).eval().to(DEV)
    rng = np.random.default_rng(5)
    pts = [make_scene(40 + i, n + 50 * i, a, c)[0] for i in range(3)]
    # smooth distortion of the voxel-unit coordinates, like transforms_3d.py:39-43
    els = [(p[:, :3] / v + 3.0 * np.sin(p[:, :3] * 2.0 + rng.uniform(0, 6, 3))).astype(np.float32) for p in pts]
    P = torch.as_tensor(np.concatenate(pts)).to(DEV)
    offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device=DEV)
    # float32 (augmentation skipped) and float64 (applied; values snapped next to integers so that the precision of the
    # subtraction decides the voxel)
    els64 = [e.astype(np.float64) + 1e-9 * rng.standard_normal(e.shape) for e in els]
    for k in range(3):
        els64[k][::7] = np.round(els64[k][::7]) - 1e-11 + els64[k].min(0)
    for variant in (els, els64):
        coords, feats, inverse, shape = ovox.voxelize(pts, v, 128, elastic_list=variant)
        E = torch.as_tensor(np.concatenate(variant)).to(DEV)
        x, inv = model.collate(P, offs, 3, E)
        assert np.array_equal(x.indices.cpu().numpy(), coords)
        assert np.array_equal(inv.cpu().numpy().astype(np.int64), inverse)
        assert x.spatial_shape == list(shape)
        assert relerr(x.features, feats) < 1e-5
    # and it differs from the plain voxelisation
    x0, _ = model.collate(P, offs, 3)
    assert x0.indices.shape != x.indices.shape or not torch.equal(x0.indices, x.indices)
