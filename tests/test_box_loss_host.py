"""CPU: the dual-number DIoU templates of unidet3d_b200/csrc/box_loss.cuh (what the GPU criterion-gradient kernel
instantiates), built for the host with g++ by this test: values against the oracle's DIoU losses (pinned to the
reference's criterion fixtures), gradients against torch.autograd (axis-aligned) and against central finite differences of
the same templates in double (rotated: the reference differentiates mmcv's oriented_box_intersection_2d with autograd,
which is not installable here -- the polygon-area gradient is the same function's derivative)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import criterion as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("bl") / "libbl_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "harness", "box_loss_host.cpp")],
                   check=True)
    return C.CDLL(so)


def _call(lib, name, pred, tgt, dtype, with_grad=True):
    pred, tgt = np.ascontiguousarray(pred, dtype), np.ascontiguousarray(tgt, dtype)
    n, dim = pred.shape
    loss, grad = np.zeros(n, dtype), np.zeros((n, dim), dtype)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    if with_grad:
        getattr(lib, name)(ptr(pred), ptr(tgt), C.c_int(n), C.c_int(dim), ptr(loss), ptr(grad))
        return loss, grad
    getattr(lib, name)(ptr(pred), ptr(tgt), C.c_int(n), C.c_int(dim), ptr(loss))
    return loss


def _boxes(rng, n, dim):
    """pairs with every overlap regime: heavy, partial, touching-free disjoint"""
    t = np.concatenate([rng.uniform(0, 4, (n, 3)), rng.uniform(0.3, 1.5, (n, 3))], 1)
    shift = rng.normal(0, 1, (n, 3)) * rng.choice([0.05, 0.3, 1.5], (n, 1))
    p = np.concatenate([t[:, :3] + shift, t[:, 3:] * rng.uniform(0.6, 1.6, (n, 3))], 1)
    if dim == 7:
        t = np.concatenate([t, rng.uniform(-3.1, 3.1, (n, 1))], 1)
        p = np.concatenate([p, t[:, 6:] + rng.normal(0, 0.4, (n, 1))], 1)
    return p, t


def test_aligned_diou_gradient_vs_torch_autograd(lib):
    rng = np.random.default_rng(0)
    p, t = _boxes(rng, 400, 6)
    loss, grad = _call(lib, "bl_pair_loss_grad_f32", p, t, np.float32)
    pt = torch.tensor(p, dtype=torch.float32, requires_grad=True)
    ref = oc.aligned_diou_loss(oc.bbox_to_loss(pt), oc.bbox_to_loss(torch.tensor(t, dtype=torch.float32)))
    ref.sum().backward()
    assert np.allclose(loss, ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    assert np.allclose(grad, pt.grad.numpy(), rtol=1e-4, atol=1e-5)
    assert (ref.detach().numpy() > 1.0).any() and (ref.detach().numpy() < 0.5).any()        # disjoint and overlapping pairs


def test_rotated_diou_value_vs_oracle_and_gradient_vs_finite_differences(lib):
    rng = np.random.default_rng(1)
    p, t = _boxes(rng, 300, 7)
    loss32, grad32 = _call(lib, "bl_pair_loss_grad_f32", p, t, np.float32)
    ref = oc.rotated_diou_loss(torch.tensor(p, dtype=torch.float32), torch.tensor(t, dtype=torch.float32)).numpy()
    assert np.allclose(loss32, ref, rtol=2e-4, atol=2e-5), np.abs(loss32 - ref).max()
    loss64, grad64 = _call(lib, "bl_pair_loss_grad_f64", p, t, np.float64)
    assert np.allclose(loss64, _call(lib, "bl_pair_loss_f64", p, t, np.float64, with_grad=False), rtol=1e-12, atol=1e-12)
    h = 1e-6
    fd = np.zeros_like(grad64)
    for k in range(7):
        e = np.zeros(7)
        e[k] = h
        fd[:, k] = (_call(lib, "bl_pair_loss_f64", p + e, t, np.float64, with_grad=False) -
                    _call(lib, "bl_pair_loss_f64", p - e, t, np.float64, with_grad=False)) / (2 * h)
    err = np.abs(grad64 - fd).max(1) / np.maximum(np.abs(fd).max(1), 1e-3)
    # a finite difference that straddles a kink of the loss (a vertex entering / leaving the polygon, a min / max switching)
    # is not a derivative: allow a handful of such pairs, everything else must agree to 1e-5
    assert np.quantile(err, 0.97) < 1e-5, np.sort(err)[-12:]
    assert (err < 1e-5).sum() >= 0.97 * len(err)
    assert np.allclose(grad32, grad64, rtol=5e-3, atol=5e-4), np.abs(grad32 - grad64).max()
    assert (np.abs(grad64[:, 6]) > 1e-3).mean() > 0.5            # the yaw gradient is exercised


@pytest.mark.parametrize("with_angle", [False, True])
def test_bbox_decode_backward_vs_torch_autograd(lib, with_angle):
    """PredBBox exp + _bbox_pred_to_bbox (encoder.py:109-111,241-283): value and J^T d_box of the dual-number template
    against torch.autograd through the oracle's restatement."""
    from oracle.encoder import bbox_pred_to_bbox
    g = torch.Generator().manual_seed(3)
    n, dim = 300, 7 if with_angle else 6
    raw = (torch.randn(n, 8, generator=g) * 0.8).requires_grad_(True)
    centers = torch.randn(n, 3, generator=g)
    d_box = torch.randn(n, dim, generator=g)
    pred = torch.cat((torch.exp(raw[:, :6]), raw[:, 6:]), 1) if with_angle else torch.exp(raw[:, :6])
    box = bbox_pred_to_bbox(centers, pred)
    box.backward(d_box)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    r = np.ascontiguousarray(raw.detach().numpy(), np.float32)
    got_box, got = np.zeros((n, dim), np.float32), np.zeros((n, 8), np.float32)
    lib.bl_bbox_decode_f32(ptr(r), ptr(np.ascontiguousarray(centers.numpy())), C.c_int(n), C.c_int(int(with_angle)), ptr(got_box))
    lib.bl_bbox_decode_backward_f32(ptr(r), C.c_int(n), C.c_int(int(with_angle)), ptr(np.ascontiguousarray(d_box.numpy())), ptr(got))
    assert np.allclose(got_box, box.detach().numpy(), rtol=1e-5, atol=1e-6)
    assert np.allclose(got, raw.grad.numpy(), rtol=1e-4, atol=1e-5), np.abs(got - raw.grad.numpy()).max()
    if not with_angle:
        assert not got[:, 6:].any()


def test_rotated_diou_invariances_and_regimes(lib):
    """Size-independent properties of the rotated DIoU templates: translating both boxes changes nothing; a yaw of
    alpha + pi is the same rectangle; swapping (w, h) together with alpha + pi/2 keeps the geometry (only the reference's
    (x, y, w) centre penalty sees the swap, so that is checked with equal w on both sides); disjoint pairs have zero
    intersection gradient but a non-zero centre-penalty gradient; a box inside the other has IoU = volume ratio."""
    rng = np.random.default_rng(7)
    p, t = _boxes(rng, 200, 7)
    base, gbase = _call(lib, "bl_pair_loss_grad_f64", p, t, np.float64)
    shift = np.zeros(7)
    shift[:3] = (3.25, -1.5, 0.75)
    moved, gmoved = _call(lib, "bl_pair_loss_grad_f64", p + shift, t + shift, np.float64)
    assert np.allclose(moved, base, atol=1e-9) and np.allclose(gmoved, gbase, atol=1e-7)
    turn = np.zeros(7)
    turn[6] = np.pi
    flipped, gflipped = _call(lib, "bl_pair_loss_grad_f64", p + turn, t, np.float64)
    assert np.allclose(flipped, base, atol=1e-9) and np.allclose(gflipped, gbase, atol=1e-6)
    # w <-> h with a quarter turn, on pairs with the same w (r2's third term is (w_p - w_t)^2 in the reference)
    p2, t2 = p.copy(), t.copy()
    p2[:, 3] = t2[:, 3] = 0.9
    p2[:, 4] = t2[:, 4] = 0.9            # squares in BEV: the swap is then exactly a quarter turn
    a = _call(lib, "bl_pair_loss_f64", p2, t2, np.float64, with_grad=False)
    q = p2.copy()
    q[:, 6] += np.pi / 2
    assert np.allclose(_call(lib, "bl_pair_loss_f64", q, t2, np.float64, with_grad=False), a, atol=1e-9)
    # disjoint: loss = 1 + r2 / c2 > 1, gradient w.r.t. the yaw comes from the enclosing box only
    far = p.copy()
    far[:, :2] = t[:, :2] + 10.0
    lf, gf = _call(lib, "bl_pair_loss_grad_f64", far, t, np.float64)
    assert (lf > 1.0).all() and (np.abs(gf[:, :2]).max(1) > 0).all()
    # containment: same centre and yaw, every size of p half of t -> IoU = 1/8, no centre penalty except (w_p - w_t)^2 / c2
    inner = t.copy()
    inner[:, 3:6] *= 0.5
    li = _call(lib, "bl_pair_loss_f64", inner, t, np.float64, with_grad=False)
    c2 = t[:, 3] ** 2 + t[:, 4] ** 2 + t[:, 5] ** 2
    # the enclosing box of two concentric, equally rotated rectangles is the larger one's axis-aligned hull, so c2 is at
    # least |size|^2: bound the penalty instead of restating it
    pen = (0.5 * t[:, 3]) ** 2 / c2
    assert (li >= 1 - 0.125 - 1e-9).all() and (li <= 1 - 0.125 + pen + 1e-9).all()
