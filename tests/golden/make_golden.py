"""Generate golden fixtures by EXECUTING THE REFERENCE'S OWN PYTHON in the build container.

Run (build container only -- /root/reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/{encoder_ref,unet_ref,post_ref,criterion_ref,criterion_grad_ref,backward_ref,train_step_ref,predict_ref,collate_ref,gt_prep_ref,augment_ref,evaluate_ref}.npz (all, or the ones named on the command line).  The reference modules are
imported unmodified from /root/reference; their absent third-party imports
(mmengine, mmdet3d, spconv, MinkowskiEngine, torch_scatter, mmcv) are replaced by
minimal stand-ins defined here:

* ``spconv.pytorch`` -> a DENSE stand-in: features are densified and the convs run
  through ``torch.nn.functional.conv3d / conv_transpose3d`` (an implementation that
  shares no code with the oracle's gather/mm/index_add restatement);
* ``mmcv.ops.nms3d*`` / ``aligned_3d_nms`` / ``scatter_mean`` -> the oracle's
  restatements (so post_ref pins the surrounding reference logic only: top-k,
  per-class loop, output order, superpoint trimming, face distances);
* registries / BaseModule / Base3DDetector -> no-op shims.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------- stubs
class _Registry:
    """``@X.register_module()`` records the class, ``X.build(dict(type=..., **kw))`` instantiates it."""
    classes = {}

    def register_module(self, *a, **k):
        def deco(cls):
            _Registry.classes[cls.__name__] = cls
            return cls
        return deco

    def build(self, cfg):
        cfg = dict(cfg)
        return _Registry.classes[cfg.pop("type")](**cfg)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, indice_dict=None):
        self.features, self.indices = features, indices
        self.spatial_shape, self.batch_size = [int(s) for s in spatial_shape], batch_size
        self.indice_dict = {} if indice_dict is None else indice_dict

    def replace_feature(self, f):
        return SparseConvTensor(f, self.indices, self.spatial_shape, self.batch_size, self.indice_dict)

    def dense(self):
        c = self.features.shape[1]
        d = self.features.new_zeros((self.batch_size, c, *self.spatial_shape))
        i = self.indices.long()
        d[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]] = self.features
        return d


class SparseModule(nn.Module):
    pass


class _Conv(SparseModule):
    def __init__(self, cin, cout, kernel_size, stride=1, padding=0, bias=False, indice_key=None):
        super().__init__()
        assert not bias
        k = kernel_size
        self.k, self.stride, self.padding, self.indice_key = k, stride, padding, indice_key
        self.weight = nn.Parameter(torch.randn(cout, k, k, k, cin) * 0.1)   # spconv-2.x layout [C_out,k,k,k,C_in]


class SubMConv3d(_Conv):
    def forward(self, x):
        w = self.weight.permute(0, 4, 1, 2, 3)
        y = F.conv3d(x.dense(), w, padding=self.padding)
        i = x.indices.long()
        return x.replace_feature(y[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]])


class SparseConv3d(_Conv):
    def forward(self, x):
        w = self.weight.permute(0, 4, 1, 2, 3)
        y = F.conv3d(x.dense(), w, stride=self.stride)
        occ = x.features.new_zeros((x.batch_size, 1, *x.spatial_shape))
        i = x.indices.long()
        occ[i[:, 0], 0, i[:, 1], i[:, 2], i[:, 3]] = 1
        occ = F.max_pool3d(occ, self.k, self.stride)[:, 0]
        oi = occ.nonzero()                                           # ascending (b,x,y,z)
        out = SparseConvTensor(y[oi[:, 0], :, oi[:, 1], oi[:, 2], oi[:, 3]], oi.int(),
                               list(y.shape[2:]), x.batch_size, x.indice_dict)
        out.indice_dict[self.indice_key] = (x.indices, x.spatial_shape)
        return out


class SparseInverseConv3d(_Conv):
    def forward(self, x):
        fine_idx, fine_shape = x.indice_dict[self.indice_key]
        w = self.weight.permute(4, 0, 1, 2, 3)                       # conv_transpose3d: [C_in, C_out, k,k,k]
        y = F.conv_transpose3d(x.dense(), w, stride=self.k)
        full = y.new_zeros((x.batch_size, y.shape[1], *fine_shape))
        s = [min(a, b) for a, b in zip(y.shape[2:], fine_shape)]
        full[:, :, :s[0], :s[1], :s[2]] = y[:, :, :s[0], :s[1], :s[2]]
        i = fine_idx.long()
        return SparseConvTensor(full[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]], fine_idx, fine_shape,
                                x.batch_size, x.indice_dict)


class SparseSequential(SparseModule):
    def __init__(self, *args):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], dict):
            for k, v in args[0].items():
                self.add_module(k, v)
        else:
            for i, m in enumerate(args):
                self.add_module(str(i), m)

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            else:
                x = x.replace_feature(m(x.features))
        return x


def install_stubs():
    from oracle import nms as onms
    from oracle.pool import scatter_mean as o_scatter_mean

    reg = _Registry()
    _mod("mmengine"); _mod("mmengine.model", BaseModule=nn.Module)

    class InstanceData:
        def __init__(self, **kw):
            self.__dict__.update(kw)

        def __len__(self):                                   # mmengine: length of the data fields
            return len(next(iter(self.__dict__.values()))) if self.__dict__ else 0
    _mod("mmengine.structures", InstanceData=InstanceData)
    _mod("mmdet3d"); _mod("mmdet3d.registry", MODELS=reg, TASK_UTILS=reg)

    class DepthInstance3DBoxes:
        """stand-in: built with origin (0.5, 0.5, 0.5) everywhere in the reference, so ``tensor[:, :3]`` is kept as
        the gravity centre (mmdet3d stores the bottom centre and converts back in ``gravity_center``)."""
        def __init__(self, tensor, box_dim=7, with_yaw=True, origin=(0.5, 0.5, 0)):
            self.tensor, self.box_dim, self.with_yaw = tensor, box_dim, with_yaw

        @property
        def gravity_center(self):
            return self.tensor[:, :3]

        def __len__(self):
            return len(self.tensor)

        def __getitem__(self, idx):
            return DepthInstance3DBoxes(self.tensor[idx], self.box_dim, self.with_yaw)

    def rotation_3d_in_axis(points, angles, axis=0):
        assert axis in (2, -1)
        s, c = torch.sin(angles), torch.cos(angles)
        o, z = torch.ones_like(c), torch.zeros_like(c)
        rot_t = torch.stack([torch.stack([c, s, z]), torch.stack([-s, c, z]), torch.stack([z, z, o])])
        return torch.einsum("aij,jka->aik", points, rot_t)

    class Base3DDetector(nn.Module):
        pass
    class AxisAlignedBboxOverlaps3D:
        """mmdet3d.structures.ops.iou3d_calculator (1.4.0), the is_aligned branch the reference uses."""
        def __call__(self, b1, b2, mode="iou", is_aligned=False):
            assert is_aligned and mode == "iou"
            a1 = (b1[..., 3] - b1[..., 0]) * (b1[..., 4] - b1[..., 1]) * (b1[..., 5] - b1[..., 2])
            a2 = (b2[..., 3] - b2[..., 0]) * (b2[..., 4] - b2[..., 1]) * (b2[..., 5] - b2[..., 2])
            lt, rb = torch.max(b1[..., :3], b2[..., :3]), torch.min(b1[..., 3:], b2[..., 3:])
            wh = (rb - lt).clamp(min=0)
            overlap = wh[..., 0] * wh[..., 1] * wh[..., 2]
            union = torch.max(a1 + a2 - overlap, a1.new_tensor([1e-6]))
            return overlap / union
    _mod("mmdet3d.structures", DepthInstance3DBoxes=DepthInstance3DBoxes, rotation_3d_in_axis=rotation_3d_in_axis,
         AxisAlignedBboxOverlaps3D=AxisAlignedBboxOverlaps3D)
    _mod("mmdet3d.models", Base3DDetector=Base3DDetector, axis_aligned_iou_loss=None, rotated_iou_3d_loss=None)

    def weighted_loss(fn):                                    # mmdet: reduction 'none', weight None -> identity
        def wrapper(pred, target, weight=None, reduction="mean", avg_factor=None, **kw):
            assert weight is None and reduction == "none" and avg_factor is None
            return fn(pred, target, **kw)
        return wrapper
    _mod("mmdet"); _mod("mmdet.models"); _mod("mmdet.models.losses")
    _mod("mmdet.models.losses.utils", weighted_loss=weighted_loss)

    def box2corners(box):                                     # mmcv.ops.diff_iou_rotated.box2corners
        B = box.size()[0]
        x, y, w, h, alpha = box.split([1, 1, 1, 1, 1], dim=-1)
        x4 = box.new_tensor([0.5, -0.5, -0.5, 0.5]) * w
        y4 = box.new_tensor([0.5, 0.5, -0.5, -0.5]) * h
        corners = torch.stack([x4, y4], dim=-1)
        sin, cos = torch.sin(alpha), torch.cos(alpha)
        rot_T = torch.stack([torch.cat([cos, sin], dim=-1), torch.cat([-sin, cos], dim=-1)], dim=-2)
        rotated = torch.bmm(corners.view([-1, 4, 2]), rot_T.view([-1, 2, 2])).view([B, -1, 4, 2])
        rotated[..., 0] += x
        rotated[..., 1] += y
        return rotated

    def oriented_box_intersection_2d(c1, c2):
        """intersection area of two convex quadrilaterals: float64 Sutherland-Hodgman clipping, independent of the
        oracle's restatement of mmcv's vertex-sorting algorithm."""
        def clip(poly, a, b):
            out = []
            for i in range(len(poly)):
                p, q = poly[i], poly[(i + 1) % len(poly)]
                sp = (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
                sq = (b[0] - a[0]) * (q[1] - a[1]) - (b[1] - a[1]) * (q[0] - a[0])
                if sp >= 0:
                    out.append(p)
                if sp * sq < 0:
                    t = sp / (sp - sq)
                    out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
            return out
        a1, a2 = c1.double().reshape(-1, 4, 2).numpy(), c2.double().reshape(-1, 4, 2).numpy()
        area = np.zeros(len(a1))
        for n in range(len(a1)):
            poly = [tuple(v) for v in a1[n]]
            q = a2[n]
            if (q[1, 0] - q[0, 0]) * (q[2, 1] - q[0, 1]) - (q[1, 1] - q[0, 1]) * (q[2, 0] - q[0, 0]) < 0:
                q = q[::-1]
            for i in range(4):
                poly = clip(poly, q[i], q[(i + 1) % 4])
                if not poly:
                    break
            if len(poly) >= 3:
                xs, ys = np.array([v[0] for v in poly]), np.array([v[1] for v in poly])
                area[n] = 0.5 * abs(np.dot(xs, np.roll(ys, -1)) - np.dot(ys, np.roll(xs, -1)))
        return torch.as_tensor(area, dtype=torch.float32).reshape(c1.shape[:-2]), None
    _mod("mmcv.ops.diff_iou_rotated", box2corners=box2corners, oriented_box_intersection_2d=oriented_box_intersection_2d)
    _mod("mmdet3d.models.layers")

    def aligned_3d_nms(boxes, scores, classes, thr):
        return torch.as_tensor(onms.aligned_3d_nms(boxes.numpy(), scores.numpy(), classes.numpy(), thr))
    _mod("mmdet3d.models.layers.box3d_nms", aligned_3d_nms=aligned_3d_nms)
    _mod("mmcv")
    _mod("mmcv.ops",
         nms3d=lambda b, s, t: torch.as_tensor(onms.nms3d(b.numpy(), s.numpy(), t)),
         nms3d_normal=lambda b, s, t: torch.as_tensor(onms.nms3d_normal(b.numpy(), s.numpy(), t)))

    def scatter_mean(src, index, dim=0):
        if dim in (-1, src.dim() - 1) and src.dim() == 2:
            return o_scatter_mean(src.t().contiguous(), index).t()
        return o_scatter_mean(src, index)
    _mod("torch_scatter", scatter_mean=scatter_mean)

    class BaseTransform:
        def __call__(self, results):
            return self.transform(results)

    class PointSample(BaseTransform):                          # mmdet3d.datasets.transforms.PointSample (ctor only)
        def __init__(self, num_points, sample_range=None, replace=False):
            self.num_points, self.sample_range, self.replace = num_points, sample_range, replace
    _mod("mmcv.transforms", BaseTransform=BaseTransform)
    _mod("mmdet3d.datasets"); _mod("mmdet3d.datasets.transforms", PointSample=PointSample)
    sys.modules["mmdet3d.registry"].TRANSFORMS = reg
    _mod("MinkowskiEngine")
    sp = _mod("spconv")
    spp = _mod("spconv.pytorch", SparseConvTensor=SparseConvTensor, SubMConv3d=SubMConv3d, SparseConv3d=SparseConv3d,
               SparseInverseConv3d=SparseInverseConv3d, SparseSequential=SparseSequential)
    _mod("spconv.pytorch.modules", SparseModule=SparseModule)
    sp.pytorch = spp
    # the reference hard-wires SyncBatchNorm (spconv_unet.py:119-121); eval math == BatchNorm1d
    nn.SyncBatchNorm = nn.BatchNorm1d
    # import the reference modules as a namespace package without running unidet3d/__init__.py
    pkg = types.ModuleType("unidet3d"); pkg.__path__ = [os.path.join(REF, "unidet3d")]
    sys.modules["unidet3d"] = pkg
    # criterion.py needs mmdet3d.registry.TASK_UTILS / MODELS only
    _mod("unidet3d.structures", InstanceData_=InstanceData)


def t2n(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in d.items()}


# ----------------------------------------------------------------------------- fixtures
def gen_encoder():
    from unidet3d.encoder import UniDet3DEncoder
    torch.manual_seed(7)
    classes = [["chair", "table", "sofa"], ["table", "board"], ["bed", "chair", "oven", "sink"]]
    cfg = dict(num_layers=2, datasets_classes=classes, in_channels=8, d_model=64, num_heads=2, hidden_dim=128,
               dropout=0.0, activation_fn="gelu", datasets=["scannet", "s3dis", "arkitscenes"],
               angles=[False, False, True])
    m = UniDet3DEncoder(**cfg).eval()
    with torch.no_grad():
        for p in m.parameters():            # non-trivial LN / bias values
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    T = [37, 21, 50]
    x = [torch.randn(t, 8) for t in T]
    c = [torch.randn(t, 3) for t in T]
    names = ["scannet", "arkitscenes", "s3dis"]
    with torch.no_grad():
        out = m(x, c, names)
    save = {"sd." + k: v for k, v in t2n(m.state_dict()).items()}
    for i in range(3):
        save[f"x{i}"], save[f"c{i}"] = x[i].numpy(), c[i].numpy()
        save[f"cls{i}"], save[f"box{i}"] = out["cls_preds"][i].numpy(), out["bboxes"][i].numpy()
        for l, aux in enumerate(out["aux_outputs"]):
            save[f"aux{l}_cls{i}"], save[f"aux{l}_box{i}"] = aux["cls_preds"][i].numpy(), aux["bboxes"][i].numpy()
    save["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "encoder_ref.npz"), **save)
    print("encoder_ref.npz", {k: v.shape for k, v in save.items() if k.startswith("cls")})


def gen_unet():
    from unidet3d.spconv_unet import SpConvUNet
    torch.manual_seed(11)
    planes = [8, 16, 24, 32, 40]
    m = SpConvUNet(planes, return_blocks=True).eval()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.BatchNorm1d):
                mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.1)
                mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
    shape = [40, 36, 27]
    rng = np.random.default_rng(3)
    B = 2
    coords = []
    for b in range(B):
        # points near two planes + random clutter -> neighbourhood statistics like surfaces
        n = 900
        xy = rng.integers(0, [shape[0], shape[1]], (n, 2))
        z = np.clip((0.3 * xy[:, 0] + rng.integers(0, 3, n)).astype(np.int64), 0, shape[2] - 1)
        c1 = np.concatenate([xy, z[:, None]], 1)
        c2 = rng.integers(0, shape, (300, 3))
        c2[:50] = np.array(shape) - 1 - rng.integers(0, 2, (50, 3))      # touch the odd upper boundary
        cc = np.unique(np.concatenate([c1, c2]), axis=0)
        cc = cc[rng.permutation(len(cc))]                                # arbitrary (non-canonical) row order
        coords.append(np.concatenate([np.full((len(cc), 1), b), cc], 1))
    coords = torch.as_tensor(np.concatenate(coords), dtype=torch.int32)
    feats = torch.randn(len(coords), planes[0])
    x = SparseConvTensor(feats, coords, shape, B)
    with torch.no_grad():
        y, blocks = m(x)
    assert torch.equal(y.indices, coords)
    save = {"sd." + k: v for k, v in t2n(m.state_dict()).items()}
    save.update(coords=coords.numpy(), feats=feats.numpy(), shape=np.array(shape), out=y.features.numpy())
    np.savez_compressed(os.path.join(HERE, "unet_ref.npz"), **save)
    print("unet_ref.npz", coords.shape, y.features.shape, float(y.features.abs().mean()))


def gen_criterion():
    """The reference's own criterion.py / axis_aligned_iou_loss.py / rotated_iou_loss.py and the GT-target helpers of
    unidet3d.py on a 3-scene batch (axis-aligned x2 incl. a scene without GT, rotated x1), final layer + 2 aux."""
    import unidet3d.axis_aligned_iou_loss  # noqa: F401  (registers UniDet3DAxisAlignedIoULoss)
    import unidet3d.rotated_iou_loss  # noqa: F401
    from unidet3d.criterion import UniDet3DCriterion
    from unidet3d.unidet3d import UniDet3D
    InstanceData = sys.modules["mmengine.structures"].InstanceData
    Boxes = sys.modules["mmdet3d.structures"].DepthInstance3DBoxes
    torch.manual_seed(21)
    rng = np.random.default_rng(21)
    datasets = ["scannet", "s3dis", "arkitscenes"]
    diou = lambda t: dict(type=t, mode="diou", reduction="none")
    simple, rotated = diou("UniDet3DAxisAlignedIoULoss"), diou("UniDet3DRotatedIoU3DLoss")
    crit = UniDet3DCriterion(
        matcher=dict(type="UniMatcher", costs=[dict(type="QueryClassificationCost", weight=0.5),
                                               dict(type="BboxCostJointTraining", weight=2.0, loss_simple=simple,
                                                    loss_rotated=rotated)]),
        loss_weight=[0.5, 1.0], non_object_weight=0.1, iter_matcher=True, bbox_loss_simple=simple,
        bbox_loss_rotated=rotated, datasets=datasets, datasets_weights=[1.0, 0.7, 1.3], topk=[6, 4, 5])
    names = ["scannet", "arkitscenes", "s3dis", "scannet"]
    T = [90, 80, 70, 40]
    G = [7, 5, 6, 0]
    C = {"scannet": 5, "s3dis": 4, "arkitscenes": 6}
    save = {"names": np.array(names)}

    def rand_boxes(n, dim):
        c = rng.uniform(0.5, 3.5, (n, 3))
        sz = rng.uniform(0.3, 1.6, (n, 3))
        b = np.concatenate([c, sz] + ([rng.uniform(-3, 3, (n, 1))] if dim == 7 else []), 1)
        return torch.as_tensor(b.astype(np.float32))
    insts = []
    for i, nm in enumerate(names):
        dim = 7 if nm == "arkitscenes" else 6
        gt = rand_boxes(G[i], dim)
        labels = torch.as_tensor(rng.integers(0, C[nm], G[i]))
        qm = torch.as_tensor(rng.random((G[i], T[i])) < 0.35)
        if G[i]:
            qm[0, :] = False
            qm[0, :3] = True                                  # a GT with fewer than topk+1 candidate queries
        insts.append(InstanceData(labels_3d=labels, bboxes_3d=Boxes(gt, box_dim=dim, with_yaw=dim == 7,
                                                                    origin=(0.5, 0.5, 0.5)), query_masks=qm))
        save.update({f"gt_boxes{i}": gt.numpy(), f"gt_labels{i}": labels.numpy(), f"qmask{i}": qm.numpy()})

    def layer():
        cls, box = [], []
        for i, nm in enumerate(names):
            dim = 7 if nm == "arkitscenes" else 6
            cls.append(torch.randn(T[i], C[nm] + 1) * 1.5)
            b = rand_boxes(T[i], dim)
            k = min(G[i], T[i])
            if k:                                             # some predictions close to a GT
                b[:k] = insts[i].bboxes_3d.tensor + 0.05 * torch.randn(k, dim)
            box.append(b)
        return dict(cls_preds=cls, bboxes=box)
    pred = layer()
    pred["aux_outputs"] = [layer(), layer()]
    out = crit(pred, insts, names)
    save["det_loss"] = out["det_loss"].numpy()
    for l, lay in enumerate([pred] + pred["aux_outputs"]):
        save[f"layer_loss{l}"] = crit.get_layer_loss(lay, insts, names).numpy()
        for i in range(len(names)):
            save[f"l{l}_cls{i}"], save[f"l{l}_box{i}"] = lay["cls_preds"][i].numpy(), lay["bboxes"][i].numpy()
            if G[i]:
                pi = InstanceData(scores=lay["cls_preds"][i], bboxes=lay["bboxes"][i])
                gi = InstanceData(labels=insts[i].labels_3d, query_masks=insts[i].query_masks, bboxes=insts[i].bboxes_3d.tensor)
                iq, ig = crit.matcher(pi, gi, crit.topk[datasets.index(names[i])])
                save[f"l{l}_iq{i}"], save[f"l{l}_ig{i}"] = iq.numpy(), ig.numpy()
    # GT-target helpers of the detector (unidet3d.py:220-275, 371-409)
    det = object.__new__(UniDet3D)
    pts = torch.as_tensor(rng.uniform(0, 4, (3000, 3)).astype(np.float32))
    inst = torch.as_tensor(rng.integers(-1, 9, 3000))
    masks = UniDet3D.get_gt_inst_masks(det, inst)
    bb = UniDet3D.get_bboxes_by_masks(det, masks.T, pts)
    save.update(bm_points=pts.numpy(), bm_inst=inst.numpy(), bm_boxes=bb.tensor.numpy())
    spc = torch.as_tensor(rng.uniform(0, 4, (400, 3)).astype(np.float32))
    gtb = Boxes(rand_boxes(9, 7), box_dim=7, with_yaw=True, origin=(0.5, 0.5, 0.5))
    tg = UniDet3D.get_targets(det, spc, gtb, 6)
    save.update(tg_centers=spc.numpy(), tg_boxes=gtb.tensor.numpy(), tg_masks=tg.numpy())
    # edge cases on one axis-aligned scene each: a single GT; T == topk + 1; duplicated predictions (exactly tied costs:
    # `cost < kth` then matches fewer than topk queries); all queries masked out for one GT
    edge = []
    for tag, T_e, G_e, dup, mask_all in [("one_gt", 30, 1, False, False), ("t_eq_k1", 7, 3, False, False),
                                         ("ties", 24, 3, True, False), ("masked", 20, 2, False, True)]:
        gt = rand_boxes(G_e, 6)
        labels = torch.as_tensor(rng.integers(0, 5, G_e))
        cls = torch.randn(T_e, 6) * 1.5
        box = rand_boxes(T_e, 6)
        box[:G_e] = gt + 0.03 * torch.randn(G_e, 6)
        if dup:
            cls[8:16] = cls[0:8]
            box[8:16] = box[0:8]
        qm = torch.ones(G_e, T_e, dtype=torch.bool)
        if mask_all:
            qm[1] = False
        inst = InstanceData(labels_3d=labels, bboxes_3d=Boxes(gt, box_dim=6, with_yaw=False, origin=(0.5, 0.5, 0.5)), query_masks=qm)
        lay = dict(cls_preds=[cls], bboxes=[box])
        loss = crit.get_layer_loss(lay, [inst], ["scannet"])
        iq, ig = crit.matcher(InstanceData(scores=cls, bboxes=box), InstanceData(labels=labels, query_masks=qm, bboxes=gt), 6)
        save.update({f"e_{tag}_cls": cls.numpy(), f"e_{tag}_box": box.numpy(), f"e_{tag}_gt": gt.numpy(),
                     f"e_{tag}_labels": labels.numpy(), f"e_{tag}_qm": qm.numpy(), f"e_{tag}_loss": loss.numpy(),
                     f"e_{tag}_iq": iq.numpy(), f"e_{tag}_ig": ig.numpy()})
        edge.append((tag, len(iq), float(loss)))
    print("criterion edge cases", edge)
    np.savez_compressed(os.path.join(HERE, "criterion_ref.npz"), **save)
    print("criterion_ref.npz det_loss", float(out["det_loss"]), "matches per layer/scene",
          [[len(save.get(f"l{l}_iq{i}", [])) for i in range(len(names))] for l in range(3)])


def gen_criterion_grad():
    """GRADIENTS of the reference's own criterion.py (with axis_aligned_iou_loss.py) under torch.autograd: det_loss of a
    3-scene axis-aligned batch (two datasets with different weights, a scene without GT, a GT with every query masked
    out), final layer + 2 aux layers, differentiated w.r.t. every layer's class logits and boxes.  (Rotated boxes are
    left out: the reference differentiates mmcv's CUDA-backed ``oriented_box_intersection_2d``, which is not installable
    here -- the rotated DIoU derivative is checked against finite differences instead, tests/test_box_loss_host.py.)"""
    import unidet3d.axis_aligned_iou_loss  # noqa: F401
    import unidet3d.rotated_iou_loss  # noqa: F401
    from unidet3d.criterion import UniDet3DCriterion
    InstanceData = sys.modules["mmengine.structures"].InstanceData
    Boxes = sys.modules["mmdet3d.structures"].DepthInstance3DBoxes
    torch.manual_seed(33)
    rng = np.random.default_rng(33)
    datasets = ["scannet", "s3dis"]
    diou = lambda t: dict(type=t, mode="diou", reduction="none")
    simple, rotated = diou("UniDet3DAxisAlignedIoULoss"), diou("UniDet3DRotatedIoU3DLoss")
    crit = UniDet3DCriterion(
        matcher=dict(type="UniMatcher", costs=[dict(type="QueryClassificationCost", weight=0.5),
                                               dict(type="BboxCostJointTraining", weight=2.0, loss_simple=simple,
                                                    loss_rotated=rotated)]),
        loss_weight=[0.5, 1.0], non_object_weight=0.1, iter_matcher=True, bbox_loss_simple=simple,
        bbox_loss_rotated=rotated, datasets=datasets, datasets_weights=[1.0, 0.7], topk=[4, 3])
    names = ["scannet", "s3dis", "scannet", "s3dis"]
    T = [80, 60, 50, 45]
    G = [6, 5, 0, 3]
    C = {"scannet": 5, "s3dis": 4}
    save = {"names": np.array(names), "datasets": np.array(datasets), "datasets_weights": np.array([1.0, 0.7]),
            "topk": np.array([4, 3])}

    def rand_boxes(n):
        return torch.as_tensor(np.concatenate([rng.uniform(0.5, 3.5, (n, 3)), rng.uniform(0.3, 1.6, (n, 3))], 1).astype(np.float32))
    insts = []
    for i, nm in enumerate(names):
        gt = rand_boxes(G[i])
        labels = torch.as_tensor(rng.integers(0, C[nm], G[i]))
        qm = torch.as_tensor(rng.random((G[i], T[i])) < 0.5)
        if i == 3:
            qm[1, :] = False                                   # a GT nobody may match: no pair from it
        insts.append(InstanceData(labels_3d=labels, bboxes_3d=Boxes(gt, box_dim=6, with_yaw=False, origin=(0.5, 0.5, 0.5)),
                                  query_masks=qm))
        save.update({f"gt_boxes{i}": gt.numpy(), f"gt_labels{i}": labels.numpy(), f"qmask{i}": qm.numpy()})

    def layer():
        cls, box = [], []
        for i, nm in enumerate(names):
            cls.append((torch.randn(T[i], C[nm] + 1) * 1.5).requires_grad_(True))
            b = rand_boxes(T[i])
            if G[i]:
                b[:G[i]] = insts[i].bboxes_3d.tensor + 0.05 * torch.randn(G[i], 6)
                b[G[i]:2 * G[i]] = insts[i].bboxes_3d.tensor + 0.2 * torch.randn(G[i], 6)
            box.append(b.requires_grad_(True))
        return dict(cls_preds=cls, bboxes=box)
    layers = [layer(), layer(), layer()]
    pred = dict(layers[0], aux_outputs=layers[1:])
    out = crit(pred, insts, names)
    out["det_loss"].backward()
    save["det_loss"] = out["det_loss"].detach().numpy()
    n_pairs = []
    for l, lay in enumerate(layers):
        for i in range(len(names)):
            save[f"l{l}_cls{i}"], save[f"l{l}_box{i}"] = lay["cls_preds"][i].detach().numpy(), lay["bboxes"][i].detach().numpy()
            save[f"l{l}_dcls{i}"] = lay["cls_preds"][i].grad.numpy()
            g = lay["bboxes"][i].grad
            save[f"l{l}_dbox{i}"] = (g if g is not None else torch.zeros_like(lay["bboxes"][i])).numpy()
            if G[i]:
                pi = InstanceData(scores=lay["cls_preds"][i].detach(), bboxes=lay["bboxes"][i].detach())
                gi = InstanceData(labels=insts[i].labels_3d, query_masks=insts[i].query_masks, bboxes=insts[i].bboxes_3d.tensor)
                iq, ig = crit.matcher(pi, gi, crit.topk[datasets.index(names[i])])
                save[f"l{l}_iq{i}"], save[f"l{l}_ig{i}"] = iq.numpy(), ig.numpy()
                n_pairs.append(len(iq))
    np.savez_compressed(os.path.join(HERE, "criterion_grad_ref.npz"), **save)
    print("criterion_grad_ref.npz det_loss", float(out["det_loss"]), "pairs", n_pairs,
          "max |dcls|", max(float(np.abs(save[k]).max()) for k in save if "_dcls" in k),
          "max |dbox|", max(float(np.abs(save[k]).max()) for k in save if "_dbox" in k))


def gen_backward():
    """PARAMETER GRADIENTS of the reference's own modules under torch.autograd, in train mode:
    * UniDet3DEncoder (encoder.py; all num_layers + 1 heads are evaluated in train mode): loss = sum of randomly weighted
      class logits and boxes of every head -> gradient of every parameter and of the input features;
    * SpConvUNet (spconv_unet.py over the dense spconv stand-in; BatchNorm with BATCH statistics): loss = sum of randomly
      weighted output features -> gradient of every conv weight / BatchNorm affine and of the input features.
    The oracle under autograd must reproduce them (tests/test_oracle_golden.py): the backward passes of the library are
    compared with autograd through that oracle."""
    from unidet3d.encoder import UniDet3DEncoder
    from unidet3d.spconv_unet import SpConvUNet
    torch.manual_seed(17)
    classes = [["chair", "table", "sofa"], ["table", "board"], ["bed", "chair", "oven", "sink"]]
    cfg = dict(num_layers=2, datasets_classes=classes, in_channels=8, d_model=64, num_heads=2, hidden_dim=128,
               dropout=0.0, activation_fn="gelu", datasets=["scannet", "s3dis", "arkitscenes"], angles=[False, False, True])
    m = UniDet3DEncoder(**cfg).train()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    T = [33, 19, 41]
    names = ["scannet", "arkitscenes", "s3dis"]
    x = [torch.randn(t, 8).requires_grad_(True) for t in T]
    c = [torch.randn(t, 3) for t in T]
    out = m(x, c, names)
    heads = out["aux_outputs"] + [dict(cls_preds=out["cls_preds"], bboxes=out["bboxes"])]
    assert len(heads) == 3
    save = {"enc_sd." + k: v for k, v in t2n(m.state_dict()).items()}
    loss = 0.0
    for h, hd in enumerate(heads):
        for i in range(3):
            gc, gb = torch.randn_like(hd["cls_preds"][i]), 0.1 * torch.randn_like(hd["bboxes"][i])
            save[f"enc_gc{h}_{i}"], save[f"enc_gb{h}_{i}"] = gc.numpy(), gb.numpy()
            loss = loss + (hd["cls_preds"][i] * gc).sum() + (hd["bboxes"][i] * gb).sum()
    loss.backward()
    for i in range(3):
        save[f"enc_x{i}"], save[f"enc_c{i}"], save[f"enc_dx{i}"] = x[i].detach().numpy(), c[i].numpy(), x[i].grad.numpy()
    for k, p in m.named_parameters():
        save["enc_grad." + k] = p.grad.numpy()
    save["enc_names"] = np.array(names)

    torch.manual_seed(19)
    planes = [8, 16, 24]
    u = SpConvUNet(planes, return_blocks=True).train()
    with torch.no_grad():
        for mod in u.modules():
            if isinstance(mod, nn.BatchNorm1d):
                mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.1)
    shape = [20, 18, 13]
    rng = np.random.default_rng(5)
    coords = []
    for b in range(2):
        n = 500
        xy = rng.integers(0, [shape[0], shape[1]], (n, 2))
        z = np.clip((0.3 * xy[:, 0] + rng.integers(0, 3, n)).astype(np.int64), 0, shape[2] - 1)
        cc = np.unique(np.concatenate([np.concatenate([xy, z[:, None]], 1), rng.integers(0, shape, (150, 3))]), axis=0)
        cc = cc[rng.permutation(len(cc))]
        coords.append(np.concatenate([np.full((len(cc), 1), b), cc], 1))
    coords = torch.as_tensor(np.concatenate(coords), dtype=torch.int32)
    feats = torch.randn(len(coords), planes[0]).requires_grad_(True)
    sd0 = {k: v.clone() for k, v in u.state_dict().items()}           # before the forward updates the running statistics
    y, _ = u(SparseConvTensor(feats, coords, shape, 2))
    R = torch.randn_like(y.features)
    (y.features * R).sum().backward()
    save.update({"unet_sd." + k: v.numpy() for k, v in sd0.items()})
    save.update(unet_coords=coords.numpy(), unet_feats=feats.detach().numpy(), unet_shape=np.array(shape), unet_R=R.numpy(),
                unet_out=y.features.detach().numpy(), unet_dfeats=feats.grad.numpy())
    for k, p in u.named_parameters():
        save["unet_grad." + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "backward_ref.npz"), **save)
    print("backward_ref.npz: encoder", sum(1 for k in save if k.startswith("enc_grad.")), "parameter gradients; unet",
          sum(1 for k in save if k.startswith("unet_grad.")), "parameter gradients,", len(coords), "voxels")


def _install_me_standin():
    """MinkowskiEngine stand-in for UniDet3D.collate (unidet3d.py:136-176): ``batch_sparse_collate`` floors the coordinates and
    prepends the batch index, ``TensorField.sparse()`` keeps one row per distinct coordinate with the UNWEIGHTED AVERAGE of its
    points' features (ME's default quantisation mode), ``inverse_mapping`` maps every point to its voxel row.  Voxel rows come
    out in ascending (b, x, y, z) order (ME's own order is hash-defined).  numpy ``unique``: shares no code with the oracle."""
    class _Sparse:
        def __init__(self, coordinates, features):
            self.coordinates, self.features, self.coordinate_map_key = coordinates, features, object()

    class TensorField:
        def __init__(self, features, coordinates):
            self.features, self.coordinates = features, coordinates

        def sparse(self):
            uniq, inv = np.unique(self.coordinates.long().numpy(), axis=0, return_inverse=True)
            inv = torch.as_tensor(inv.reshape(-1)).long()
            cnt = torch.zeros(len(uniq)).index_add_(0, inv, torch.ones(len(inv)))
            feats = torch.zeros(len(uniq), self.features.shape[1]).index_add_(0, inv, self.features) / cnt[:, None]
            self._inv = inv
            return _Sparse(torch.as_tensor(uniq).int(), feats)

        def inverse_mapping(self, key):
            return self._inv

    def batch_sparse_collate(data):
        cs, fs = [], []
        for b, (c, f) in enumerate(data):
            cs.append(torch.cat((torch.full((len(c), 1), b, dtype=torch.int32), torch.floor(c).int()), 1))
            fs.append(f)
        return torch.cat(cs), torch.cat(fs)
    me = sys.modules["MinkowskiEngine"]
    me.TensorField = TensorField
    me.utils = types.SimpleNamespace(batch_sparse_collate=batch_sparse_collate)


def gen_train_step():
    """The reference's own ``UniDet3D.loss`` (unidet3d.py:277-364) END TO END in train mode under torch.autograd: GT boxes by
    instance masks (scannet scene) / shifted GT boxes + distance targets (3rscan scene), collate (ME stand-in), SpConvUNet over
    the dense spconv stand-in with batch-statistics BatchNorm, superpoint pooling, the encoder with all heads, the criterion
    -> det_loss and the gradient of EVERY parameter of the detector.  The oracle pipeline under autograd must reproduce them
    (tests/test_oracle_golden.py), which pins the composition the GPU training step is compared with."""
    import unidet3d.axis_aligned_iou_loss  # noqa: F401
    import unidet3d.rotated_iou_loss  # noqa: F401
    import unidet3d.criterion  # noqa: F401
    import unidet3d.encoder  # noqa: F401
    import unidet3d.spconv_unet  # noqa: F401
    from unidet3d.unidet3d import UniDet3D
    from unidet3d_b200.synthetic import make_scene, SCENE_PRESETS
    _install_me_standin()
    InstanceData = sys.modules["mmengine.structures"].InstanceData
    Boxes = sys.modules["mmdet3d.structures"].DepthInstance3DBoxes
    MODELS = sys.modules["mmdet3d.registry"].MODELS
    torch.manual_seed(41)
    rng = np.random.default_rng(41)
    datasets = ["scannet", "3rscan"]
    classes = [["chair", "table", "sofa", "bed", "sink"], ["table", "board", "bed", "oven"]]
    planes, voxel = [8, 16, 24], 0.05
    diou = lambda t: dict(type=t, mode="diou", reduction="none")
    simple, rotated = diou("UniDet3DAxisAlignedIoULoss"), diou("UniDet3DRotatedIoU3DLoss")
    cfg = dict(
        backbone=dict(type="SpConvUNet", num_planes=planes, return_blocks=True),
        decoder=dict(type="UniDet3DEncoder", num_layers=2, datasets_classes=classes, in_channels=planes[0], d_model=64, num_heads=2,
                     hidden_dim=128, dropout=0.0, activation_fn="gelu", datasets=datasets, angles=[False, False]),
        criterion=dict(type="UniDet3DCriterion", datasets=datasets, datasets_weights=[1.0, 0.7], bbox_loss_simple=simple,
                       bbox_loss_rotated=rotated,
                       matcher=dict(type="UniMatcher", costs=[dict(type="QueryClassificationCost", weight=0.5),
                                                              dict(type="BboxCostJointTraining", weight=2.0, loss_simple=simple,
                                                                   loss_rotated=rotated)]),
                       loss_weight=[0.5, 1.0], non_object_weight=0.1, topk=[3, 2], iter_matcher=True))
    det = object.__new__(UniDet3D)                       # the constructor's body (unidet3d.py:76-94) without mmengine's BaseModel
    nn.Module.__init__(det)
    det.unet, det.decoder, det.criterion = MODELS.build(cfg["backbone"]), MODELS.build(cfg["decoder"]), MODELS.build(cfg["criterion"])
    det.voxel_size, det.min_spatial_shape, det.query_thr = voxel, 32, 3000
    det.use_superpoints, det.bbox_by_mask, det.target_by_distance, det.fast_nms = [True, False], [True, False], [False, True], [True, True]
    det.train_cfg, det.test_cfg, det.use_sync_bn = types.SimpleNamespace(topk=4), None, True
    det._init_layers(6, planes[0])
    with torch.no_grad():
        for mod in det.modules():
            if isinstance(mod, nn.BatchNorm1d):
                mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.1)
        for k, p in det.decoder.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    det.train()
    n, _, area, cell = SCENE_PRESETS["tiny"]
    scenes = [make_scene(300 + i, n, area, cell) for i in range(2)]
    names = ["scannet", "3rscan"]
    save = {"names": np.array(names), "voxel_size": np.array(voxel), "planes": np.array(planes)}
    samples, pts_in = [], []
    for i, (pts, sp) in enumerate(scenes):
        P, S = torch.as_tensor(pts), torch.as_tensor(sp)
        n_sp = int(sp.max()) + 1
        save[f"points{i}"], save[f"sp{i}"] = pts, sp
        if names[i] == "scannet":
            G = 5
            sp_inst = rng.integers(-1, G, n_sp)
            sp_inst[:G] = np.arange(G)
            inst = sp_inst[sp]
            labels = rng.integers(0, 5, G)
            sp_masks = sp_inst[None, :] == np.arange(G)[:, None]
            gi = InstanceData(labels_3d=torch.as_tensor(labels), sp_masks=torch.as_tensor(sp_masks))
            seg = types.SimpleNamespace(sp_pts_mask=S.clone(), pts_instance_mask=torch.as_tensor(inst))
            save.update({f"inst{i}": inst, f"labels{i}": labels, f"sp_masks{i}": sp_masks})
        else:
            G = 4
            lo, hi = pts[:, :3].min(0), pts[:, :3].max(0)
            gt = np.concatenate([rng.uniform(lo, hi, (G, 3)), rng.uniform(0.3, 1.0, (G, 3))], 1).astype(np.float32)
            labels = rng.integers(0, 4, G)
            gi = InstanceData(labels_3d=torch.as_tensor(labels), bboxes_3d=Boxes(torch.as_tensor(gt), box_dim=6, with_yaw=False,
                                                                                origin=(0.5, 0.5, 0.5)))
            seg = types.SimpleNamespace(sp_pts_mask=S.clone())
            save.update({f"gt_boxes{i}": gt, f"labels{i}": labels})
        samples.append(types.SimpleNamespace(lidar_path=f"data/{names[i]}/points/x.bin", gt_pts_seg=seg, gt_instances_3d=gi))
        pts_in.append(P)
    sd0 = {k: v.clone() for k, v in det.state_dict().items()}
    out = det.loss(dict(points=pts_in), samples)
    out["det_loss"].backward()
    save["det_loss"] = out["det_loss"].detach().numpy()
    save.update({"sd." + k: v.numpy() for k, v in sd0.items()})
    n_grad = 0
    for k, p in det.named_parameters():
        assert p.grad is not None, k
        save["grad." + k] = p.grad.numpy()
        n_grad += 1
    for i, smp in enumerate(samples):                    # what loss() derived on the way (unidet3d.py:306-336)
        gi = smp.gt_instances_3d
        save[f"used_boxes{i}"] = gi.bboxes_3d.tensor.numpy()
        save[f"used_sp_masks{i}"] = gi.sp_masks.numpy()
    np.savez_compressed(os.path.join(HERE, "train_step_ref.npz"), **save)
    print("train_step_ref.npz det_loss", float(save["det_loss"]), "parameter gradients", n_grad,
          "max |grad|", max(float(np.abs(save[k]).max()) for k in save if k.startswith("grad.")))


def gen_predict():
    """The reference's own ``UniDet3D.predict`` (unidet3d.py:411-473) END TO END in eval mode, one call per scene (the
    reference post-processes scene 0 of a batch only, :498-502): collate (ME stand-in), input conv + SpConvUNet over the
    dense spconv stand-in, output BN + ReLU, superpoint pooling, the encoder, ``predict_by_feat`` (top-k, multi-class NMS,
    superpoint trimming).  Three flavours: scannet (fast NMS + superpoint trim), s3dis (3-D aligned NMS + trim), 3rscan
    (fast NMS, no superpoints: [n, 7] boxes with yaw 0).  Pins the COMPOSITION oracle/detector.py::forward_scenes restates
    (the NMS kernels themselves are the oracle's restatements on both sides, see the module docstring)."""
    import unidet3d.encoder  # noqa: F401
    import unidet3d.spconv_unet  # noqa: F401
    from unidet3d.unidet3d import UniDet3D
    from unidet3d_b200.synthetic import make_scene, SCENE_PRESETS
    _install_me_standin()
    MODELS = sys.modules["mmdet3d.registry"].MODELS
    InstanceData = sys.modules["mmengine.structures"].InstanceData
    sys.modules["unidet3d.structures"].InstanceData_ = InstanceData
    torch.manual_seed(53)
    datasets = ["scannet", "s3dis", "3rscan"]
    classes = [["chair", "table", "sofa", "bed", "sink"], ["table", "board", "bed", "oven"], ["chair", "sofa", "lamp"]]
    planes, voxel = [8, 16, 24], 0.05
    det = object.__new__(UniDet3D)
    nn.Module.__init__(det)
    det.unet = MODELS.build(dict(type="SpConvUNet", num_planes=planes, return_blocks=True))
    det.decoder = MODELS.build(dict(type="UniDet3DEncoder", num_layers=2, datasets_classes=classes, in_channels=planes[0], d_model=64,
                                    num_heads=2, hidden_dim=128, dropout=0.0, activation_fn="gelu", datasets=datasets,
                                    angles=[False, False, False]))
    det.voxel_size, det.min_spatial_shape, det.query_thr = voxel, 32, 3000
    det.use_superpoints, det.fast_nms = [True, True, False], [True, False, True]
    det.test_cfg = types.SimpleNamespace(topk_insts=120, score_thr=0.0, iou_thr=[0.5, 0.55, 0.55], low_sp_thr=0.18, up_sp_thr=0.81)
    det.use_sync_bn = True
    det._init_layers(6, planes[0])
    with torch.no_grad():
        for mod in det.modules():
            if isinstance(mod, nn.BatchNorm1d):
                mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.1)
                mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
        for k, p in det.decoder.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
        det.decoder.out_bboxes.linear.bias.add_(-1.0)                       # box sizes of a few decimetres: NMS has work to do
    det.eval()
    n, _, area, cell = SCENE_PRESETS["tiny"]
    save = {"names": np.array(datasets), "voxel_size": np.array(voxel), "planes": np.array(planes)}
    save.update({"sd." + k: v.numpy() for k, v in det.state_dict().items()})
    for i, name in enumerate(datasets):
        pts, sp = make_scene(400 + i, n, area, cell)
        sample = types.SimpleNamespace(lidar_path=f"data/{name}/points/x.bin",
                                       gt_pts_seg=types.SimpleNamespace(sp_pts_mask=torch.as_tensor(sp).clone()))
        with torch.no_grad():
            res = det.predict(dict(points=[torch.as_tensor(pts)]), [sample])
        pi = res[0].pred_instances_3d
        save.update({f"points{i}": pts, f"sp{i}": sp, f"boxes{i}": pi.bboxes_3d.tensor.numpy(), f"labels{i}": pi.labels_3d.numpy(),
                     f"scores{i}": pi.scores_3d.numpy()})
        print("predict", name, pi.bboxes_3d.tensor.shape, float(pi.scores_3d.max()))
    np.savez_compressed(os.path.join(HERE, "predict_ref.npz"), **save)


def gen_collate():
    """The reference's own ``UniDet3D.collate`` (unidet3d.py:136-176) over the numpy MinkowskiEngine stand-in, plain and with
    ``elastic_points`` (float64 where ElasticTransfrom applied, float32 where its coin flip skipped it): the coordinate
    recipe (per-scene min shift, division by the voxel size, floor), the feature recipe (colour | xyz - per-scene mean), the
    spatial shape clip and the inverse mapping."""
    from unidet3d.unidet3d import UniDet3D
    _install_me_standin()
    rng = np.random.default_rng(61)
    det = object.__new__(UniDet3D)
    nn.Module.__init__(det)
    det.voxel_size, det.min_spatial_shape = 0.05, 16
    pts = [rng.uniform(-1.5, 1.7, (1500, 6)).astype(np.float32), rng.uniform(0.2, 2.4, (900, 6)).astype(np.float32)]
    pts[0][:40, :3] = pts[0][40:80, :3]                                    # exact duplicates share a voxel
    el = [(pts[0][:, :3].astype(np.float64) / 0.05 + rng.normal(0, 0.8, (1500, 3))), (pts[1][:, :3] / np.float32(0.05))]
    assert el[0].dtype == np.float64 and el[1].dtype == np.float32
    save = {"points0": pts[0], "points1": pts[1], "elastic0": el[0], "elastic1": el[1], "voxel_size": np.array(0.05)}
    for tag, elastic in (("plain", None), ("elastic", [torch.as_tensor(e) for e in el])):
        c, f, inv, shape = det.collate([torch.as_tensor(p) for p in pts], elastic)
        save.update({f"{tag}_coords": c.numpy(), f"{tag}_feats": f.numpy(), f"{tag}_inverse": inv.numpy(),
                     f"{tag}_shape": shape.numpy()})
        print("collate", tag, tuple(c.shape), shape.tolist())
    np.savez_compressed(os.path.join(HERE, "collate_ref.npz"), **save)


def gen_gt_prep():
    """The reference's GT-preparation transforms (unidet3d/transforms_3d.py) on synthetic masks."""
    from unidet3d.transforms_3d import PointDetClassMappingScanNet, PointDetClassMappingS3DIS, PointSample_
    rng = np.random.default_rng(33)
    n = 5000
    save = {}
    # superpoints: 120 ids; instances: unions of superpoints with 8 % label noise (so the > 0.5 majority matters)
    sp = rng.integers(0, 120, n)
    sp_inst = rng.integers(0, 14, 120)
    inst = sp_inst[sp].copy()
    noise = rng.random(n) < 0.08
    inst[noise] = rng.integers(0, 14, int(noise.sum()))
    inst_sem = rng.integers(0, 21, 14)                        # 0,1 = stuff, 20 = unlabelled
    inst_sem[:3] = [0, 1, 20]
    sem = inst_sem[inst]
    d = dict(pts_instance_mask=inst.astype(np.int64), pts_semantic_mask=sem.astype(np.int64), sp_pts_mask=sp.astype(np.int64))
    out = PointDetClassMappingScanNet(num_classes=20, stuff_classes=[0, 1]).transform({k: v.copy() for k, v in d.items()})
    save.update({"sn_" + k: v for k, v in d.items()})
    save.update(sn_out_inst=out["pts_instance_mask"], sn_out_labels=out["gt_labels_3d"], sn_out_sp_masks=out["gt_sp_masks"].numpy())
    # S3DIS: instance ids 0..11, classes subset
    inst2 = rng.integers(0, 12, 60)[rng.integers(0, 60, n)]
    inst2[:12] = np.arange(12)
    sem2 = rng.integers(0, 13, 12)[inst2]
    d2 = dict(pts_instance_mask=inst2.astype(np.int64), pts_semantic_mask=sem2.astype(np.int64), sp_pts_mask=sp.astype(np.int64))
    out2 = PointDetClassMappingS3DIS(classes=[7, 8, 9, 10, 11]).transform({k: v.copy() for k, v in d2.items()})
    save.update({"s3_" + k: v for k, v in d2.items()})
    save.update(s3_out_inst=out2["pts_instance_mask"], s3_out_labels=out2["gt_labels_3d"].numpy(), s3_out_sp_masks=out2["gt_sp_masks"].numpy())
    # PointSample_: fixed choices through a patched sampler
    choices = rng.choice(n, 1800, replace=False)
    ps = PointSample_(num_points=1800)
    ps._points_random_sampling = lambda points, num: (points[choices], choices)
    d3 = dict(points=np.zeros((n, 6), np.float32), pts_instance_mask=out["pts_instance_mask"].copy(), pts_semantic_mask=sem.copy(),
              sp_pts_mask=sp.copy())
    out3 = ps.transform(d3)
    save.update(ps_choices=choices, ps_in_inst=out["pts_instance_mask"], ps_out_inst=out3["pts_instance_mask"],
                ps_out_sem=out3["pts_semantic_mask"], ps_out_sp=out3["sp_pts_mask"])
    np.savez_compressed(os.path.join(HERE, "gt_prep_ref.npz"), **save)
    print("gt_prep_ref.npz", out["gt_sp_masks"].shape, out2["gt_sp_masks"].shape, len(np.unique(out3["sp_pts_mask"])))


def gen_augment():
    """The reference's ElasticTransfrom (unidet3d/transforms_3d.py:12-83) on synthetic points, numpy RNG seeded."""
    import scipy.ndimage
    if not hasattr(scipy.ndimage, "filters"):
        scipy.ndimage.filters = scipy.ndimage
    from unidet3d.transforms_3d import ElasticTransfrom
    rng = np.random.default_rng(77)
    n = 3000
    pts = (rng.random((n, 6)) * np.array([7.5, 5.0, 2.8, 1, 1, 1]) - np.array([3.0, 2.0, 0.1, 0, 0, 0])).astype(np.float32)
    save = dict(points=pts)
    for tag, gran, mag, vs, p, seed in [("a", [6, 20], [40, 160], 0.02, 1.0, 11), ("b", [6, 20], [40, 160], 0.02, 0.1, 12),
                                        ("c", [4, 12], [20, 60], 0.05, 1.0, 13)]:
        np.random.seed(seed)
        t = ElasticTransfrom(gran=gran, mag=mag, voxel_size=vs, p=p)
        out = t.transform(dict(points=types.SimpleNamespace(tensor=torch.as_tensor(pts))))["elastic_coords"]
        save[f"{tag}_cfg"] = np.array([gran[0], gran[1], mag[0], mag[1], vs, p, seed], dtype=np.float64)
        save[f"{tag}_out"] = np.asarray(out)
        print("augment", tag, np.asarray(out).dtype, float(np.abs(np.asarray(out) - pts[:, :3] / vs).max()))
    np.savez_compressed(os.path.join(HERE, "augment_ref.npz"), **save)


def gen_evaluate():
    """The reference's own indoor_eval (unidet3d/indoor_eval.py) on synthetic detections / ground truth.  Boxes are stub
    objects whose ``overlaps`` is the oracle's restatement of mmdet3d's (third-party, absent): the fixture pins the
    evaluation LOGIC (class bookkeeping, TP / FP marking, AP, nan conventions), not the IoU."""
    import importlib.util
    for name, attrs in (("mmengine.logging", dict(print_log=lambda *a, **k: None)),
                        ("terminaltables", dict(AsciiTable=type("AsciiTable", (), {"__init__": lambda self, d: setattr(self, "table", ""),
                                                                                   "inner_footing_row_border": False})))):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_indoor_eval", os.path.join(REF, "unidet3d", "indoor_eval.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import evaluate as oev

    class Boxes:
        def __init__(self, t):
            self.tensor = torch.as_tensor(t, dtype=torch.float32).reshape(-1, 7)
        def __len__(self):
            return self.tensor.shape[0]
        def __getitem__(self, i):
            return Boxes(self.tensor[i].reshape(-1, 7))
        def new_box(self, t):
            return Boxes(t)
        def convert_to(self, mode):
            return self
        @classmethod
        def overlaps(cls, a, b):
            return torch.as_tensor(oev.overlaps_3d(a.tensor.numpy(), b.tensor.numpy()))

    rng = np.random.default_rng(21)
    n_img, n_cls = 6, 7                       # class 5 has detections but no ground truth, class 6 ground truth but no detections
    gt_b, gt_l, gt_i, dt_b, dt_s, dt_l, dt_i = [], [], [], [], [], [], []
    gt_annos, dt_annos = [], []
    for img in range(n_img):
        m = int(rng.integers(3, 9))
        gb = np.concatenate([rng.uniform(0, 6, (m, 3)), rng.uniform(0.3, 1.5, (m, 3)), np.zeros((m, 1))], 1).astype(np.float32)
        if img % 2:
            gb[:, 6] = rng.uniform(-1.5, 1.5, m)
        gl = rng.integers(0, 5, m)
        if img == 2:
            gl[0] = 6
        # detections: jittered copies of the ground truth (several per box, some duplicates of the same box) + clutter
        k = int(rng.integers(10, 30))
        src = rng.integers(0, m, k)
        db = gb[src] + rng.normal(0, 0.08, (k, 7)).astype(np.float32) * np.array([1, 1, 1, 1, 1, 1, 0], np.float32)
        dl = gl[src].copy()
        flip = rng.random(k) < 0.2
        dl[flip] = rng.integers(0, 6, int(flip.sum()))
        dl[dl == 6] = 0
        ds = rng.random(k).astype(np.float32)
        ds[:3] = ds[0]                                   # tied scores
        if img == 4:
            db, dl, ds = db[:0], dl[:0], ds[:0]          # an image without detections
        gt_b.append(gb); gt_l.append(gl); gt_i.append(np.full(m, img)); dt_b.append(db); dt_s.append(ds); dt_l.append(dl)
        dt_i.append(np.full(len(dl), img))
        gt_annos.append(dict(gt_bboxes_3d=Boxes(gb), gt_labels_3d=[int(x) for x in gl]))
        dt_annos.append(dict(labels_3d=torch.as_tensor(dl), bboxes_3d=Boxes(db), scores_3d=torch.as_tensor(ds)))
    label2cat = {i: f"c{i}" for i in range(n_cls)}
    # the reference sorts with np.argsort(-confidence) (unstable for ties): make the order reproducible for the fixture
    _argsort = np.argsort
    np.argsort = lambda a, *args, **kw: _argsort(a, kind="stable")
    try:
        ret = mod.indoor_eval(gt_annos, dt_annos, [0.25, 0.5], label2cat)
    finally:
        np.argsort = _argsort
    save = dict(gt_boxes=np.concatenate(gt_b), gt_labels=np.concatenate(gt_l), gt_img=np.concatenate(gt_i),
                det_boxes=np.concatenate(dt_b), det_scores=np.concatenate(dt_s), det_labels=np.concatenate(dt_l),
                det_img=np.concatenate(dt_i), metric=np.array([0.25, 0.5]),
                keys=np.array(list(ret.keys())), values=np.array([ret[k] for k in ret], dtype=np.float64))
    np.savez_compressed(os.path.join(HERE, "evaluate_ref.npz"), **save)
    print("evaluate_ref.npz", {k: round(v, 4) for k, v in ret.items() if k.startswith("m")}, len(ret))


def gen_post():
    from unidet3d.unidet3d import UniDet3D, get_face_distances
    from unidet3d.encoder import _bbox_pred_to_bbox
    torch.manual_seed(5)
    rng = np.random.default_rng(9)
    save = {}
    n_pts, S, T = 4000, 60, 90
    pts = torch.as_tensor(rng.uniform(0, 4, (n_pts, 3)).astype(np.float32))
    g = np.floor(pts.numpy() / 1.0).astype(np.int64)
    _, sp = np.unique((g[:, 0] * 8 + g[:, 1]) * 8 + g[:, 2], return_inverse=True)
    sp = torch.as_tensor(sp)
    centers = torch.as_tensor(rng.uniform(0.5, 3.5, (T, 3)).astype(np.float32))
    save.update(points=pts.numpy(), sp=sp.numpy(), centers=centers.numpy())
    for tag, ncls, fast, angle, use_sp, thr in [("scannet", 5, True, False, True, 0.5),
                                                ("s3dis", 4, False, False, True, 0.55),
                                                ("arkit", 6, None, True, False, 0.55)]:
        det = object.__new__(UniDet3D)
        det.__dict__["_modules"] = {}
        det.__dict__["_parameters"] = {}
        det.__dict__["_buffers"] = {}
        object.__setattr__(det, "test_cfg", types.SimpleNamespace(topk_insts=300, score_thr=0.0, iou_thr=[thr],
                                                                 low_sp_thr=0.18, up_sp_thr=0.81))
        object.__setattr__(det, "fast_nms", [fast])
        object.__setattr__(det, "use_superpoints", [use_sp])
        object.__setattr__(det, "decoder", types.SimpleNamespace(datasets=[tag]))
        cls = torch.randn(T, ncls + 1) * 2
        raw = torch.cat([torch.as_tensor(rng.uniform(0.2, 1.0, (T, 6)).astype(np.float32)),
                         torch.randn(T, 2) * 0.5], 1)
        box = _bbox_pred_to_bbox(centers, raw if angle else raw[:, :6])
        out = dict(cls_preds=[cls], bboxes=[box])
        res = det.predict_by_feat(out, [sp], [pts], [tag])
        bb, labels, scores = res[0]
        save.update({f"{tag}_cls": cls.numpy(), f"{tag}_raw": raw.numpy(), f"{tag}_box": box.numpy(),
                     f"{tag}_out_boxes": bb.tensor.numpy(), f"{tag}_out_labels": labels.numpy(),
                     f"{tag}_out_scores": scores.numpy()})
        print("post", tag, bb.tensor.shape)
    # face distances on rotated boxes
    boxes = torch.cat([centers[:7], torch.as_tensor(rng.uniform(0.5, 2, (7, 3)).astype(np.float32)),
                       torch.as_tensor(rng.uniform(-3, 3, (7, 1)).astype(np.float32))], 1)
    fd = get_face_distances(pts[:500].unsqueeze(1).expand(500, 7, 3), boxes.unsqueeze(0).expand(500, 7, 7))
    save.update(fd_boxes=boxes.numpy(), fd_out=fd.numpy())
    np.savez_compressed(os.path.join(HERE, "post_ref.npz"), **save)


if __name__ == "__main__":
    assert os.path.isdir(REF), "reference checkout not present: goldens can only be generated in the build container"
    install_stubs()
    which = sys.argv[1:] or ["encoder", "unet", "post", "criterion", "criterion_grad", "backward", "train_step", "predict", "collate", "gt_prep", "augment", "evaluate"]
    for name in which:
        {"encoder": gen_encoder, "unet": gen_unet, "post": gen_post, "criterion": gen_criterion, "criterion_grad": gen_criterion_grad, "backward": gen_backward, "train_step": gen_train_step, "predict": gen_predict, "collate": gen_collate, "gt_prep": gen_gt_prep, "augment": gen_augment, "evaluate": gen_evaluate}[name]()
