"""Generate golden fixtures by EXECUTING THE REFERENCE'S OWN PYTHON in the build container.

Run (build container only -- /root/reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/{encoder_ref,unet_ref,post_ref}.npz.  The reference modules are
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
    def register_module(self, *a, **k):
        return lambda cls: cls

    def build(self, cfg):
        raise RuntimeError("not used")


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
    _mod("mmengine.structures", InstanceData=InstanceData)
    _mod("mmdet3d"); _mod("mmdet3d.registry", MODELS=reg, TASK_UTILS=reg)

    class DepthInstance3DBoxes:
        def __init__(self, tensor, box_dim=7, with_yaw=True, origin=(0.5, 0.5, 0)):
            self.tensor, self.box_dim, self.with_yaw = tensor, box_dim, with_yaw

    def rotation_3d_in_axis(points, angles, axis=0):
        assert axis in (2, -1)
        s, c = torch.sin(angles), torch.cos(angles)
        o, z = torch.ones_like(c), torch.zeros_like(c)
        rot_t = torch.stack([torch.stack([c, s, z]), torch.stack([-s, c, z]), torch.stack([z, z, o])])
        return torch.einsum("aij,jka->aik", points, rot_t)

    class Base3DDetector(nn.Module):
        pass
    _mod("mmdet3d.structures", DepthInstance3DBoxes=DepthInstance3DBoxes, rotation_3d_in_axis=rotation_3d_in_axis)
    _mod("mmdet3d.models", Base3DDetector=Base3DDetector)
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
    gen_encoder()
    gen_unet()
    gen_post()
