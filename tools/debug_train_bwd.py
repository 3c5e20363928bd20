"""Debug helper: where the backbone backward (unidet3d_b200/train.py) departs from torch.autograd through the oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unidet3d_b200 as u  # noqa: E402
from unidet3d_b200 import configs, train, ops  # noqa: E402
from unidet3d_b200.synthetic import make_model_state_dict, make_scene, SCENE_PRESETS  # noqa: E402
from oracle import spconv as ospconv, unet as ounet, voxelize as ovox  # noqa: E402
from oracle.spconv import sparse_conv, weight_to_koc  # noqa: E402
from oracle.pool import superpoint_pool  # noqa: E402

DEV = "cuda"


def relerr(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


cfg = configs.model_cfg(("scannet",), topk_insts=100)
n, v, a, c = SCENE_PRESETS["tiny"]
cfg["voxel_size"] = v
model = u.MODELS.build(cfg)
sd = make_model_state_dict(cfg, 0)
model.load_state_dict(sd, strict=False)
model.to(DEV).train()
scenes = [make_scene(70 + i, n, a, c) for i in range(2)]
pts, sps = [s[0] for s in scenes], [s[1] for s in scenes]
P = torch.as_tensor(np.concatenate(pts)).to(DEV)
offs = torch.tensor(np.cumsum([0] + [len(p) for p in pts]), dtype=torch.int32, device=DEV)
n_sps = [int(s.max()) + 1 for s in sps]
sp_off = np.concatenate([[0], np.cumsum(n_sps)])
sp_all = np.concatenate([s + o for s, o in zip(sps, sp_off[:-1])])
d_pooled = torch.randn(int(sp_off[-1]), 32, generator=torch.Generator().manual_seed(9))
with torch.no_grad():
    x, inv = model.collate(P, offs, 2)
    pooled, tape = train.backbone_forward(model, x, torch.as_tensor(sp_all).to(DEV), inv, int(sp_off[-1]))
    train.backbone_backward(tape, pooled, d_pooled.to(DEV))
det_sd = {k: t.clone().float() for k, t in sd.items() if not k.startswith("decoder.")}
params = {k: t.requires_grad_(True) for k, t in det_sd.items()
          if t.is_floating_point() and not k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
coords, feats, inverse, shape = ovox.voxelize(pts, v, 128)
unet_sd = {k[len("unet."):]: t for k, t in det_sd.items() if k.startswith("unet.")}
levels = ounet.build_pyramid(coords, shape, ounet.n_levels_of(unet_sd))
ospconv.TRAIN_MODE = True
stem = sparse_conv(torch.as_tensor(feats), levels[0]["subm"], weight_to_koc(det_sd["input_conv.0.weight"]))
stem.retain_grad()
xo_pre = ounet.unet_forward(unet_sd, stem, levels)
xo_pre.retain_grad()
xo = ospconv.bn_relu(xo_pre, ounet._bn(det_sd, "output_layer.0"))
xo.retain_grad()
ospconv.TRAIN_MODE = False
ref = superpoint_pool(xo, inverse, sp_all, int(sp_off[-1]))
(ref * d_pooled).sum().backward()
print("fwd pooled", relerr(pooled, ref), "unet_out", relerr(tape.named["unet_out"], xo_pre), "stem", relerr(tape.named["stem"], stem))
print("grad unet_out", relerr(tape.grad(tape.named["unet_out"]), xo_pre.grad), "grad stem", relerr(tape.grad(tape.named["stem"]), stem.grad))
da = ops.segmented_mean_backward(d_pooled.to(DEV), torch.as_tensor(sp_all).to(DEV), xo.shape[0], gather=inv)
print("pool bwd", relerr(da, xo.grad))
got = {k: p.grad for k, p in model.named_parameters() if not k.startswith("decoder.")}
for k in ("output_layer.0.weight", "output_layer.0.bias", "unet.blocks_tail.block1.conv_branch.5.weight", "unet.blocks_tail.block1.conv_branch.3.weight",
          "unet.blocks_tail.block1.conv_branch.2.weight", "unet.blocks_tail.block0.i_branch.0.weight", "unet.deconv.2.weight", "unet.conv.2.weight",
          "unet.blocks.block0.conv_branch.2.weight", "input_conv.0.weight"):
    print(f"{k:50s} {relerr(got[k], params[k].grad):.3e}")
# ---- the output layer alone, with torch ops on the GPU tensors
y = tape.named["unet_out"]
bn = model.output_layer[0]
print("bn eps", bn.eps, "momentum", bn.momentum, type(bn).__name__, "training", bn.training)
sc, sh, mean, invstd = ops.bn_train(y, bn, update_running=False)
mask = (y * sc + sh > 0).float()
g = da * mask
print("dbeta torch-on-gpu vs oracle", relerr(g.sum(0), params["output_layer.0.bias"].grad))
xhat = (y - mean) * invstd
print("dgamma torch-on-gpu vs oracle", relerr((g * xhat).sum(0), params["output_layer.0.weight"].grad))
dx, dgamma, dbeta = ops.bn_relu_backward(y, da, sc, sh, mean, invstd)
print("kernel dbeta vs torch-on-gpu", relerr(dbeta, g.sum(0)), "dgamma", relerr(dgamma, (g * xhat).sum(0)))
# oracle-side statistics
xp = xo_pre.detach()
print("mean vs oracle", relerr(mean, xp.mean(0)), "invstd", relerr(invstd, 1.0 / torch.sqrt(xp.var(0, unbiased=False) + 1e-4)))
w, b = det_sd["output_layer.0.weight"].detach(), det_sd["output_layer.0.bias"].detach()
print("scale", relerr(sc, w / torch.sqrt(xp.var(0, unbiased=False) + 1e-4)), "gamma equal", relerr(bn.weight, w), relerr(bn.bias, b))
mask_o = (xo.detach() > 0).float()
print("mask mismatches", float((mask.cpu() != mask_o).float().mean()), "oracle dbeta from mask", relerr((xo.grad * mask_o).sum(0), params["output_layer.0.bias"].grad))
