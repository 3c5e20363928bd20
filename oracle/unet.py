"""Oracle: SpConvUNet forward (reference: unidet3d/spconv_unet.py:13-240) and the
detector-side input conv / output BN+ReLU (unidet3d/unidet3d.py:95-134).

Functional restatement over a reference-layout ``state_dict`` (same key names and
shapes as the reference modules: ``blocks.block0.conv_branch.{0,2,3,5}``,
``conv.{0,2}``, ``u.*``, ``deconv.{0,2}``, ``blocks_tail.block{0,1}.*``;
conv weights ``[C_out,k,k,k,C_in]``).  Eval-mode BatchNorm (eps 1e-4) only.
"""
import numpy as np
import torch

from . import rulebook
from .spconv import sparse_conv, weight_to_koc, bn_relu


def build_pyramid(coords, spatial_shape, n_levels):
    """Rulebooks for every level: list of dict(coords, shape, subm, child, up)."""
    levels = []
    c, s = np.asarray(coords, np.int32), np.asarray(spatial_shape, np.int64)
    for l in range(n_levels):
        lv = dict(coords=c, shape=s, subm=rulebook.subm3_table(c, s))
        if l + 1 < n_levels:
            cc, child, up, out_shape = rulebook.down2(c, s)
            lv.update(child=child, up=up)
            c, s = cc, out_shape
        levels.append(lv)
    return levels


def _bn(sd, p):
    return (sd[p + ".weight"], sd[p + ".bias"], sd[p + ".running_mean"], sd[p + ".running_var"])


def residual_block(sd, p, x, subm):
    """spconv_unet.py:74-91 (normalize_before=True branch :41-56)."""
    if (p + ".i_branch.0.weight") in sd:
        w = sd[p + ".i_branch.0.weight"]                   # [C,1,1,1,C_in]
        identity = x @ w.reshape(w.shape[0], w.shape[-1]).t()
    else:
        identity = x
    y = bn_relu(x, _bn(sd, p + ".conv_branch.0"))
    y = sparse_conv(y, subm, weight_to_koc(sd[p + ".conv_branch.2.weight"]))
    y = bn_relu(y, _bn(sd, p + ".conv_branch.3"))
    y = sparse_conv(y, subm, weight_to_koc(sd[p + ".conv_branch.5.weight"]))
    return y + identity


def unet_forward(sd, x, levels, l=0, p=""):
    """spconv_unet.py:205-240. x: [N_l, C_l] features at level l."""
    lv = levels[l]
    block_reps = 0
    while (p + f"blocks.block{block_reps}.conv_branch.2.weight") in sd:
        block_reps += 1
    for i in range(block_reps):
        x = residual_block(sd, p + f"blocks.block{i}", x, lv["subm"])
    if (p + "conv.2.weight") in sd:
        identity = x
        d = bn_relu(x, _bn(sd, p + "conv.0"))
        d = sparse_conv(d, lv["child"], weight_to_koc(sd[p + "conv.2.weight"]))
        d = unet_forward(sd, d, levels, l + 1, p + "u.")
        d = bn_relu(d, _bn(sd, p + "deconv.0"))
        u = sparse_conv(d, lv["up"], weight_to_koc(sd[p + "deconv.2.weight"]))
        x = torch.cat((identity, u), dim=1)
        for i in range(block_reps):
            x = residual_block(sd, p + f"blocks_tail.block{i}", x, lv["subm"])
    return x


def n_levels_of(sd, p=""):
    n = 1
    while (p + "conv.2.weight") in sd:
        n += 1
        p += "u."
    return n


def backbone_forward(det_sd, coords, feats, spatial_shape):
    """unidet3d.py:127-129: input_conv -> unet -> output_layer.  ``det_sd`` uses the
    detector's key names (``input_conv.0.weight``, ``unet.*``, ``output_layer.0.*``)."""
    unet_sd = {k[len("unet."):]: v for k, v in det_sd.items() if k.startswith("unet.")}
    levels = build_pyramid(coords, spatial_shape, n_levels_of(unet_sd))
    x = sparse_conv(feats, levels[0]["subm"], weight_to_koc(det_sd["input_conv.0.weight"]))
    x = unet_forward(unet_sd, x, levels)
    x = bn_relu(x, _bn(det_sd, "output_layer.0"))
    return x, levels


from unidet3d_b200.synthetic import make_unet_state_dict, make_detector_backbone_state_dict  # noqa: E402,F401  (shared synthetic weights)
