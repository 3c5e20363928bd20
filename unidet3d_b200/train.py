"""Backward pass of the backbone in train mode: gradients of every parameter of ``input_conv`` / ``SpConvUNet`` /
``output_layer`` (reference unidet3d/unidet3d.py:95-134, spconv_unet.py:13-240 under torch.autograd) from the
gradient of the pooled superpoint features, on the library's kernels:

* sparse conv: weight gradient ``ud3d_conv_wgrad``; input gradient = the forward gather-GEMM on dY with the transposed
  weight over the transposed rulebook (SubM3: same table, offsets reversed; k2s2 conv <-> its inverse conv's table);
* train-mode (Sync)BatchNorm + ReLU: ``ud3d_bn_backward_sums`` / ``_apply`` (one all-reduce of 2C + 1 doubles per
  BatchNorm under torch.distributed);
* superpoint mean-pool: ``ud3d_segmented_mean_backward``.

The forward is the train-mode executor of ``SpConvUNet`` re-run op by op while a tape records one closure per op; the
backward replays the tape in reverse, accumulating gradients per tensor (a residual input or the skip concat has two
consumers).  Activations after BatchNorm + ReLU are recomputed, not stored.  Scope: the backbone only -- the encoder /
criterion backward kernels do not exist yet, so this is not a training step.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .spconv_unet import SparseConvWeight, SpConvUNet


class Tape:
    def __init__(self, group=None):
        self.steps, self.grads, self.keep, self.group = [], {}, [], group
        self.named = {}                       # a few intermediate tensors by name (tests / debugging)

    def grad(self, t: torch.Tensor) -> Optional[torch.Tensor]:
        return self.grads.get(id(t))

    def add(self, t: torch.Tensor, g: torch.Tensor):
        self.keep.append(t)                       # ids stay unique while the tape is alive
        cur = self.grads.get(id(t))
        self.grads[id(t)] = g if cur is None else cur + g

    def backward(self):
        for step in reversed(self.steps):
            step()


def _acc_param(p: torch.nn.Parameter, g: torch.Tensor):
    g = g.reshape(p.shape).to(p.dtype)
    p.grad = g if p.grad is None else p.grad + g


def conv(tape: Tape, x: torch.Tensor, conv_mod: SparseConvWeight, K: int, table, mask, n_out: int, table_t, reverse: bool,
         bn=None, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         need_input_grad: bool = True) -> torch.Tensor:
    """y = sparse_conv(relu(bn_train(x)) if bn else x) (+ residual); records its backward."""
    s = h = mean = invstd = None
    if bn is not None:
        s, h, mean, invstd = ops.bn_train(x, bn, group=tape.group)
    y = ops.gemm(x, ops.PackedWeight(conv_mod.weight), table=table, tile_mask=mask, n_out=n_out, in_scale=s, in_shift=h,
                 in_relu=bn is not None, residual=residual, out=out)
    n_in = x.shape[0]

    def bwd():
        dy = tape.grad(y)
        if dy is None:
            return
        if residual is not None:
            tape.add(residual, dy)
        a = ops.bn_relu_apply(x, s, h) if bn is not None else x
        _acc_param(conv_mod.weight, ops.conv_wgrad(a, dy, K, table))
        if not need_input_grad:
            return
        da = ops.conv_dgrad(dy, conv_mod.weight.detach(), table_t, n_in, reverse_offsets=reverse)
        if bn is None:
            tape.add(x, da)
            return
        dx, dgamma, dbeta = ops.bn_relu_backward(x, da, s, h, mean, invstd, group=tape.group)
        _acc_param(bn.weight, dgamma)
        _acc_param(bn.bias, dbeta)
        tape.add(x, dx)

    tape.steps.append(bwd)
    return y


def _block(tape, blk, x, lv):
    cb = blk.conv_branch
    if isinstance(blk.i_branch[0], SparseConvWeight):
        identity = conv(tape, x, blk.i_branch[0], 1, None, None, x.shape[0], None, False)
    else:
        identity = x
    y = conv(tape, x, cb[2], 27, lv.subm, lv.subm_mask, lv.n, lv.subm, True, bn=cb[0])
    return conv(tape, y, cb[5], 27, lv.subm, lv.subm_mask, lv.n, lv.subm, True, bn=cb[3], residual=identity)


def unet_forward(tape: Tape, unet: SpConvUNet, x: torch.Tensor, pyr, l: int = 0) -> torch.Tensor:
    """spconv_unet.py:205-240 in train mode (the op sequence of SpConvUNet._forward_level_train) with the tape."""
    lv = pyr.levels[l]
    c = unet.num_planes[0]
    for blk in unet.blocks:
        x = _block(tape, blk, x, lv)
    if len(unet.num_planes) > 1:
        nxt = pyr.levels[l + 1]
        cat = torch.empty((lv.n, 2 * c), dtype=torch.float32, device=x.device)
        cat[:, :c] = x
        x_in = x
        d = conv(tape, x, unet.conv[2], 8, lv.child, lv.child_mask, nxt.n, lv.up, False, bn=unet.conv[0])
        d = unet_forward(tape, unet.u, d, pyr, l + 1)
        up = conv(tape, d, unet.deconv[2], 8, lv.up, lv.up_mask, lv.n, lv.child, False, bn=unet.deconv[0], out=cat[:, c:])

        def split():                                   # gradient of the skip concat [identity | decoder]
            dc = tape.grad(cat)
            if dc is not None:
                tape.add(x_in, dc[:, :c])
                tape.add(up, dc[:, c:])

        tape.steps.append(split)
        x = cat
        for blk in unet.blocks_tail:
            x = _block(tape, blk, x, lv)
    return x


def backbone_forward(model, x, superpoints: torch.Tensor, inverse_mapping: torch.Tensor, n_superpoints: int, group=None):
    """``UniDet3D.extract_feat`` in train mode with a tape.  -> (pooled [n_sp, C], tape).  ``model.training`` must be True
    (running statistics are updated like in any train-mode forward)."""
    tape = Tape(group)
    lv0 = x.pyramid.levels[0]
    f = conv(tape, x.features, model.input_conv[0], 27, lv0.subm, lv0.subm_mask, lv0.n, lv0.subm, True, need_input_grad=False)
    y = unet_forward(tape, model.unet, f, x.pyramid)
    tape.named.update(stem=f, unet_out=y)
    bn = model.output_layer[0]
    sc, sh, mean, invstd = ops.bn_train(y, bn, group=group)
    pooled = ops.segmented_mean(y, superpoints, n_superpoints, gather=inverse_mapping, scale=sc, shift=sh, relu=True)

    def bwd():
        dp = tape.grad(pooled)
        da = ops.segmented_mean_backward(dp.contiguous(), superpoints, y.shape[0], gather=inverse_mapping)
        dy, dgamma, dbeta = ops.bn_relu_backward(y, da, sc, sh, mean, invstd, group=group)
        _acc_param(bn.weight, dgamma)
        _acc_param(bn.bias, dbeta)
        tape.add(y, dy)

    tape.steps.append(bwd)
    return pooled, tape


def backbone_backward(tape: Tape, pooled: torch.Tensor, d_pooled: torch.Tensor):
    """Fills ``.grad`` of every backbone parameter (accumulating into existing gradients like torch)."""
    tape.add(pooled, d_pooled)
    with torch.no_grad():
        tape.backward()


def allreduce_gradients(params, group=None, bucket_bytes: int = 32 << 20):
    """Data-parallel gradient exchange (tools/train.py:49-52 runs the reference under DDP): average the gradients over
    the ranks in buckets of ``bucket_bytes`` (one flat buffer per bucket, one all-reduce each)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    n_coll, i = 0, 0
    while i < len(grads):
        j, size = i, 0
        while j < len(grads) and (j == i or size + grads[j].numel() * 4 <= bucket_bytes):
            size += grads[j].numel() * 4
            j += 1
        flat = torch.cat([g.reshape(-1) for g in grads[i:j]])
        dist.all_reduce(flat, group=group)
        flat /= world
        off = 0
        for g in grads[i:j]:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_coll += 1
        i = j
    return n_coll
