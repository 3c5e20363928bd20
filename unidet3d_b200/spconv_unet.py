"""SpConvUNet: the sparse-conv residual U-Net backbone (reference: unidet3d/spconv_unet.py:13-240),
re-implemented on the tcgen05 gather-GEMM.  Same registry name, constructor arguments and
state_dict keys/shapes as the reference; no spconv import.

Execution plan (eval mode):
  * every BatchNorm(eval)+ReLU that precedes a conv is folded into that conv's operand load
    (``in_scale``/``in_shift``/``in_relu``), every residual add into the conv epilogue, and the skip
    concat ``[identity | decoder]`` is never materialised: the encoder-side block and the inverse
    conv write straight into the two column halves of one [N, 2c] buffer;
  * => each feature map is read once and written once per conv.
"""
from __future__ import annotations

import functools
from collections import OrderedDict
from typing import List, Optional

import torch
from torch import nn

import ctypes as C

from . import _lib, ops
from .registry import register_model
from .rulebook import Pyramid, build_pyramid
from .structures import SparseConvTensor


class SparseConvWeight(nn.Module):
    """Parameter holder with the spconv-2.x layout ``[C_out, k, k, k, C_in]`` (no bias)."""

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        k = kernel_size
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, k
        w = torch.empty(out_channels, k, k, k, in_channels)
        nn.init.kaiming_uniform_(w.view(out_channels, -1), a=5 ** 0.5)
        self.weight = nn.Parameter(w)


def fold_bn(bn: nn.Module):
    """Eval-mode BatchNorm -> per-channel (scale, shift): y = x*scale + shift."""
    inv = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    scale = bn.weight.detach().float() * inv
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale.contiguous(), shift.contiguous()


class ResidualBlock(nn.Module):
    """Parameter tree of the reference block (spconv_unet.py:13-72); executed by SpConvUNet."""

    def __init__(self, in_channels, out_channels, norm_fn, indice_key=None, normalize_before=True):
        super().__init__()
        if not normalize_before:
            raise NotImplementedError("only normalize_before=True is supported (the reference configs use it)")
        if in_channels == out_channels:
            self.i_branch = nn.Sequential(nn.Identity())
        else:
            self.i_branch = nn.Sequential(SparseConvWeight(in_channels, out_channels, 1))
        self.conv_branch = nn.Sequential(
            norm_fn(in_channels), nn.ReLU(), SparseConvWeight(in_channels, out_channels, 3),
            norm_fn(out_channels), nn.ReLU(), SparseConvWeight(out_channels, out_channels, 3))
        self.in_channels, self.out_channels = in_channels, out_channels


@register_model
class SpConvUNet(nn.Module):
    """SpConv U-Net model (drop-in for the reference class of the same name).

    Args mirror unidet3d/spconv_unet.py:108-115.  The reference's recursion passes ``norm_fn``
    positionally into ``use_sync_bn`` (:166-173); any truthy value selects SyncBatchNorm there.  Here
    both variants hold plain BatchNorm statistics (identical state_dict keys; eval math identical).
    """

    def __init__(self, num_planes, use_sync_bn=True, block_reps=2, block=ResidualBlock, indice_key_id=1,
                 normalize_before=True, return_blocks=False):
        super().__init__()
        self.return_blocks = return_blocks
        self.num_planes = list(num_planes)
        self.block_reps = block_reps
        self.indice_key_id = indice_key_id
        norm_fn = functools.partial(nn.BatchNorm1d, eps=1e-4, momentum=0.1)
        if isinstance(block, str):
            assert block in ("residual",), f"only the residual block is implemented, got {block}"
            block = ResidualBlock
        c = num_planes[0]
        self.blocks = nn.Sequential(OrderedDict(
            (f"block{i}", block(c, c, norm_fn, normalize_before=normalize_before, indice_key=f"subm{indice_key_id}"))
            for i in range(block_reps)))
        if len(num_planes) > 1:
            if not normalize_before:
                raise NotImplementedError("only normalize_before=True is supported")
            self.conv = nn.Sequential(norm_fn(c), nn.ReLU(), SparseConvWeight(c, num_planes[1], 2))
            self.u = SpConvUNet(num_planes[1:], use_sync_bn, block_reps, block, indice_key_id=indice_key_id + 1,
                                normalize_before=normalize_before, return_blocks=return_blocks)
            self.deconv = nn.Sequential(norm_fn(num_planes[1]), nn.ReLU(), SparseConvWeight(num_planes[1], c, 2))
            self.blocks_tail = nn.Sequential(OrderedDict(
                (f"block{i}", block(c * (2 - i), c, norm_fn, indice_key=f"subm{indice_key_id}",
                                    normalize_before=normalize_before)) for i in range(block_reps)))
        self._plan = None
        self._cplan = None
        # eval-mode executor: True = the whole recursion issued by one C call (ud3d_unet_forward, csrc/unet_plan.cu);
        # False = the Python recursion below (_forward_level), one ctypes call per convolution.  Same kernels, same
        # arguments, bit-identical results (tests/test_gpu_model.py); the C plan saves ~1.2 ms of host time per batch.
        self.use_stage_plan = True
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate_plan())

    # ------------------------------------------------------------------ plan (packed weights, folded BN)
    def invalidate_plan(self):
        self._plan = None
        self._cplan = None
        if hasattr(self, "u"):
            self.u.invalidate_plan()

    def _apply(self, fn, *a, **k):
        self.invalidate_plan()
        return super()._apply(fn, *a, **k)

    def first_bn(self):
        """Folded (scale, shift) of ``blocks.block0.conv_branch.0``: lets the producer of the input
        feature map emit its operand-form copy directly (``SparseConvTensor.features_act``)."""
        return self._get_plan()["blocks"][0]["bn0"]

    @staticmethod
    def _block_plan(blk: ResidualBlock):
        cb = blk.conv_branch
        p = dict(bn0=fold_bn(cb[0]), w0=ops.PackedWeight(cb[2].weight), bn1=fold_bn(cb[3]), w1=ops.PackedWeight(cb[5].weight))
        if isinstance(blk.i_branch[0], SparseConvWeight):
            p["wi"] = ops.PackedWeight(blk.i_branch[0].weight)
        return p

    def _get_plan(self):
        if self._plan is None:
            p = dict(blocks=[self._block_plan(b) for b in self.blocks])
            if len(self.num_planes) > 1:
                p["down_bn"] = fold_bn(self.conv[0]); p["down_w"] = ops.PackedWeight(self.conv[2].weight)
                p["up_bn"] = fold_bn(self.deconv[0]); p["up_w"] = ops.PackedWeight(self.deconv[2].weight)
                p["tail"] = [self._block_plan(b) for b in self.blocks_tail]
            self._plan = p
        return self._plan

    def _c_plan(self):
        """ctypes mirror (``ud3d_unet_plan``) of the plan tree; the tensors it points to are owned by ``_get_plan()``."""
        if self._cplan is None:
            mods = [self]
            while hasattr(mods[-1], "u"):
                mods.append(mods[-1].u)
            if len(mods) > 8 or self.block_reps > 4:
                raise RuntimeError("ud3d_unet_plan holds at most 8 levels and 4 blocks per stage")
            P = _lib.UnetPlan()
            P.n_levels, P.block_reps = len(mods), self.block_reps

            def fill(dst, bp):
                dst.w0, dst.w1 = bp["w0"].data.data_ptr(), bp["w1"].data.data_ptr()
                dst.wi = bp["wi"].data.data_ptr() if "wi" in bp else None
                dst.bn0_scale, dst.bn0_shift = bp["bn0"][0].data_ptr(), bp["bn0"][1].data_ptr()
                dst.bn1_scale, dst.bn1_shift = bp["bn1"][0].data_ptr(), bp["bn1"][1].data_ptr()

            for l, m in enumerate(mods):
                p, L = m._get_plan(), P.level[l]
                L.c = m.num_planes[0]
                for i, bp in enumerate(p["blocks"]):
                    fill(L.blocks[i], bp)
                if "tail" in p:
                    for i, bp in enumerate(p["tail"]):
                        fill(L.tail[i], bp)
                    L.down_w, L.up_w = p["down_w"].data.data_ptr(), p["up_w"].data.data_ptr()
                    L.down_scale, L.down_shift = p["down_bn"][0].data_ptr(), p["down_bn"][1].data_ptr()
                    L.up_scale, L.up_shift = p["up_bn"][0].data_ptr(), p["up_bn"][1].data_ptr()
            self._cplan = P
        return self._cplan

    def _forward_plan(self, x_raw, x_act, pyr: Pyramid, outputs: list, want_blocks: bool):
        """Eval forward through ud3d_unet_forward.  Returns the fp32 level-0 output."""
        nl = self.n_levels()
        P = self._c_plan()
        T = pyr.c_tables(nl)
        lib = _lib.load()
        dev = x_raw.device
        assert x_raw.is_contiguous() and x_act.is_contiguous() and x_raw.shape[1] == self.num_planes[0] == x_act.shape[1]
        ws_bytes = int(lib.ud3d_unet_workspace_bytes(C.byref(P), T))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out = torch.empty((pyr.levels[0].n, self.num_planes[0]), dtype=torch.float32, device=dev)
        level_out, keep = None, []
        if want_blocks:
            level_out = (C.c_void_p * nl)()
            m = self
            for l in range(1, nl):
                m = m.u
                keep.append(torch.empty((pyr.levels[l].n, m.num_planes[0]), dtype=torch.float32, device=dev))
                level_out[l] = keep[-1].data_ptr()
        ops.check(lib.ud3d_unet_forward(C.byref(P), T, x_raw.data_ptr(), x_act.data_ptr(), out.data_ptr(), level_out,
                                        ws.data_ptr(), ws_bytes, ops._stream()), "ud3d_unet_forward")
        for l in range(nl - 1, 0, -1):                      # deepest level first, like the recursion
            outputs.append((l, keep[l - 1] if want_blocks else None))
        outputs.append((0, out))
        return out

    # ------------------------------------------------------------------ execution
    # Feature maps travel between convs in "operand form": relu(bn(x)) already applied (the CONSUMER's
    # folded BatchNorm), split into bf16 hi|lo -- the exact bytes of the tensor-core tile, gathered by
    # cp.async with no per-use conversion.  A conv epilogue emits one operand-form copy per consumer
    # BatchNorm (at most two) plus the raw fp32 map only where it is needed as a residual / output.
    @staticmethod
    def _conv(x_act, w, table, mask, n_out, *, residual=None, raw=None, want_raw=True, acts=(), perm=None):
        """x_act: operand-form input.  acts: sequence of (buffer|None, (scale, shift)).  Returns (raw, [act bufs])."""
        dev = x_act.device
        bufs = []
        for buf, bn in acts:
            if buf is None:
                buf = torch.empty((n_out, w.c_out), dtype=torch.float32, device=dev)
            bufs.append((buf, bn[0], bn[1]))
        out = ops.gemm(x_act, w, table=table, tile_mask=mask, n_out=n_out, in_split=True, residual=residual, out=raw,
                       no_raw=not want_raw, acts=bufs, row_perm=perm)
        return (out if want_raw else None), [b[0] for b in bufs]

    def _run_block(self, bp, x_raw, x_act, lv, *, out_raw=None, want_raw=True, out_acts=()):
        """ResidualBlock (spconv_unet.py:74-91) with equal channel counts: x + SubM3(BN.SubM3(BN.x)).
        x_act = operand form of relu(bn0(x))."""
        tb, tm, pm = lv.subm_conv
        _, (y_act,) = self._conv(x_act, bp["w0"], tb, tm, lv.n, want_raw=False, acts=[(None, bp["bn1"])], perm=pm)
        return self._conv(y_act, bp["w1"], tb, tm, lv.n, residual=x_raw, raw=out_raw, want_raw=want_raw,
                          acts=out_acts, perm=pm)

    def _forward_level(self, x_raw, x_act, pyr: Pyramid, l: int, outputs: list, out_bn=None):
        """x_raw: fp32 [N_l, c]; x_act: operand form of relu(blocks.block0.bn0(x)).
        out_bn: folded BN of the parent's ``deconv.0`` (None at the top level).
        Returns (raw fp32 level output, operand-form output for the parent or None)."""
        plan = self._get_plan()
        lv = pyr.levels[l]
        c = self.num_planes[0]
        dev = x_raw.device
        blocks = plan["blocks"]
        has_sub = len(self.num_planes) > 1
        out_acts_final = [(None, out_bn)] if out_bn is not None else []
        if not has_sub:
            raw, act = x_raw, x_act
            for i, bp in enumerate(blocks):
                last = i == len(blocks) - 1
                acts = out_acts_final if last else [(None, blocks[i + 1]["bn0"])]
                raw, a = self._run_block(bp, raw, act, lv, out_acts=acts)
                act = a[0] if a else None
            outputs.append((l, raw))
            return raw, act
        tail = plan["tail"]
        cat_raw = torch.empty((lv.n, 2 * c), dtype=torch.float32, device=dev)     # [identity | decoder], never concatenated
        cat_act = torch.empty((lv.n, 2 * c), dtype=torch.float32, device=dev)     # operand form under tail.block0.bn0
        t_sc, t_sh = tail[0]["bn0"]
        raw, act = x_raw, x_act
        a_down = None
        for i, bp in enumerate(blocks):
            last = i == len(blocks) - 1
            if last:
                raw, (a_down, _) = self._run_block(bp, raw, act, lv, out_raw=cat_raw[:, :c],
                                                   out_acts=[(None, plan["down_bn"]), (cat_act[:, :c], (t_sc[:c], t_sh[:c]))])
            else:
                raw, (act,) = self._run_block(bp, raw, act, lv, out_acts=[(None, blocks[i + 1]["bn0"])])
        nxt = pyr.levels[l + 1]
        sub_bn0 = self.u._get_plan()["blocks"][0]["bn0"]
        d_raw, (d_act,) = self._conv(a_down, plan["down_w"], lv.child, lv.child_mask, nxt.n, acts=[(None, sub_bn0)])
        _, u_act = self.u._forward_level(d_raw, d_act, pyr, l + 1, outputs, out_bn=plan["up_bn"])
        self._conv(u_act, plan["up_w"], lv.up, lv.up_mask, lv.n, raw=cat_raw[:, c:],
                   acts=[(cat_act[:, c:], (t_sc[c:], t_sh[c:]))])
        # tail block 0: SubM1(cat) + SubM3(BN.SubM3(BN.cat))   (in 2c -> c, spconv_unet.py:36-38)
        r = ops.gemm(cat_raw, tail[0]["wi"])
        tb, tm, pm = lv.subm_conv
        _, (y_act,) = self._conv(cat_act, tail[0]["w0"], tb, tm, lv.n, want_raw=False, acts=[(None, tail[0]["bn1"])], perm=pm)
        raw, a = self._conv(y_act, tail[0]["w1"], tb, tm, lv.n, residual=r,
                            acts=[(None, tail[1]["bn0"])] if len(tail) > 1 else out_acts_final, perm=pm)
        act = a[0] if a else None
        for i in range(1, len(tail)):
            last = i == len(tail) - 1
            acts = out_acts_final if last else [(None, tail[i + 1]["bn0"])]
            raw, a = self._run_block(tail[i], raw, act, lv, out_acts=acts)
            act = a[0] if a else None
        outputs.append((l, raw))
        return raw, act

    # ---- fallback executor: fp32 feature maps, BN+ReLU folded into every consumer's operand load
    #      (used when a level's channel count is not a multiple of 32, e.g. tiny test models)
    @staticmethod
    def _run_block_fp32(bp, x, lv, out=None):
        identity = ops.gemm(x, bp["wi"]) if "wi" in bp else x
        y = ops.gemm(x, bp["w0"], table=lv.subm, tile_mask=lv.subm_mask, in_scale=bp["bn0"][0], in_shift=bp["bn0"][1],
                     in_relu=True)
        return ops.gemm(y, bp["w1"], table=lv.subm, tile_mask=lv.subm_mask, in_scale=bp["bn1"][0], in_shift=bp["bn1"][1],
                        in_relu=True, residual=identity, out=out)

    def _forward_level_fp32(self, x: torch.Tensor, pyr: Pyramid, l: int, outputs: list) -> torch.Tensor:
        plan = self._get_plan()
        lv = pyr.levels[l]
        c = self.num_planes[0]
        has_sub = len(self.num_planes) > 1
        cat = torch.empty((lv.n, 2 * c), dtype=torch.float32, device=x.device) if has_sub else None
        nb = len(plan["blocks"])
        for i, bp in enumerate(plan["blocks"]):
            dst = cat[:, :c] if (has_sub and i == nb - 1) else None
            x = self._run_block_fp32(bp, x, lv, out=dst)
        if has_sub:
            nxt = pyr.levels[l + 1]
            d = ops.gemm(x, plan["down_w"], table=lv.child, tile_mask=lv.child_mask, n_out=nxt.n,
                         in_scale=plan["down_bn"][0], in_shift=plan["down_bn"][1], in_relu=True)
            d = self.u._forward_level_fp32(d, pyr, l + 1, outputs)
            ops.gemm(d, plan["up_w"], table=lv.up, tile_mask=lv.up_mask, n_out=lv.n, in_scale=plan["up_bn"][0],
                     in_shift=plan["up_bn"][1], in_relu=True, out=cat[:, c:])
            x = cat
            for bp in plan["tail"]:
                x = self._run_block_fp32(bp, x, lv)
        outputs.append((l, x))
        return x

    # ---- train-mode executor: BatchNorm with BATCH statistics (spconv_unet.py:119-124: SyncBatchNorm over all active
    #      voxels of the global batch; running statistics updated with momentum 0.1).  A BatchNorm's statistics need the
    #      producer's complete output, so the producer-epilogue fusion of the eval executor does not apply: each conv
    #      writes its fp32 map, ud3d_bn_batch_sums reduces it (+ one all-reduce under torch.distributed), and the
    #      normalisation + ReLU is folded into the CONSUMER's operand load.  Forward values only here -- the same op sequence
    #      with a tape and the complete backward pass is unidet3d_b200/train.py (unet_forward / backbone_backward).
    @staticmethod
    def _block_train(blk: ResidualBlock, x, lv):
        cb = blk.conv_branch
        identity = ops.gemm(x, ops.PackedWeight(blk.i_branch[0].weight)) if isinstance(blk.i_branch[0], SparseConvWeight) else x
        s0, h0, _, _ = ops.bn_train(x, cb[0])
        y = ops.gemm(x, ops.PackedWeight(cb[2].weight), table=lv.subm, tile_mask=lv.subm_mask, in_scale=s0, in_shift=h0, in_relu=True)
        s1, h1, _, _ = ops.bn_train(y, cb[3])
        return ops.gemm(y, ops.PackedWeight(cb[5].weight), table=lv.subm, tile_mask=lv.subm_mask, in_scale=s1, in_shift=h1,
                        in_relu=True, residual=identity)

    def _forward_level_train(self, x: torch.Tensor, pyr: Pyramid, l: int, outputs: list) -> torch.Tensor:
        lv = pyr.levels[l]
        c = self.num_planes[0]
        has_sub = len(self.num_planes) > 1
        for blk in self.blocks:
            x = self._block_train(blk, x, lv)
        if has_sub:
            nxt = pyr.levels[l + 1]
            cat = torch.empty((lv.n, 2 * c), dtype=torch.float32, device=x.device)
            cat[:, :c] = x
            s, h, _, _ = ops.bn_train(x, self.conv[0])
            d = ops.gemm(x, ops.PackedWeight(self.conv[2].weight), table=lv.child, tile_mask=lv.child_mask, n_out=nxt.n,
                         in_scale=s, in_shift=h, in_relu=True)
            d = self.u._forward_level_train(d, pyr, l + 1, outputs)
            s, h, _, _ = ops.bn_train(d, self.deconv[0])
            ops.gemm(d, ops.PackedWeight(self.deconv[2].weight), table=lv.up, tile_mask=lv.up_mask, n_out=lv.n, in_scale=s,
                     in_shift=h, in_relu=True, out=cat[:, c:])
            x = cat
            for blk in self.blocks_tail:
                x = self._block_train(blk, x, lv)
        outputs.append((l, x))
        return x

    def operand_form_ok(self):
        return all(c % 32 == 0 for c in self.num_planes)

    def n_levels(self):
        return len(self.num_planes)

    def forward(self, input: SparseConvTensor, previous_outputs: Optional[List] = None):
        if input.pyramid is None or len(input.pyramid) < self.n_levels():
            input.pyramid = build_pyramid(input.indices, input.spatial_shape, input.batch_size, self.n_levels(),
                                          canonical=input.canonical, extents=input.extents)
        pyr = input.pyramid
        x = input.features
        if not (x.is_cuda and x.dtype == torch.float32):
            raise RuntimeError("SpConvUNet needs CUDA fp32 features (no CPU fallback)")
        outs: list = []
        x = x.contiguous() if x.stride(1) != 1 else x
        if self.training:
            y = self._forward_level_train(x, pyr, 0, outs)
        elif self.operand_form_ok():
            x_act = getattr(input, "features_act", None)   # operand form under blocks.block0.bn0, if the caller has it
            if x_act is None:
                bn0 = self._get_plan()["blocks"][0]["bn0"]
                x_act = ops.act_split(x, bn0[0], bn0[1], relu=True)
            if self.use_stage_plan:
                y = self._forward_plan(x, x_act, pyr, outs, self.return_blocks)
            else:
                y, _ = self._forward_level(x, x_act, pyr, 0, outs)
        else:
            y = self._forward_level_fp32(x, pyr, 0, outs)
        output = input.replace_feature(y)
        if self.return_blocks:
            if previous_outputs is None:
                previous_outputs = []
            for l, f in outs:       # deepest level first, like the reference's recursion (spconv_unet.py:233-238)
                lv = pyr.levels[l]
                previous_outputs.append(output if l == 0 else SparseConvTensor(f, lv.coords, lv.shape, input.batch_size))
            return output, previous_outputs
        return output
