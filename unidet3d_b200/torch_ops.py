"""``torch.library`` registration of the tensor-in / tensor-out entry points ("called from Python via ctypes /
torch.library custom ops"), namespace ``unidet3d_b200``:

    torch.ops.unidet3d_b200.gather_gemm   ud3d_gemm_fwd       (sparse conv / linear on a packed weight image)
    torch.ops.unidet3d_b200.act_split     ud3d_act_split      (fp32 -> operand form under a folded BatchNorm + ReLU)
    torch.ops.unidet3d_b200.segmented_mean  ud3d_segmented_mean
    torch.ops.unidet3d_b200.layernorm     ud3d_layernorm
    torch.ops.unidet3d_b200.attention     ud3d_attention_fwd_tc (operand-form q|k|v in, operand-form context out)
    torch.ops.unidet3d_b200.voxel_mean    ud3d_voxel_mean

Each op has a fake (meta) implementation, so FakeTensor tracing / ``torch.compile`` graphs and export see the output
shapes without running a kernel; the real implementations launch on the current stream and are CUDA-graph capturable
(no host synchronisation, workspaces allocated through the caching allocator).  They are thin wrappers over
``unidet3d_b200.ops`` -- the module classes call ``ops`` directly.  Autograd formulas are not registered: the backward
pass is the library's own tape (``unidet3d_b200.train``), not torch.autograd.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

_lib = torch.library.Library("unidet3d_b200", "DEF")          # keeps the definitions alive for the process lifetime


class _Packed:
    """View of a packed weight image (uint8 tensor) + its logical shape, in the duck type ``ops.gemm`` expects."""

    def __init__(self, data, K, c_in, c_out):
        self.data, self.data_ts, self.K, self.c_in, self.c_out = data, data, K, c_in, c_out


@torch.library.custom_op("unidet3d_b200::gather_gemm", mutates_args=())
def gather_gemm(x: torch.Tensor, w_packed: torch.Tensor, K: int, c_in: int, c_out: int, n_out: int,
                table: Optional[torch.Tensor] = None, tile_mask: Optional[torch.Tensor] = None,
                bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, act: str = "none",
                in_split: bool = False, in_scale: Optional[torch.Tensor] = None, in_shift: Optional[torch.Tensor] = None,
                in_relu: bool = False) -> torch.Tensor:
    return ops.gemm(x, _Packed(w_packed, K, c_in, c_out), table=table, tile_mask=tile_mask, n_out=n_out, bias=bias,
                    residual=residual, act=act, in_split=in_split, in_scale=in_scale, in_shift=in_shift, in_relu=in_relu)


@gather_gemm.register_fake
def _(x, w_packed, K, c_in, c_out, n_out, table=None, tile_mask=None, bias=None, residual=None, act="none", in_split=False,
      in_scale=None, in_shift=None, in_relu=False):
    return x.new_empty((n_out, c_out))


@torch.library.custom_op("unidet3d_b200::act_split", mutates_args=())
def act_split(raw: torch.Tensor, scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
              relu: bool = True) -> torch.Tensor:
    return ops.act_split(raw, scale, shift, relu=relu)


@act_split.register_fake
def _(raw, scale=None, shift=None, relu=True):
    return raw.new_empty((raw.shape[0], (raw.shape[1] + 31) // 32 * 32))


@torch.library.custom_op("unidet3d_b200::segmented_mean", mutates_args=())
def segmented_mean(src: torch.Tensor, seg: torch.Tensor, n_seg: int, gather: Optional[torch.Tensor] = None, channels: int = -1,
                   scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    return ops.segmented_mean(src, seg, n_seg, gather=gather, channels=None if channels < 0 else channels, scale=scale,
                              shift=shift, relu=relu)


@segmented_mean.register_fake
def _(src, seg, n_seg, gather=None, channels=-1, scale=None, shift=None, relu=False):
    return src.new_empty((n_seg, src.shape[1] if channels < 0 else channels))


@torch.library.custom_op("unidet3d_b200::layernorm", mutates_args=())
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, residual: Optional[torch.Tensor] = None,
              eps: float = 1e-5) -> torch.Tensor:
    return ops.layernorm(x, gamma, beta, residual=residual, eps=eps)


@layernorm.register_fake
def _(x, gamma, beta, residual=None, eps=1e-5):
    return torch.empty_like(x)


@torch.library.custom_op("unidet3d_b200::attention", mutates_args=())
def attention(qkv_split: torch.Tensor, cu_seqlens: torch.Tensor, max_T: int, num_heads: int) -> torch.Tensor:
    return ops.attention(qkv_split, cu_seqlens, max_T, num_heads, split_out=True, split_in=True, tcgen05=True)


@attention.register_fake
def _(qkv_split, cu_seqlens, max_T, num_heads):
    return qkv_split.new_empty((qkv_split.shape[0], qkv_split.shape[1] // 3))


@torch.library.custom_op("unidet3d_b200::voxel_mean", mutates_args=())
def voxel_mean(feats_pts: torch.Tensor, rank: torch.Tensor, n_vox: int) -> torch.Tensor:
    return ops.voxel_mean(feats_pts, rank, n_vox)


@voxel_mean.register_fake
def _(feats_pts, rank, n_vox):
    return feats_pts.new_empty((n_vox, feats_pts.shape[1]))
