"""Thin torch-tensor wrappers over the C-ABI (device memory + streams come from PyTorch; every
kernel is ours).  All functions require CUDA tensors and raise otherwise -- there is no CPU path."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import CriterionArgs, GemmArgs, check

TILE_M = 128


def _L():
    return _lib.load()


_raw_stream = torch._C._cuda_getCurrentRawStream
_cur_device = torch._C._cuda_getDevice


def _stream():
    # raw cudaStream_t of torch's current stream: the two C calls cost ~0.3 us; torch.cuda.current_stream() builds a
    # Stream object through several Python layers (~4 us, 170 calls per forward step)
    return C.c_void_p(_raw_stream(_cur_device()))


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.Ud3dError(f"{name}: expected a CUDA tensor (unidet3d_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.Ud3dError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.Ud3dError(f"{name}: expected a contiguous tensor")
    return t


def _dims(d: Sequence[int]):
    return (C.c_int32 * len(d))(*[int(x) for x in d])


def launch_count(reset: bool = False) -> int:
    return int(_L().ud3d_launch_count(1 if reset else 0))


# ------------------------------------------------------------------ voxelisation / grid
def point_coords(points: torch.Tensor, scene_offsets: torch.Tensor, voxel_size: float):
    """points fp32 [n,6], scene_offsets int32 [B+1] -> coords int32 [n,4], feats [n,6], stats [B,6], max_coord int32 [3]."""
    _req(points, torch.float32, "points"), _req(scene_offsets, torch.int32, "scene_offsets")
    n, B = points.shape[0], scene_offsets.numel() - 1
    dev = points.device
    coords = torch.empty((n, 4), dtype=torch.int32, device=dev)
    feats = torch.empty((n, 6), dtype=torch.float32, device=dev)
    stats = torch.empty((B, 6), dtype=torch.float32, device=dev)
    maxc = torch.empty(3, dtype=torch.int32, device=dev)
    wsb = _L().ud3d_point_coords_workspace_bytes(B)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(_L().ud3d_point_coords(_p(points), n, _p(scene_offsets), B, float(voxel_size), _p(coords), _p(feats),
                                 _p(stats), _p(maxc), _p(ws), wsb, _stream()), "ud3d_point_coords")
    return coords, feats, stats, maxc


class Grid:
    """Occupancy grid over the dense (B,X,Y,Z) box: bitmap + popcount prefix (see the C header)."""

    def __init__(self, dims: Sequence[int], device):
        self.dims = [int(d) for d in dims]
        self._cd = _dims(self.dims)
        self.ws_bytes = int(_L().ud3d_grid_workspace_bytes(self._cd))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=device)
        self.n_unique = torch.zeros(1, dtype=torch.int32, device=device)

    def build(self, coords: torch.Tensor):
        _req(coords, torch.int32, "coords")
        check(_L().ud3d_grid_build(_p(coords), coords.shape[0], self._cd, _p(self.ws), self.ws_bytes,
                                   _p(self.n_unique), _stream()), "ud3d_grid_build")
        return self.n_unique

    def build_from_finer(self, fine: "Grid"):
        """Occupancy of the k=2,s=2 down-sampled level from the finer grid's bitmap (ud3d_grid_build_coarser)."""
        check(_L().ud3d_grid_build_coarser(fine._cd, _p(fine.ws), self._cd, _p(self.ws), self.ws_bytes, _p(self.n_unique),
                                           _stream()), "ud3d_grid_build_coarser")
        return self.n_unique

    def rank(self, coords: torch.Tensor) -> torch.Tensor:
        _req(coords, torch.int32, "coords")
        out = torch.empty(coords.shape[0], dtype=torch.int32, device=coords.device)
        check(_L().ud3d_grid_rank(_p(coords), coords.shape[0], self._cd, _p(self.ws), _p(out), _stream()), "ud3d_grid_rank")
        return out

    def coords(self, n_unique: int) -> torch.Tensor:
        out = torch.empty((n_unique, 4), dtype=torch.int32, device=self.ws.device)
        check(_L().ud3d_grid_coords(self._cd, _p(self.ws), _p(out), n_unique, _stream()), "ud3d_grid_coords")
        return out


def voxel_mean(feats_pts: torch.Tensor, rank: torch.Tensor, n_vox: int) -> torch.Tensor:
    _req(feats_pts, torch.float32, "feats_pts"), _req(rank, torch.int32, "rank")
    n, Cc = feats_pts.shape
    out = torch.empty((n_vox, Cc), dtype=torch.float32, device=feats_pts.device)
    ws = torch.empty(max(n_vox, 1) * 4, dtype=torch.uint8, device=feats_pts.device)
    check(_L().ud3d_voxel_mean(_p(feats_pts), _p(rank), n, Cc, n_vox, _p(out), _p(ws), ws.numel(), _stream()), "ud3d_voxel_mean")
    return out


def rulebook_subm3(coords: torch.Tensor, grid: Grid, canonical: bool, with_mask: bool = True):
    """-> table int32 [27,N], tile_mask uint32-as-int32 [ceil(N/128)] (or None)."""
    _req(coords, torch.int32, "coords")
    n = coords.shape[0]
    dev = coords.device
    table = torch.empty((27, n), dtype=torch.int32, device=dev)
    mask = torch.empty((n + TILE_M - 1) // TILE_M, dtype=torch.int32, device=dev) if with_mask else None
    ror = None if canonical else torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    check(_L().ud3d_rulebook_subm3(_p(coords), n, grid._cd, _p(grid.ws), _p(ror), _p(table), _p(mask), _stream()),
          "ud3d_rulebook_subm3")
    return table, mask


def subm3_tile_order(table: torch.Tensor):
    """-> (perm int32 [n], table_p int32 [27,n], tile_mask_p [ceil(n/128)]): rows regrouped by neighbourhood pattern
    (see ud3d_subm3_tile_order)."""
    _req(table, torch.int32, "table")
    n = table.shape[1]
    dev = table.device
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    table_p = torch.empty_like(table)
    mask_p = torch.empty((n + TILE_M - 1) // TILE_M, dtype=torch.int32, device=dev)
    ws = torch.empty(int(_L().ud3d_subm3_tile_order_workspace_bytes(n)), dtype=torch.uint8, device=dev)
    check(_L().ud3d_subm3_tile_order(_p(table), n, _p(perm), _p(table_p), _p(mask_p), _p(ws), ws.numel(), _stream()),
          "ud3d_subm3_tile_order")
    return perm, table_p, mask_p


def down2_parents(coords: torch.Tensor, in_shape: Sequence[int]) -> torch.Tensor:
    _req(coords, torch.int32, "coords")
    parents = torch.empty_like(coords)
    check(_L().ud3d_down2_parents(_p(coords), coords.shape[0], _dims(in_shape), _p(parents), _stream()), "ud3d_down2_parents")
    return parents


def down_ancestors(coords: torch.Tensor, in_shape: Sequence[int], levels: int) -> torch.Tensor:
    """coords after ``levels`` k2/s2 down-samplings (b = -1 when dropped on the way); see ud3d_down_ancestors."""
    _req(coords, torch.int32, "coords")
    out = torch.empty_like(coords)
    check(_L().ud3d_down_ancestors(_p(coords), coords.shape[0], _dims(in_shape), int(levels), _p(out), _stream()),
          "ud3d_down_ancestors")
    return out


def rulebook_down2(coords: torch.Tensor, parents: torch.Tensor, n_coarse: int, coarse_grid: Grid, with_mask: bool = True):
    n_fine = coords.shape[0]
    dev = coords.device
    child = torch.empty((8, n_coarse), dtype=torch.int32, device=dev)
    up = torch.empty((8, n_fine), dtype=torch.int32, device=dev)
    cm = torch.empty((n_coarse + TILE_M - 1) // TILE_M, dtype=torch.int32, device=dev) if with_mask else None
    um = torch.empty((n_fine + TILE_M - 1) // TILE_M, dtype=torch.int32, device=dev) if with_mask else None
    check(_L().ud3d_rulebook_down2(_p(coords), _p(parents), n_fine, n_coarse, coarse_grid._cd, _p(coarse_grid.ws),
                                   _p(child), _p(up), _p(cm), _p(um), _stream()), "ud3d_rulebook_down2")
    return child, up, cm, um


# ------------------------------------------------------------------ gather-GEMM
class PackedWeight:
    """Weight in the kernel's shared-memory image.  ``w``: [C_out, K, C_in] fp32 (reference layout
    flattened: spconv [C_out,k,k,k,C_in]; nn.Linear [C_out,C_in] with K=1)."""

    def __init__(self, w: torch.Tensor, ts: bool = True):
        if w.dim() == 2:
            w = w.unsqueeze(1)
        if w.dim() == 5:
            w = w.reshape(w.shape[0], -1, w.shape[-1])
        w = _req(w.detach().to(torch.float32).contiguous(), torch.float32, "weight")
        self.c_out, self.K, self.c_in = w.shape
        nbytes = int(_L().ud3d_gemm_packed_weight_bytes(self.K, self.c_in, self.c_out))
        self.data = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        check(_L().ud3d_gemm_pack_weight(_p(w), self.K, self.c_in, self.c_out, _p(self.data), _stream()), "ud3d_gemm_pack_weight")
        # second image in the K order of the kernel variant whose A operand goes through registers into TMEM (``ts``: the
        # training step re-packs every weight every step and never uses that variant)
        self.data_ts = None
        if not ts:
            return
        self.data_ts = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        check(_L().ud3d_gemm_pack_weight_ts(_p(w), self.K, self.c_in, self.c_out, _p(self.data_ts), _stream()),
              "ud3d_gemm_pack_weight_ts")


def _gemm_args(x, pw_K, c_in, c_out, n_out, table, tile_mask, out, in_scale, in_shift, in_relu, bias, act, residual,
               w_packed_ptr, in_split=False, no_raw=False, acts=None, row_perm=None, w_packed_ts_ptr=None):
    a = GemmArgs()
    a.in_ = x.data_ptr(); a.ld_in = x.stride(0); a.c_in = c_in
    a.table = table.data_ptr() if table is not None else None
    a.tile_mask = tile_mask.data_ptr() if tile_mask is not None else None
    a.K = pw_K; a.n_out = n_out
    a.w_packed = w_packed_ptr
    a.out = out.data_ptr(); a.ld_out = out.stride(0); a.c_out = c_out
    a.in_scale = in_scale.data_ptr() if in_scale is not None else None
    a.in_shift = in_shift.data_ptr() if in_shift is not None else None
    a.in_relu = 1 if in_relu else 0
    a.bias = bias.data_ptr() if bias is not None else None
    a.act = {None: 0, "none": 0, "relu": 1, "gelu": 2}[act]
    a.residual = residual.data_ptr() if residual is not None else None
    a.ld_res = residual.stride(0) if residual is not None else 0
    a.in_split = int(in_split)
    a.no_raw = 1 if no_raw else 0
    norelu = 0
    for i, spec in enumerate(acts or ()):
        buf, sc, sh = spec[0], spec[1], spec[2]
        a.out_act[i] = buf.data_ptr(); a.ld_act[i] = buf.stride(0)
        a.act_scale[i] = sc.data_ptr() if sc is not None else None
        a.act_shift[i] = sh.data_ptr() if sh is not None else None
        if len(spec) > 3 and not spec[3]:
            norelu |= 1 << i
    a.act_norelu = norelu
    a.row_perm = row_perm.data_ptr() if row_perm is not None else None
    a.w_packed_ts = w_packed_ts_ptr
    return a


def gemm(x: torch.Tensor, w: PackedWeight, *, table: Optional[torch.Tensor] = None, tile_mask=None,
         n_out: Optional[int] = None, out: Optional[torch.Tensor] = None, in_scale=None, in_shift=None,
         in_relu: bool = False, bias=None, act=None, residual=None, in_split: bool = False, no_raw: bool = False,
         acts=None, row_perm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = act(sum_k pre(x[table[k]]) @ W_k + bias) + residual   (see ud3d_gemm_fwd).
    ``x``/``out``/``residual`` may be column slices of wider row-major buffers (stride(1) == 1).
    ``in_split``: x is an operand-form (pre-activated, bf16 hi|lo) feature map, see ``act_split``.
    ``acts``: up to two (buffer, scale, shift[, relu=True]): also store relu?(out*scale+shift) in operand
    form (scale/shift None = identity).
    ``no_raw``: the fp32 result itself is not needed (``out`` is then scratch).
    ``row_perm``: int32 [n_out]; ``table`` / ``tile_mask`` are in the regrouped row order of ``subm3_tile_order`` and
    position i produces output row row_perm[i] (same results, fewer active offsets per tile)."""
    if not x.is_cuda or x.dtype != torch.float32 or x.stride(1) != 1:
        raise _lib.Ud3dError("gemm: x must be a CUDA fp32 matrix with unit column stride")
    if n_out is None:
        n_out = table.shape[1] if table is not None else x.shape[0]
    if out is None:
        out = torch.empty((n_out, w.c_out), dtype=torch.float32, device=x.device)
    c_in = w.c_in if not in_split else (w.c_in + 31) // 32 * 32     # operand form pads the last chunk with zeros
    assert out.stride(1) == 1 and out.shape[1] == w.c_out and x.shape[1] == c_in
    a = _gemm_args(x, w.K, c_in, w.c_out, n_out, table, tile_mask, out, in_scale, in_shift, in_relu, bias, act,
                   residual, w.data.data_ptr(), in_split, no_raw, acts, row_perm, w.data_ts.data_ptr() if w.data_ts is not None else None)
    check(_L().ud3d_gemm_fwd(C.byref(a), _stream()), "ud3d_gemm_fwd")
    return out


def operand_form_interleave(x_split: torch.Tensor) -> torch.Tensor:
    """operand form -> interleaved operand form (``gemm(..., in_split=2)``, the experimental TMEM-operand kernel): per
    32-channel chunk the eight 16-byte pieces [h0 h1 h2 h3 l0 l1 l2 l3] become [h0 l0 h1 l1 h2 l2 h3 l3]."""
    n, c = x_split.shape
    v = x_split.contiguous().view(n, c // 32, 2, 4, 4)             # [row, chunk, hi|lo, quarter, 4 words]
    return v.permute(0, 1, 3, 2, 4).contiguous().view(n, c)


def act_split(raw: torch.Tensor, scale=None, shift=None, relu: bool = True, out: Optional[torch.Tensor] = None):
    """fp32 [n,c] -> operand form [n,c] (same bytes; bf16 hi|lo per 32-channel chunk) of relu(raw*scale+shift)."""
    if not raw.is_cuda or raw.dtype != torch.float32 or raw.stride(1) != 1:
        raise _lib.Ud3dError("act_split: raw must be a CUDA fp32 matrix with unit column stride")
    n, c = raw.shape
    if out is None:
        out = torch.empty((n, (c + 31) // 32 * 32), dtype=torch.float32, device=raw.device)
    check(_L().ud3d_act_split(_p(raw), raw.stride(0), n, c, _p(scale), _p(shift), 1 if relu else 0, _p(out), out.stride(0),
                              _stream()), "ud3d_act_split")
    return out


def gemm_simt(x, w_raw: torch.Tensor, *, table=None, n_out=None, in_scale=None, in_shift=None, in_relu=False,
              bias=None, act=None, residual=None) -> torch.Tensor:
    """Diagnostic fp32 CUDA-core version on the unpacked weight [C_out,K,C_in]."""
    w_raw = w_raw.reshape(w_raw.shape[0], -1, w_raw.shape[-1]).contiguous() if w_raw.dim() != 3 else w_raw.contiguous()
    c_out, K, c_in = w_raw.shape
    if n_out is None:
        n_out = table.shape[1] if table is not None else x.shape[0]
    out = torch.empty((n_out, c_out), dtype=torch.float32, device=x.device)
    a = _gemm_args(x, K, c_in, c_out, n_out, table, None, out, in_scale, in_shift, in_relu, bias, act, residual, None)
    check(_L().ud3d_gemm_fwd_simt(C.byref(a), _p(w_raw), _stream()), "ud3d_gemm_fwd_simt")
    return out


# ------------------------------------------------------------------ pooling / encoder pieces
def segmented_mean(src: torch.Tensor, seg: torch.Tensor, n_seg: int, *, gather: Optional[torch.Tensor] = None,
                   channels: Optional[int] = None, scale=None, shift=None, relu: bool = False) -> torch.Tensor:
    _req(seg, torch.int64, "seg")
    if not src.is_cuda or src.dtype != torch.float32 or src.stride(1) != 1:
        raise _lib.Ud3dError("segmented_mean: src must be a CUDA fp32 matrix with unit column stride")
    Cc = src.shape[1] if channels is None else channels
    n = seg.shape[0]
    out = torch.empty((n_seg, Cc), dtype=torch.float32, device=src.device)
    ws = torch.empty(max(int(_L().ud3d_segmented_mean_workspace_bytes(n_seg, Cc)), 8), dtype=torch.uint8, device=src.device)
    check(_L().ud3d_segmented_mean(_p(src), src.stride(0), Cc, _p(gather), _p(seg), n, n_seg, _p(scale), _p(shift),
                                   1 if relu else 0, _p(out), _p(ws), ws.numel(), _stream()), "ud3d_segmented_mean")
    return out


def layernorm(x: torch.Tensor, gamma, beta, residual=None, eps: float = 1e-5, out=None) -> torch.Tensor:
    _req(x, torch.float32, "x")
    if out is None:
        out = torch.empty_like(x)
    check(_L().ud3d_layernorm(_p(x), _p(residual), _p(gamma), _p(beta), _p(out), x.shape[0], x.shape[1], float(eps),
                              _stream()), "ud3d_layernorm")
    return out


def layernorm_split(x: torch.Tensor, gamma, beta, residual=None, eps: float = 1e-5, want_raw: bool = True):
    """-> (fp32 result or None, operand-form result)."""
    _req(x, torch.float32, "x")
    out = torch.empty_like(x) if want_raw else None
    out_s = torch.empty_like(x)
    check(_L().ud3d_layernorm_split(_p(x), _p(residual), _p(gamma), _p(beta), _p(out), _p(out_s), x.shape[0], x.shape[1],
                                    float(eps), _stream()), "ud3d_layernorm_split")
    return out, out_s


def attention(qkv: torch.Tensor, cu_seqlens: torch.Tensor, max_T: int, num_heads: int, split_out: bool = False,
              split_in: bool = False, tcgen05: bool = False) -> torch.Tensor:
    """softmax(QK^T/sqrt(32))V per scene; ``split_out``: result in operand form for the out-projection GEMM;
    ``split_in``: qkv itself is in operand form (implies split_out)."""
    _req(qkv, torch.float32, "qkv"), _req(cu_seqlens, torch.int32, "cu_seqlens")
    d = qkv.shape[1] // 3
    if d != num_heads * 32:
        raise _lib.Ud3dError("attention: head_dim must be 32")
    out = torch.empty((qkv.shape[0], d), dtype=torch.float32, device=qkv.device)
    if tcgen05 and not split_in:
        raise _lib.Ud3dError("attention: the tcgen05 kernel takes operand-form q|k|v (split_in=True)")
    if tcgen05:
        qkv = _req(qkv, torch.float32, "qkv")
        check(_L().ud3d_attention_fwd_tc(_p(qkv), _p(cu_seqlens), cu_seqlens.numel() - 1, int(max_T), int(qkv.shape[0]),
                                         num_heads, _p(out), _stream()), "ud3d_attention_fwd_tc")
        return out
    fn = _L().ud3d_attention_fwd_opform if split_in else (
        _L().ud3d_attention_fwd_split if split_out else _L().ud3d_attention_fwd)
    check(fn(_p(qkv), _p(cu_seqlens), cu_seqlens.numel() - 1, int(max_T), num_heads, _p(out), _stream()),
          "ud3d_attention_fwd")
    return out


def bbox_decode(raw: torch.Tensor, centers: torch.Tensor, with_angle: bool) -> torch.Tensor:
    _req(centers, torch.float32, "centers")
    T = raw.shape[0]
    out = torch.empty((T, 7 if with_angle else 6), dtype=torch.float32, device=raw.device)
    check(_L().ud3d_bbox_decode(_p(raw), raw.stride(0), _p(centers), T, 1 if with_angle else 0, _p(out), _stream()),
          "ud3d_bbox_decode")
    return out


def gather_columns(src: torch.Tensor, cols: torch.Tensor) -> torch.Tensor:
    _req(cols, torch.int32, "cols")
    T = src.shape[0]
    out = torch.empty((T, cols.numel()), dtype=torch.float32, device=src.device)
    check(_L().ud3d_gather_columns(_p(src), src.stride(0), _p(cols), cols.numel(), T, _p(out), _stream()), "ud3d_gather_columns")
    return out


# ------------------------------------------------------------------ post-processing
def topk_scores(logits: torch.Tensor, k: int):
    _req(logits, torch.float32, "logits")
    T, C1 = logits.shape
    dev = logits.device
    scores = torch.empty(k, dtype=torch.float32, device=dev)
    labels = torch.empty(k, dtype=torch.int32, device=dev)
    query = torch.empty(k, dtype=torch.int32, device=dev)
    ws = torch.empty(max(T * (C1 - 1), 1) * 4, dtype=torch.uint8, device=dev)
    check(_L().ud3d_topk_scores(_p(logits), T, C1, k, _p(scores), _p(labels), _p(query), _p(ws), ws.numel(), _stream()),
          "ud3d_topk_scores")
    return scores, labels, query


NMS_ROTATED_BEV, NMS_ALIGNED_BEV, NMS_ALIGNED_3D = 0, 1, 2


def nms_multiclass(boxes: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor, mode: int, iou_thr: float,
                   score_thr: float = 0.0):
    """-> keep int32 [n] (first n_keep valid), n_keep int32 [1] (device)."""
    _req(boxes, torch.float32, "boxes"), _req(scores, torch.float32, "scores"), _req(labels, torch.int32, "labels")
    n = boxes.shape[0]
    dev = boxes.device
    keep = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)   # tail beyond n_keep stays a valid index
    n_keep = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = int(_L().ud3d_nms_workspace_bytes(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(_L().ud3d_nms_multiclass(_p(boxes), boxes.shape[1], _p(scores), _p(labels), n, mode, float(iou_thr),
                                   float(score_thr), _p(keep), _p(n_keep), _p(ws), wsb, _stream()), "ud3d_nms_multiclass")
    return keep, n_keep


def trim_boxes(points: torch.Tensor, sp: torch.Tensor, n_sp: int, boxes: torch.Tensor, low_thr: float, up_thr: float,
               box_index: Optional[torch.Tensor] = None, m: Optional[int] = None,
               m_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(sp, torch.int64, "sp"), _req(boxes, torch.float32, "boxes")
    if not points.is_cuda or points.dtype != torch.float32 or points.stride(1) != 1:
        raise _lib.Ud3dError("trim_boxes: points must be a CUDA fp32 matrix with unit column stride")
    m = (box_index.numel() if box_index is not None else boxes.shape[0]) if m is None else m
    out = torch.empty((m, 6), dtype=torch.float32, device=boxes.device)
    wsb = int(_L().ud3d_trim_workspace_bytes(n_sp, points.shape[0], m))
    ws = torch.empty(wsb, dtype=torch.uint8, device=boxes.device)
    check(_L().ud3d_trim_boxes(_p(points), points.stride(0), _p(sp), points.shape[0], n_sp, _p(boxes), boxes.shape[1],
                               _p(box_index), m, _p(m_dev), float(low_thr), float(up_thr), _p(out), _p(ws), wsb, _stream()), "ud3d_trim_boxes")
    return out


def postprocess_scene(logits: torch.Tensor, boxes: torch.Tensor, k: int, nms_mode: int, iou_thr: float, score_thr: float,
                      points: Optional[torch.Tensor] = None, sp: Optional[torch.Tensor] = None, n_sp: int = 0,
                      low_thr: float = 0.0, up_thr: float = 1.0):
    """predict_by_feat for one scene in ONE host call (see ud3d_postprocess_scene).  ``points``/``sp`` given =>
    superpoint trimming.  Returns dict(scores, labels, cand, keep, n_keep, trimmed|None), all on the device."""
    from ._lib import PostArgs
    if not logits.is_cuda or logits.dtype != torch.float32 or logits.stride(1) != 1:
        raise _lib.Ud3dError("postprocess_scene: logits must be a CUDA fp32 matrix with unit column stride")
    _req(boxes, torch.float32, "boxes")
    dev = logits.device
    T, C1 = logits.shape
    use_trim = points is not None
    a = PostArgs()
    a.logits = logits.data_ptr(); a.ld_logits = logits.stride(0); a.T = T; a.C1 = C1
    a.boxes = boxes.data_ptr(); a.box_dim = boxes.shape[1]
    a.k = k; a.nms_mode = nms_mode; a.iou_thr = float(iou_thr); a.score_thr = float(score_thr)
    a.use_trim = 1 if use_trim else 0
    if use_trim:
        _req(sp, torch.int64, "sp")
        a.points = points.data_ptr(); a.ld_pts = points.stride(0); a.sp = sp.data_ptr()
        a.n_pts = points.shape[0]; a.n_sp = int(n_sp); a.low_thr = float(low_thr); a.up_thr = float(up_thr)
    # one allocation for all outputs: scores | labels | keep | n_keep | cand | trimmed
    bd = boxes.shape[1]
    buf = torch.empty(k * (3 + bd + 6) + 8, dtype=torch.float32, device=dev)
    scores = buf[:k]
    labels = buf[k:2 * k].view(torch.int32)
    keep = buf[2 * k:3 * k].view(torch.int32)
    n_keep = buf[3 * k:3 * k + 1].view(torch.int32)
    cand = buf[3 * k + 8:3 * k + 8 + k * bd].view(k, bd)
    trimmed = buf[3 * k + 8 + k * bd:].view(k, 6) if use_trim else None
    a.scores = scores.data_ptr(); a.labels = labels.data_ptr(); a.cand = cand.data_ptr(); a.keep = keep.data_ptr()
    a.n_keep = n_keep.data_ptr(); a.trimmed = trimmed.data_ptr() if use_trim else None
    wsb = int(_L().ud3d_postprocess_workspace_bytes(C.byref(a)))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(_L().ud3d_postprocess_scene(C.byref(a), _p(ws), wsb, _stream()), "ud3d_postprocess_scene")
    return dict(scores=scores, labels=labels, cand=cand, keep=keep, n_keep=n_keep, trimmed=trimmed, _buf=buf)


# ------------------------------------------------------------------ training-side targets / matcher / loss values
def boxes_by_instance(points: torch.Tensor, inst: torch.Tensor, n_inst: int) -> torch.Tensor:
    """get_bboxes_by_masks (unidet3d.py:220-275): points fp32 [n, >=3], inst int64 [n] (-1 = none) -> [n_inst, 6]."""
    _req(inst, torch.int64, "inst")
    if not points.is_cuda or points.dtype != torch.float32 or points.stride(1) != 1:
        raise _lib.Ud3dError("boxes_by_instance: points must be a CUDA fp32 matrix with unit column stride")
    out = torch.empty((n_inst, 6), dtype=torch.float32, device=points.device)
    ws = torch.empty(max(n_inst, 1) * 24, dtype=torch.uint8, device=points.device)
    check(_L().ud3d_boxes_by_instance(_p(points), points.stride(0), _p(inst), inst.numel(), int(n_inst), _p(out), _p(ws),
                                      ws.numel(), _stream()), "ud3d_boxes_by_instance")
    return out


def targets_by_distance(centers: torch.Tensor, gt_boxes: torch.Tensor, topk: int) -> torch.Tensor:
    """get_targets (unidet3d.py:371-409): centres [S,3], boxes [G, 6|7] -> bool [G, S]."""
    _req(centers, torch.float32, "centers"), _req(gt_boxes, torch.float32, "gt_boxes")
    S, G = centers.shape[0], gt_boxes.shape[0]
    masks = torch.zeros((G, S), dtype=torch.uint8, device=centers.device)
    ws = torch.empty(max(G, 1) * 4, dtype=torch.uint8, device=centers.device)
    check(_L().ud3d_targets_by_distance(_p(centers), S, _p(gt_boxes), gt_boxes.shape[1] if G else 6, G, int(topk),
                                        _p(masks), _p(ws), ws.numel(), _stream()), "ud3d_targets_by_distance")
    return masks.bool()


def criterion_layer(logits: torch.Tensor, boxes: torch.Tensor, gt_boxes: torch.Tensor, gt_labels: torch.Tensor,
                    query_masks: torch.Tensor, topk: int, w_cls: float, w_box: float, non_object_weight: float):
    """UniMatcher + the loss terms of one (layer, scene) (ud3d_criterion_layer).
    -> (match bool [T, G], sums fp32 [4] = CE numerator, CE denominator, box-loss sum, number of matched pairs)."""
    if not logits.is_cuda or logits.dtype != torch.float32 or logits.stride(1) != 1:
        raise _lib.Ud3dError("criterion_layer: logits must be a CUDA fp32 matrix with unit column stride")
    _req(boxes, torch.float32, "boxes")
    T, G = logits.shape[0], gt_labels.shape[0]
    dev = logits.device
    match = torch.empty((T, G), dtype=torch.uint8, device=dev)
    sums = torch.empty(4, dtype=torch.float32, device=dev)
    a = CriterionArgs()
    a.logits = logits.data_ptr(); a.ld_logits = logits.stride(0); a.T = T; a.C1 = logits.shape[1]
    a.boxes = boxes.data_ptr(); a.box_dim = boxes.shape[1]
    if G:
        _req(gt_boxes, torch.float32, "gt_boxes"), _req(gt_labels, torch.int64, "gt_labels")
        qm = query_masks if query_masks.dtype == torch.uint8 else query_masks.to(torch.uint8)
        qm = _req(qm.contiguous(), torch.uint8, "query_masks")
        if tuple(qm.shape) != (G, T) or gt_boxes.shape[1] != boxes.shape[1]:
            raise _lib.Ud3dError("criterion_layer: query_masks must be [G, T] and gt_boxes match the predicted box_dim")
        a.gt_boxes = gt_boxes.data_ptr(); a.gt_labels = gt_labels.data_ptr(); a.query_masks = qm.data_ptr()
    a.G = G
    a.topk = int(topk); a.w_cls = float(w_cls); a.w_box = float(w_box); a.non_object_weight = float(non_object_weight)
    a.match = match.data_ptr() if G else None
    a.sums = sums.data_ptr()
    ws = torch.empty(int(_L().ud3d_criterion_workspace_bytes(T, G)), dtype=torch.uint8, device=dev)
    check(_L().ud3d_criterion_layer(C.byref(a), _p(ws), ws.numel(), _stream()), "ud3d_criterion_layer")
    return match.bool(), sums


def criterion_layer_grad(logits: torch.Tensor, boxes: torch.Tensor, gt_boxes: torch.Tensor, gt_labels: torch.Tensor,
                         match: torch.Tensor, sums: torch.Tensor, scales: torch.Tensor, non_object_weight: float):
    """Gradients of one (layer, scene) of the criterion (ud3d_criterion_layer_grad) given ``match`` / ``sums`` of
    ``criterion_layer`` and the DEVICE scalars ``scales`` = (d loss / d CE term, d loss / d box term) of this scene.
    -> (d_logits [T, C+1], d_boxes [T, box_dim])."""
    if not logits.is_cuda or logits.dtype != torch.float32 or logits.stride(1) != 1:
        raise _lib.Ud3dError("criterion_layer_grad: logits must be a CUDA fp32 matrix with unit column stride")
    _req(boxes, torch.float32, "boxes"), _req(sums, torch.float32, "sums"), _req(scales, torch.float32, "scales")
    if sums.numel() != 4 or scales.numel() != 2:
        raise _lib.Ud3dError("criterion_layer_grad: sums must hold 4 and scales 2 floats")
    T, G = logits.shape[0], gt_labels.shape[0]
    d_logits = torch.empty((T, logits.shape[1]), dtype=torch.float32, device=logits.device)
    d_boxes = torch.empty_like(boxes)
    a = _lib.CriterionGradArgs()
    a.logits = logits.data_ptr(); a.ld_logits = logits.stride(0); a.T = T; a.C1 = logits.shape[1]
    a.boxes = boxes.data_ptr(); a.box_dim = boxes.shape[1]
    if G:
        _req(gt_boxes, torch.float32, "gt_boxes"), _req(gt_labels, torch.int64, "gt_labels")
        m = match.view(torch.uint8) if match.dtype == torch.bool else match
        m = _req(m, torch.uint8, "match")
        if tuple(m.shape) != (T, G) or gt_boxes.shape[1] != boxes.shape[1]:
            raise _lib.Ud3dError("criterion_layer_grad: match must be [T, G] and gt_boxes match the predicted box_dim")
        a.gt_boxes = gt_boxes.data_ptr(); a.gt_labels = gt_labels.data_ptr(); a.match = m.data_ptr()
    a.G = G
    a.sums = sums.data_ptr(); a.scales = scales.data_ptr(); a.non_object_weight = float(non_object_weight)
    a.d_logits = d_logits.data_ptr(); a.ld_dlogits = d_logits.stride(0); a.d_boxes = d_boxes.data_ptr()
    check(_L().ud3d_criterion_layer_grad(C.byref(a), _stream()), "ud3d_criterion_layer_grad")
    return d_logits, d_boxes


def head_backward(raw: torch.Tensor, d_box: Optional[torch.Tensor], with_angle: bool, d_cls: Optional[torch.Tensor],
                  cols: torch.Tensor, d_raw: torch.Tensor, d_logits: torch.Tensor):
    """Backward of one scene's head outputs (ud3d_head_backward): writes ``d_raw`` [T, 8] and ``d_logits`` [T, n_union]
    (row slices of the packed buffers) from the gradients of the decoded boxes and of the gathered class columns."""
    for t, nme in ((raw, "raw"), (d_raw, "d_raw"), (d_logits, "d_logits")):
        if not t.is_cuda or t.dtype != torch.float32 or t.stride(1) != 1:
            raise _lib.Ud3dError(f"head_backward: {nme} must be a CUDA fp32 matrix with unit column stride")
    T = raw.shape[0]
    if d_box is not None:
        _req(d_box, torch.float32, "d_box")
        if tuple(d_box.shape) != (T, 7 if with_angle else 6):
            raise _lib.Ud3dError("head_backward: d_box must be [T, 7] with an angle and [T, 6] without")
    _req(cols, torch.int32, "cols")
    if d_cls is not None:
        _req(d_cls, torch.float32, "d_cls")
        if tuple(d_cls.shape) != (T, cols.numel()):
            raise _lib.Ud3dError("head_backward: d_cls must be [T, len(cols)]")
    if d_raw.shape[0] != T or d_logits.shape[0] != T or d_raw.shape[1] != 8:
        raise _lib.Ud3dError("head_backward: d_raw must be [T, 8] and d_logits [T, n_union]")
    check(_L().ud3d_head_backward(_p(raw), raw.stride(0), _p(d_box), 1 if with_angle else 0, _p(d_cls), _p(cols), cols.numel(), T,
                                  _p(d_raw), d_raw.stride(0), _p(d_logits), d_logits.stride(0), d_logits.shape[1], _stream()),
          "ud3d_head_backward")


# ------------------------------------------------------------------ training side of the backbone
def bn_batch_sums(x: torch.Tensor) -> torch.Tensor:
    """-> fp64 [2, C]: per-channel sum and sum of squares over the rows of ``x`` (deterministic)."""
    if not x.is_cuda or x.dtype != torch.float32 or x.stride(1) != 1:
        raise _lib.Ud3dError("bn_batch_sums: x must be a CUDA fp32 matrix with unit column stride")
    n, c = x.shape
    sums = torch.empty((2, c), dtype=torch.float64, device=x.device)
    wsb = int(_L().ud3d_bn_batch_sums_workspace_bytes(n, c))
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=x.device)
    check(_L().ud3d_bn_batch_sums(_p(x), x.stride(0), n, c, _p(sums), _p(ws), ws.numel(), _stream()), "ud3d_bn_batch_sums")
    return sums


def bn_train_fold(sums: torch.Tensor, count: float, bn, update_running: bool = True, count_dev: Optional[torch.Tensor] = None):
    """Train-mode BatchNorm as (scale, shift) [+ (mean, invstd) for the backward pass]: ``bn`` is the nn.BatchNorm1d /
    SyncBatchNorm holding gamma, beta, eps, momentum and the running statistics (updated in place like torch)."""
    c = sums.shape[1]
    dev = sums.device
    scale = torch.empty(c, dtype=torch.float32, device=dev)
    shift = torch.empty(c, dtype=torch.float32, device=dev)
    mean = torch.empty(c, dtype=torch.float32, device=dev)
    invstd = torch.empty(c, dtype=torch.float32, device=dev)
    rm = bn.running_mean if update_running else None
    rv = bn.running_var if update_running else None
    mom = 0.1 if bn.momentum is None else float(bn.momentum)
    check(_L().ud3d_bn_train_fold(_p(sums), float(count), c, _p(bn.weight.detach()), _p(bn.bias.detach()), float(bn.eps), mom,
                                  _p(rm), _p(rv), _p(scale), _p(shift), _p(mean), _p(invstd), _p(count_dev), _stream()), "ud3d_bn_train_fold")
    if update_running and getattr(bn, "num_batches_tracked", None) is not None:
        bn.num_batches_tracked += 1
    return scale, shift, mean, invstd


def sync_bn_sums(sums: torch.Tensor, count: float, group=None, device_count: bool = False):
    """The SyncBatchNorm exchange (spconv_unet.py:119-121): ONE all-reduce of the 2C channel sums + the row count.  Pure
    torch.distributed plumbing (device-agnostic: the gloo test in tests/test_sharding_gloo.py drives it on CPU tensors).
    -> (sums, count), or with ``device_count`` (sums, local count, global count as a 1-element tensor or None when there is
    nothing to exchange): the kernels then read the count on the device and the host never waits for the collective."""
    import torch.distributed as dist
    count_dev = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        buf = torch.cat((sums.flatten(), sums.new_tensor([count])))
        dist.all_reduce(buf, group=group)
        sums = buf[:-1].view(2, -1).contiguous()
        if device_count:
            count_dev = buf[-1:]
        else:
            count = float(buf[-1].item())
    return (sums, count, count_dev) if device_count else (sums, count)


def bn_train(x: torch.Tensor, bn, group=None, update_running: bool = True):
    """Batch statistics of ``x`` [N, C] (all active voxels of the batch) -> (scale, shift, mean, invstd).  With an
    initialised torch.distributed process group of more than one rank the sums and the row count are all-reduced first
    (SyncBatchNorm, spconv_unet.py:119-121): ONE collective of 2C + 1 doubles per BatchNorm, no host synchronisation."""
    sums, count, count_dev = sync_bn_sums(bn_batch_sums(x), float(x.shape[0]), group, device_count=x.is_cuda)
    return bn_train_fold(sums, max(count, 1.0), bn, update_running, count_dev=count_dev)


def bn_relu_apply(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, relu: bool = True) -> torch.Tensor:
    """relu?(x * scale + shift) as an fp32 map (recomputed in the backward pass as the X operand of ``conv_wgrad``)."""
    if not x.is_cuda or x.dtype != torch.float32 or x.stride(1) != 1:
        raise _lib.Ud3dError("bn_relu_apply: x must be a CUDA fp32 matrix with unit column stride")
    n, c = x.shape
    out = torch.empty((n, c), dtype=torch.float32, device=x.device)
    check(_L().ud3d_bn_relu_apply(_p(x), x.stride(0), n, c, _p(scale), _p(shift), 1 if relu else 0, _p(out), c, _stream()),
          "ud3d_bn_relu_apply")
    return out


def bn_relu_backward(x: torch.Tensor, d_act: torch.Tensor, scale, shift, mean, invstd, relu: bool = True, group=None,
                     dx: Optional[torch.Tensor] = None, accumulate: bool = False):
    """Backward of a = relu?(batchnorm_train(x)) given d_act = dL/da -> (dx, dgamma, dbeta).  With a torch.distributed
    group of more than one rank the two per-channel sums are all-reduced (SyncBatchNorm backward: ONE collective of
    2C + 1 doubles)."""
    for t, nme in ((x, "x"), (d_act, "d_act")):
        if not t.is_cuda or t.dtype != torch.float32 or t.stride(1) != 1:
            raise _lib.Ud3dError(f"bn_relu_backward: {nme} must be a CUDA fp32 matrix with unit column stride")
    n, c = x.shape
    sums = torch.empty((2, c), dtype=torch.float64, device=x.device)
    wsb = int(_L().ud3d_bn_batch_sums_workspace_bytes(n, c))
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=x.device)
    check(_L().ud3d_bn_backward_sums(_p(x), x.stride(0), _p(d_act), d_act.stride(0), n, c, _p(scale), _p(shift), _p(mean),
                                     _p(invstd), 1 if relu else 0, _p(sums), _p(ws), ws.numel(), _stream()), "ud3d_bn_backward_sums")
    sums, count, count_dev = sync_bn_sums(sums, float(n), group, device_count=True)
    if dx is None:
        dx = torch.empty((n, c), dtype=torch.float32, device=x.device)
        accumulate = False
    check(_L().ud3d_bn_backward_apply(_p(x), x.stride(0), _p(d_act), d_act.stride(0), n, c, _p(scale), _p(shift), _p(mean),
                                      _p(invstd), 1 if relu else 0, _p(sums), float(max(count, 1.0)), _p(dx), dx.stride(0),
                                      1 if accumulate else 0, _p(count_dev), _stream()), "ud3d_bn_backward_apply")
    return dx, sums[1].float(), sums[0].float()


def segmented_mean_backward(d_pooled: torch.Tensor, seg: torch.Tensor, n_rows: int, gather: Optional[torch.Tensor] = None):
    """d_src [n_rows, C]: the gradient of ``segmented_mean`` (without its affine) w.r.t. its source rows."""
    _req(d_pooled, torch.float32, "d_pooled"), _req(seg, torch.int64, "seg")
    n_seg, c = d_pooled.shape
    out = torch.empty((n_rows, c), dtype=torch.float32, device=d_pooled.device)
    wsb = int(_L().ud3d_segmented_mean_backward_workspace_bytes(n_rows, n_seg, c))
    ws = torch.empty(wsb, dtype=torch.uint8, device=d_pooled.device)
    check(_L().ud3d_segmented_mean_backward(_p(d_pooled), c, _p(gather), _p(seg), seg.numel(), n_seg, n_rows, _p(out), _p(ws), wsb,
                                            _stream()), "ud3d_segmented_mean_backward")
    return out


ATTN_BWD_VARIANT = os.environ.get("UD3D_ATTN_BWD", "reg")      # measurement switch: "warp" = the first implementation


def attention_backward(qkv: torch.Tensor, cu_seqlens: torch.Tensor, num_heads: int, out: torch.Tensor, d_out: torch.Tensor,
                       variant: Optional[str] = None):
    """fp32 qkv [T, 3d], forward result ``out`` [T, d] and its gradient -> dqkv [T, 3d].  ``variant``: "reg" (thread per
    query / key, register-resident rows; the default) or "warp" (warp per token, the first implementation)."""
    for t, nme in ((qkv, "qkv"), (out, "out"), (d_out, "d_out")):
        _req(t, torch.float32, nme)
    _req(cu_seqlens, torch.int32, "cu_seqlens")
    T = qkv.shape[0]
    dqkv = torch.empty_like(qkv)
    wsb = int(_L().ud3d_attention_bwd_workspace_bytes(T, num_heads))
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=qkv.device)
    fn = {"reg": _L().ud3d_attention_bwd_reg, "warp": _L().ud3d_attention_bwd}[variant or ATTN_BWD_VARIANT]
    check(fn(_p(qkv), _p(cu_seqlens), cu_seqlens.numel() - 1, T, num_heads, _p(out), _p(d_out), _p(dqkv), _p(ws), ws.numel(), _stream()),
          "ud3d_attention_bwd")
    return dqkv


def layernorm_backward(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5):
    """x = the LayerNorm's input (residual already added) -> (dx, dgamma, dbeta)."""
    _req(x, torch.float32, "x"), _req(dy, torch.float32, "dy")
    rows, c = x.shape
    dx = torch.empty_like(x)
    sums = torch.empty((2, c), dtype=torch.float64, device=x.device)
    wsb = int(_L().ud3d_layernorm_backward_workspace_bytes(rows, c))
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=x.device)
    check(_L().ud3d_layernorm_backward(_p(x), _p(dy), _p(gamma), rows, c, float(eps), _p(dx), _p(sums), _p(ws), ws.numel(), _stream()),
          "ud3d_layernorm_backward")
    return dx, sums[0].float(), sums[1].float()


def activation_forward(pre: torch.Tensor, act: str) -> torch.Tensor:
    _req(pre, torch.float32, "pre")
    out = torch.empty_like(pre)
    check(_L().ud3d_activation_forward(_p(pre), pre.numel(), {"relu": 1, "gelu": 2}[act], _p(out), _stream()), "ud3d_activation_forward")
    return out


def activation_backward(pre: torch.Tensor, dy: torch.Tensor, act: str) -> torch.Tensor:
    """dy * act'(pre) for act in {"relu", "gelu"} (pre = the pre-activation values)."""
    _req(pre, torch.float32, "pre"), _req(dy, torch.float32, "dy")
    dx = torch.empty_like(pre)
    check(_L().ud3d_activation_backward(_p(pre), _p(dy), pre.numel(), {"relu": 1, "gelu": 2}[act], _p(dx), _stream()),
          "ud3d_activation_backward")
    return dx


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, K: int, table: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """dW [C_out, K, C_in] (+)= sum_o dy[o]^T x[table[k][o]]   (``x`` = the conv's input after its BatchNorm + ReLU)."""
    for t, nme in ((x, "x"), (dy, "dy")):
        if not t.is_cuda or t.dtype != torch.float32 or t.stride(1) != 1:
            raise _lib.Ud3dError(f"conv_wgrad: {nme} must be a CUDA fp32 matrix with unit column stride")
    n_out, c_out = dy.shape
    c_in = x.shape[1]
    if table is not None:
        _req(table, torch.int32, "table")
    if out is None:
        out = torch.empty((c_out, K, c_in), dtype=torch.float32, device=x.device)
        accumulate = False
    wsb = int(_L().ud3d_conv_wgrad_workspace_bytes(n_out, K, c_in, c_out))
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=x.device)
    check(_L().ud3d_conv_wgrad(_p(x), x.stride(0), c_in, _p(dy), dy.stride(0), c_out, _p(table), n_out, K, _p(out),
                               1 if accumulate else 0, _p(ws), ws.numel(), _stream()), "ud3d_conv_wgrad")
    return out


def conv_dgrad(dy: torch.Tensor, weight: torch.Tensor, table_t: Optional[torch.Tensor], n_in: int, *, reverse_offsets: bool,
               tile_mask_t=None) -> torch.Tensor:
    """dX [n_in, C_in] = sum_k dy[table_t[k][i]] @ W_k^T: the input gradient of a sparse conv IS a sparse conv of dy with
    the transposed weight over the transposed rulebook -- for SubM3 the same table with the kernel offsets reversed
    (``reverse_offsets``), for the k2/s2 conv the table of its inverse conv and vice versa -- on the same tcgen05 kernel.
    ``weight``: the reference parameter [C_out, K, C_in] (flattened)."""
    w = weight.reshape(weight.shape[0], -1, weight.shape[-1])
    wt = w.permute(2, 1, 0)                                   # [C_in, K, C_out]
    if reverse_offsets:
        wt = wt.flip(1)
    return gemm(dy, PackedWeight(wt.contiguous(), ts=False), table=table_t, tile_mask=tile_mask_t, n_out=n_in)
