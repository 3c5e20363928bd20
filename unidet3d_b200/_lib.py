"""ctypes binding of libunidet3d_b200.so (the C-ABI in include/unidet3d_b200.h).

There is NO fallback: if the shared library is missing or fails to load, importing the
compute modules raises.  Build it with ``python -m unidet3d_b200.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libunidet3d_b200.so")

c_i32p = C.POINTER(C.c_int32)


class GemmArgs(C.Structure):
    """Mirror of ``ud3d_gemm_args`` (include/unidet3d_b200.h)."""
    _fields_ = [
        ("in_", C.c_void_p), ("ld_in", C.c_int32), ("c_in", C.c_int32),
        ("table", C.c_void_p), ("tile_mask", C.c_void_p),
        ("K", C.c_int32), ("n_out", C.c_int32),
        ("w_packed", C.c_void_p),
        ("out", C.c_void_p), ("ld_out", C.c_int32), ("c_out", C.c_int32),
        ("in_scale", C.c_void_p), ("in_shift", C.c_void_p), ("in_relu", C.c_int32),
        ("bias", C.c_void_p),
        ("act", C.c_int32),
        ("residual", C.c_void_p), ("ld_res", C.c_int32),
        ("in_split", C.c_int32), ("no_raw", C.c_int32),
        ("out_act", C.c_void_p * 2), ("ld_act", C.c_int32 * 2),
        ("act_scale", C.c_void_p * 2), ("act_shift", C.c_void_p * 2),
        ("act_norelu", C.c_int32),
        ("row_perm", C.c_void_p),
        ("w_packed_ts", C.c_void_p),
    ]


class UnetBlock(C.Structure):
    """Mirror of ``ud3d_unet_block``."""
    _fields_ = [("w0", C.c_void_p), ("w1", C.c_void_p), ("wi", C.c_void_p),
                ("bn0_scale", C.c_void_p), ("bn0_shift", C.c_void_p), ("bn1_scale", C.c_void_p), ("bn1_shift", C.c_void_p)]


class UnetLevel(C.Structure):
    """Mirror of ``ud3d_unet_level``."""
    _fields_ = [("c", C.c_int32), ("blocks", UnetBlock * 4), ("tail", UnetBlock * 4),
                ("down_w", C.c_void_p), ("up_w", C.c_void_p),
                ("down_scale", C.c_void_p), ("down_shift", C.c_void_p), ("up_scale", C.c_void_p), ("up_shift", C.c_void_p)]


class UnetPlan(C.Structure):
    """Mirror of ``ud3d_unet_plan``."""
    _fields_ = [("n_levels", C.c_int32), ("block_reps", C.c_int32), ("level", UnetLevel * 8)]


class UnetTables(C.Structure):
    """Mirror of ``ud3d_unet_tables``."""
    _fields_ = [("n", C.c_int32), ("subm", C.c_void_p), ("subm_mask", C.c_void_p), ("row_perm", C.c_void_p),
                ("child", C.c_void_p), ("child_mask", C.c_void_p), ("up", C.c_void_p), ("up_mask", C.c_void_p)]


class Linear(C.Structure):
    """Mirror of ``ud3d_linear``."""
    _fields_ = [("w", C.c_void_p), ("bias", C.c_void_p)]


class EncoderLayer(C.Structure):
    """Mirror of ``ud3d_encoder_layer``."""
    _fields_ = [("qkv", Linear), ("out", Linear), ("f1", Linear), ("f2", Linear),
                ("n1_gamma", C.c_void_p), ("n1_beta", C.c_void_p), ("n1_eps", C.c_float),
                ("n2_gamma", C.c_void_p), ("n2_beta", C.c_void_p), ("n2_eps", C.c_float)]


class EncoderPlan(C.Structure):
    """Mirror of ``ud3d_encoder_plan``."""
    _fields_ = [("num_layers", C.c_int32), ("in_channels", C.c_int32), ("d_model", C.c_int32), ("num_heads", C.c_int32),
                ("hidden", C.c_int32), ("n_union", C.c_int32), ("activation", C.c_int32),
                ("ip0", Linear), ("ip2", Linear), ("layer", EncoderLayer * 12),
                ("on_gamma", C.c_void_p), ("on_beta", C.c_void_p), ("on_eps", C.c_float),
                ("c0", Linear), ("c2", Linear), ("bb", Linear)]


class PostArgs(C.Structure):
    """Mirror of ``ud3d_post_args``."""
    _fields_ = [
        ("logits", C.c_void_p), ("ld_logits", C.c_int32), ("T", C.c_int32), ("C1", C.c_int32),
        ("boxes", C.c_void_p), ("box_dim", C.c_int32),
        ("k", C.c_int32), ("nms_mode", C.c_int32), ("iou_thr", C.c_float), ("score_thr", C.c_float),
        ("use_trim", C.c_int32), ("points", C.c_void_p), ("ld_pts", C.c_int32), ("sp", C.c_void_p),
        ("n_pts", C.c_int32), ("n_sp", C.c_int32), ("low_thr", C.c_float), ("up_thr", C.c_float),
        ("scores", C.c_void_p), ("labels", C.c_void_p), ("cand", C.c_void_p), ("keep", C.c_void_p),
        ("n_keep", C.c_void_p), ("trimmed", C.c_void_p),
    ]


class CriterionArgs(C.Structure):
    """Mirror of ``ud3d_criterion_args``."""
    _fields_ = [
        ("logits", C.c_void_p), ("ld_logits", C.c_int32), ("T", C.c_int32), ("C1", C.c_int32),
        ("boxes", C.c_void_p), ("box_dim", C.c_int32),
        ("gt_boxes", C.c_void_p), ("gt_labels", C.c_void_p), ("G", C.c_int32),
        ("query_masks", C.c_void_p),
        ("topk", C.c_int32), ("w_cls", C.c_float), ("w_box", C.c_float), ("non_object_weight", C.c_float),
        ("match", C.c_void_p), ("sums", C.c_void_p),
    ]


class CriterionGradArgs(C.Structure):
    """Mirror of ``ud3d_criterion_grad_args``."""
    _fields_ = [
        ("logits", C.c_void_p), ("ld_logits", C.c_int32), ("T", C.c_int32), ("C1", C.c_int32),
        ("boxes", C.c_void_p), ("box_dim", C.c_int32),
        ("gt_boxes", C.c_void_p), ("gt_labels", C.c_void_p), ("G", C.c_int32),
        ("match", C.c_void_p), ("sums", C.c_void_p), ("scales", C.c_void_p),
        ("non_object_weight", C.c_float),
        ("d_logits", C.c_void_p), ("ld_dlogits", C.c_int32),
        ("d_boxes", C.c_void_p),
    ]


# name -> (restype, argtypes); must list EVERY symbol declared in include/unidet3d_b200.h
_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
SIGNATURES = {
    "ud3d_version": (C.c_int, []),
    "ud3d_last_error": (C.c_char_p, []),
    "ud3d_launch_count": (C.c_int64, [_i]),
    "ud3d_point_coords_workspace_bytes": (_sz, [_i]),
    "ud3d_point_coords": (_i, [_vp, _i, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ud3d_grid_workspace_bytes": (_sz, [c_i32p]),
    "ud3d_grid_build": (_i, [_vp, _i, c_i32p, _vp, _sz, _vp, _vp]),
    "ud3d_subm3_tile_order_workspace_bytes": (_sz, [_i]),
    "ud3d_subm3_tile_order": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ud3d_grid_build_coarser": (_i, [c_i32p, _vp, c_i32p, _vp, _sz, _vp, _vp]),
    "ud3d_grid_rank": (_i, [_vp, _i, c_i32p, _vp, _vp, _vp]),
    "ud3d_grid_coords": (_i, [c_i32p, _vp, _vp, _i, _vp]),
    "ud3d_voxel_mean": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ud3d_rulebook_subm3": (_i, [_vp, _i, c_i32p, _vp, _vp, _vp, _vp, _vp]),
    "ud3d_down2_parents": (_i, [_vp, _i, c_i32p, _vp, _vp]),
    "ud3d_down_ancestors": (_i, [_vp, _i, c_i32p, _i, _vp, _vp]),
    "ud3d_rulebook_down2": (_i, [_vp, _vp, _i, _i, c_i32p, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ud3d_gemm_packed_weight_bytes": (_sz, [_i, _i, _i]),
    "ud3d_gemm_pack_weight": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "ud3d_gemm_pack_weight_ts": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "ud3d_bn_batch_sums_workspace_bytes": (C.c_size_t, [_i, _i]),
    "ud3d_bn_batch_sums": (_i, [_vp, _i, _i, _i, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_bn_train_fold": (_i, [_vp, C.c_double, _i, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ud3d_conv_wgrad_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "ud3d_segmented_mean_workspace_bytes": (C.c_size_t, [_i, _i]),
    "ud3d_bn_backward_sums": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_bn_backward_apply": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, C.c_double, _vp, _i, _i, _vp, _vp]),
    "ud3d_bn_relu_apply": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "ud3d_attention_bwd_workspace_bytes": (C.c_size_t, [_i, _i]),
    "ud3d_attention_bwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_attention_bwd_reg": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_layernorm_backward_workspace_bytes": (C.c_size_t, [_i, _i]),
    "ud3d_layernorm_backward": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_activation_backward": (_i, [_vp, _vp, C.c_longlong, _i, _vp, _vp]),
    "ud3d_activation_forward": (_i, [_vp, C.c_longlong, _i, _vp, _vp]),
    "ud3d_segmented_mean_backward_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "ud3d_segmented_mean_backward": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_elastic_workspace_bytes": (C.c_size_t, [_vp]),
    "ud3d_elastic_blur": (_i, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_points_to_voxel_units": (_i, [_vp, _i, _i, _f, _vp, _vp]),
    "ud3d_elastic_apply": (_i, [_vp, _i, _vp, _vp, C.c_double, C.c_double, _vp, _vp]),
    "ud3d_elastic_voxel_coords_workspace_bytes": (C.c_size_t, [_i]),
    "ud3d_elastic_voxel_coords": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_compact_ids_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "ud3d_compact_ids": (_i, [_vp, _i, C.c_int64, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_ctx_current": (_vp, []),
    "ud3d_ctx_device": (_i, [_vp]),
    "ud3d_ctx_sm_count": (_i, [_vp]),
    "ud3d_eval_workspace_bytes": (C.c_size_t, [_i, _i, _i]),
    "ud3d_eval_detections": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_encoder_workspace_bytes": (C.c_size_t, [_vp, _i]),
    "ud3d_encoder_forward": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_unet_workspace_bytes": (C.c_size_t, [_vp, _vp]),
    "ud3d_unet_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "ud3d_conv_wgrad": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i, _vp, C.c_size_t, _vp]),
    "ud3d_gemm_fwd": (_i, [C.POINTER(GemmArgs), _vp]),
    "ud3d_act_split": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "ud3d_gemm_fwd_simt": (_i, [C.POINTER(GemmArgs), _vp, _vp]),
    "ud3d_segmented_mean": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "ud3d_layernorm": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "ud3d_attention_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "ud3d_attention_fwd_split": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "ud3d_attention_fwd_opform": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "ud3d_attention_fwd_tc": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ud3d_layernorm_split": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "ud3d_bbox_decode": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp]),
    "ud3d_gather_columns": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp]),
    "ud3d_topk_scores": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ud3d_nms_workspace_bytes": (_sz, [_i]),
    "ud3d_nms_multiclass": (_i, [_vp, _i, _vp, _vp, _i, _i, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "ud3d_trim_workspace_bytes": (_sz, [_i, _i, _i]),
    "ud3d_postprocess_workspace_bytes": (_sz, [C.POINTER(PostArgs)]),
    "ud3d_postprocess_scene": (_i, [C.POINTER(PostArgs), _vp, _sz, _vp]),
    "ud3d_boxes_by_instance": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "ud3d_targets_by_distance": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "ud3d_criterion_workspace_bytes": (_sz, [_i, _i]),
    "ud3d_criterion_layer": (_i, [C.POINTER(CriterionArgs), _vp, _sz, _vp]),
    "ud3d_criterion_layer_grad": (_i, [C.POINTER(CriterionGradArgs), _vp]),
    "ud3d_head_backward": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _i, _vp, _i, _i, _vp]),
    "ud3d_trim_boxes": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _f, _f, _vp, _vp, _sz, _vp]),
}

_lib = None


class Ud3dError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Ud3dError(
            f"{LIB_PATH} not found: the CUDA extension is not built.  Run `python -m unidet3d_b200.build` "
            "(needs nvcc).  unidet3d_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if a declared symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ud3d_last_error().decode("utf-8", "replace")
        raise Ud3dError(f"{what} failed (code {rc}): {msg}")
