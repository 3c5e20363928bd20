"""CPU: the C-ABI shared library loads and exports every symbol include/unidet3d_b200.h declares
(no compute calls -- there is no GPU in the build container)."""
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "unidet3d_b200.h")).read()
    return set(re.findall(r"\b(ud3d_[a-z0-9_]+)\s*\(", hdr))


def test_library_exports_every_declared_symbol():
    from unidet3d_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    lib = _lib.load()
    declared = _declared()
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ud3d_version() >= 100


def test_argument_errors_are_reported_not_crashing():
    from unidet3d_b200 import _lib
    lib = _lib.load()
    rc = lib.ud3d_gemm_fwd(None, None)
    assert rc == -1 and b"NULL" in lib.ud3d_last_error()
    assert lib.ud3d_gemm_packed_weight_bytes(27, 32, 32) == 27 * 32 * 128
    assert lib.ud3d_gemm_packed_weight_bytes(1, 256, 768) == 768 * 8 * 128


def test_no_cpu_fallback():
    import torch
    from unidet3d_b200 import ops, _lib
    with pytest.raises(_lib.Ud3dError):
        ops.layernorm(torch.zeros(4, 256), torch.ones(256), torch.zeros(256))


def test_stage_plan_evaluator_and_augment_entry_points_validate_arguments():
    """Host-side validation of the entry points added for the stage plans, the evaluator and the augmentations: errors come
    back as codes + messages before any CUDA call (this box has no GPU), workspace queries are pure host arithmetic."""
    import ctypes as C
    from unidet3d_b200 import _lib
    lib = _lib.load()
    EINVAL = -1
    assert lib.ud3d_unet_forward(None, None, None, None, None, None, None, 0, None) == EINVAL
    assert b"NULL" in lib.ud3d_last_error()
    assert lib.ud3d_unet_workspace_bytes(None, None) == 0
    # a two-level plan: workspace = the fixed buffer set of each level (see csrc/unet_plan.cu:carve)
    P = _lib.UnetPlan()
    P.n_levels, P.block_reps = 2, 2
    P.level[0].c, P.level[1].c = 32, 64
    T = (_lib.UnetTables * 2)()
    T[0].n, T[1].n = 1000, 300
    ws = lib.ud3d_unet_workspace_bytes(C.byref(P), T)
    lo = 4 * (1000 * 32 * 11 + 300 * 64 * 3 + 300 * 64 * 5)
    assert lo <= ws <= lo + 64 * 256
    P.level[0].c = 48                                        # not a multiple of 32
    assert lib.ud3d_unet_forward(C.byref(P), T, C.c_void_p(256), C.c_void_p(256), C.c_void_p(256), None, C.c_void_p(256), 1 << 30,
                                 None) == EINVAL
    assert b"multiple of 32" in lib.ud3d_last_error()
    E = _lib.EncoderPlan()
    E.num_layers, E.in_channels, E.d_model, E.num_heads, E.hidden, E.n_union, E.activation = 6, 32, 256, 8, 1024, 19, 2
    assert lib.ud3d_encoder_workspace_bytes(C.byref(E), 1000) >= 4 * 1000 * (256 * 10 + 768 + 1024)
    E.num_heads = 4                                          # head_dim 64: the attention kernel is specialised for 32
    assert lib.ud3d_encoder_forward(C.byref(E), C.c_void_p(256), 10, C.c_void_p(256), 1, 10, C.c_void_p(256), C.c_void_p(256), None,
                                    C.c_void_p(256), 1 << 30, None) == EINVAL
    assert lib.ud3d_eval_detections(None, None, None, 0, None, None, 18, None, None, None, 0, 0, None, 2, None, None, None, None, 0,
                                    None) == EINVAL
    assert lib.ud3d_eval_workspace_bytes(1000, 100, 2) >= 1000 * 4 * 2 + 2 * 100 * 4 + 2 * 1000 * 4
    assert lib.ud3d_elastic_blur(None, None, None, 0, None) == EINVAL
    d = (C.c_int32 * 3)(10, 12, 7)
    assert lib.ud3d_elastic_workspace_bytes(d) == 3 * 10 * 12 * 7 * 4
    assert lib.ud3d_compact_ids(C.c_void_p(8), 4, -5, C.c_void_p(8), C.c_void_p(8), C.c_void_p(8), 1 << 20, None) == EINVAL
    assert lib.ud3d_compact_ids_workspace_bytes(63) == 2 * 8 + 16
    assert lib.ud3d_segmented_mean_workspace_bytes(100, 32) == 100 * (32 * 8 + 4)
    # no CUDA device here: the per-device context cannot be created and says so
    import torch
    if not torch.cuda.is_available():
        assert not lib.ud3d_ctx_current()
        assert b"device" in lib.ud3d_last_error()
        assert lib.ud3d_ctx_sm_count(None) == 0 and lib.ud3d_ctx_device(None) == -1
