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
