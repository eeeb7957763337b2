"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/ih_api.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ih_api.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ih_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from imagharmony_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ih_api.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == set(declared)
    assert lib.ih_version() == 1


def test_product_path_fails_loudly_without_gpu_tensors():
    import torch
    from imagharmony_b200 import ops
    from imagharmony_b200._lib import IHError
    with pytest.raises(IHError):
        ops.linear(torch.zeros(8, 8, dtype=torch.float16), torch.zeros(8, 8, dtype=torch.float16))


def test_argument_validation_returns_error_codes():
    from imagharmony_b200 import _lib
    lib = _lib.load()
    rc = lib.ih_gemm_f16(None, 0, None, None, None, 0, 0, None, 0, None, 0, 1, 1, 1, 0, 0, None)
    assert rc < 0 and b"null" in lib.ih_last_error()
    rc = lib.ih_layernorm_f16(1, 1, 1, 1, 4, 7, 1e-5, None)   # C not a multiple of 8 (checked before any launch)
    assert rc < 0
