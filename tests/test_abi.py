"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol that
include/w2v2_b200.h declares, with the binding in _lib.py covering all of them (no compute calls)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "w2v2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(w2v2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from w2v2_speaker_b200 import _lib, build
    build.build()
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
        assert s in _lib.SIGNATURES, f"{s} declared in the header but not bound in _lib.py"
    for s in _lib.SIGNATURES:
        assert s in syms, f"{s} bound in _lib.py but not declared in include/w2v2_b200.h"
    assert lib.w2v2_abi_version() >= 1


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from w2v2_speaker_b200 import ops
    from w2v2_speaker_b200._lib import W2V2Error
    with pytest.raises(W2V2Error):
        ops.gemm_f16(torch.zeros(4, 64, dtype=torch.float16), torch.zeros(4, 64, dtype=torch.float16))
