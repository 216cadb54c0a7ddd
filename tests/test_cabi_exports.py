"""The C-ABI library loads, exports every symbol include/keaki_b200.h declares, and the product path
fails loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "keaki_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from keaki_b200 import _ffi
    lib = _ffi.load_library()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"libkeaki_b200.so does not export {s}"
    assert sorted(_ffi.SIGNATURES) == syms, "python binding table and header disagree"
    assert b"sm_100a" in lib.kb_version()


def test_null_context_is_rejected_not_crashing():
    from keaki_b200 import _ffi
    lib = _ffi.load_library()
    assert lib.kb_msm_g1(None, None, 0, 0, None, None) == _ffi.KB_ERR_ARG
    assert lib.kb_launch_count(None) == 0
    assert lib.kb_last_error(None) == b"null context"


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for CPU-only boxes")
    from keaki_b200 import _ffi
    with pytest.raises(_ffi.KeakiB200Error):
        _ffi.Context(0)


def test_product_package_never_imports_the_oracle():
    """a product path routed through oracle/ would void every parity claim"""
    pkg = os.path.join(ROOT, "keaki_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"
                assert "hostemu" not in text or f.endswith(".cuh"), f"{f} references the test-only host emulation"
