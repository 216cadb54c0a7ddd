"""The C++ host layer (include/keaki_b200.hpp — the keaki API restated over the C ABI, since the reference's Rust
toolchain is absent here) and its test program, which re-instantiates the reference's own tests on BN254
(tests/cpp/test_keaki_host.cpp).  CPU: it builds, its host-side Fr arithmetic checks pass and, with no GPU, it fails
loudly instead of falling back.  GPU (marked): every reference test passes through the C++ layer."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
BIN = os.path.join(CPP, "_build", "test_keaki_host")
PTAU = os.path.join(ROOT, "tests", "golden", "ppot_0080_01_mini.ptau")
GOLDEN = os.path.join(ROOT, "tests", "golden", "oracle_vectors.txt")


def _build():
    subprocess.check_call(["make", "-s", "-C", CPP])
    assert os.path.exists(BIN)


def _run():
    return subprocess.run([BIN, PTAU, GOLDEN], capture_output=True, text=True, timeout=900)


def test_cpp_host_layer_builds_and_fails_loudly_without_gpu():
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for CPU-only boxes")
    r = _run()
    assert r.returncode == 3, r.stderr
    assert "[ DONE ] test_fr_host_arithmetic" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_reference_test_suite_through_the_cpp_host_layer():
    _build()
    r = _run()
    assert r.returncode == 0, r.stderr[-4000:]
    assert " 0 failed" in r.stderr
