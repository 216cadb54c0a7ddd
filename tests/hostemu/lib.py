"""Builds and loads the TEST-ONLY host emulation of the device arithmetic headers."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostemu.cpp")
OUT = os.path.join(HERE, "_build", "libkb_hostemu.so")
CSRC = os.path.join(HERE, "..", "..", "keaki_b200", "csrc")


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    if _stale():
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", OUT, SRC])
    return ctypes.CDLL(OUT)


def u32(x):
    return np.ascontiguousarray(x, dtype=np.uint32)


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)
