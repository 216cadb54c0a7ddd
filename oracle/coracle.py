"""Loader for the C restatement of the reference CPU path (oracle/c/keaki_oracle.c).
TEST INFRASTRUCTURE + CPU BASELINE: only tests/, __graft_entry__.smoke() and bench.py may import this."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libkeaki_oracle.so")
SRC = os.path.join(HERE, "c", "keaki_oracle.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "c")] + (["-B"] if force else []))
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.ko_max_threads.restype = ctypes.c_int
        _lib.ko_init()
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def count_muls(enable: bool) -> int:
    """Field products (Fq and Fr, Montgomery) executed by single-thread calls since the last call of this function;
    `enable` switches the counter for what follows.  Instrumentation for the per-unit work figures in DESIGN.md."""
    lib = load()
    lib.ko_count_muls.restype = ctypes.c_uint64
    return int(lib.ko_count_muls(1 if enable else 0))


def max_threads():
    return int(load().ko_max_threads())


def msm_g1(bases_xy, scalars, threads=1):
    """bases (n,16) u32 Montgomery affine, scalars (n,8) u32 Montgomery Fr -> (xy[16], inf)"""
    n = scalars.shape[0]
    out, inf = np.zeros(16, np.uint32), np.zeros(1, np.uint8)
    load().ko_msm_g1(_p(np.ascontiguousarray(bases_xy[:n])), _p(np.ascontiguousarray(scalars)), ctypes.c_size_t(n), _p(out), _p(inf), int(threads))
    return out, int(inf[0])


def g1_multiples(base_xy, n, threads=1):
    """harness helper: (n, 16) u32 array of (i + 1) * base, i < n - distinct valid bases for the timed CPU arm"""
    out = np.zeros((n, 16), np.uint32)
    load().ko_g1_multiples(_p(np.ascontiguousarray(base_xy, np.uint32)), ctypes.c_size_t(n), _p(out), int(threads))
    return out


def g1_mul(p_xy, p_inf, k_limbs):
    out, inf = np.zeros(16, np.uint32), np.zeros(1, np.uint8)
    load().ko_g1_mul(_p(np.ascontiguousarray(p_xy, np.uint32)), ctypes.c_uint8(int(p_inf)), _p(np.ascontiguousarray(k_limbs, np.uint32)), _p(out), _p(inf))
    return out, int(inf[0])


def pairing_batch(g1_xy, g1_inf, g2_xy, g2_inf, threads=1):
    n = g1_xy.shape[0]
    out = np.zeros((n, 384), np.uint8)
    load().ko_pairing_batch(_p(g1_xy), _p(g1_inf), _p(g2_xy), _p(g2_inf), ctypes.c_size_t(n), _p(out), int(threads))
    return out


def encrypt_batch(com_xy, com_inf, tau_g2_xy, points, values, r, msgs, off, threads=1):
    n = points.shape[0]
    total = int(off[-1]) if n else 0
    ct, ct_inf, msg_ct = np.zeros((n, 32), np.uint32), np.zeros(n, np.uint8), np.zeros(max(total, 1), np.uint8)
    load().ko_encrypt_batch(_p(np.ascontiguousarray(com_xy, np.uint32)), ctypes.c_uint8(int(com_inf)), _p(np.ascontiguousarray(tau_g2_xy, np.uint32)),
                            _p(points), _p(values), _p(r), _p(msgs), _p(off), ctypes.c_size_t(n), _p(ct), _p(ct_inf), _p(msg_ct), int(threads))
    return ct, ct_inf, msg_ct


def decrypt_batch(proofs_xy, proofs_inf, ct_xy, ct_inf, msg_ct, off, threads=1):
    n = proofs_xy.shape[0]
    total = int(off[-1]) if n else 0
    out = np.zeros(max(total, 1), np.uint8)
    load().ko_decrypt_batch(_p(proofs_xy), _p(proofs_inf), _p(ct_xy), _p(ct_inf), _p(msg_ct), _p(off), ctypes.c_size_t(n), _p(out), int(threads))
    return out


def blake3_xof(data: bytes, out_len: int) -> bytes:
    out = (ctypes.c_uint8 * max(out_len, 1))()
    load().ko_blake3_xof(data, ctypes.c_size_t(len(data)), out, ctypes.c_size_t(out_len))
    return bytes(out)[:out_len]
