"""ctypes binding of libkeaki_b200.so (include/keaki_b200.h).

Buffers are numpy arrays (host) or torch CUDA tensors (device; their data_ptr is passed and the
library skips the PCIe copies).  There is no fallback: if the shared library is missing, or no
B200-class GPU is visible, importing/creating a context raises."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libkeaki_b200.so")

KB_OK, KB_ERR_CUDA, KB_ERR_ARG, KB_ERR_POLY_TOO_LARGE, KB_ERR_NO_SRS, KB_ERR_DOMAIN, KB_ERR_INVALID_POINT = 0, -1, -2, -3, -4, -5, -6

_vp, _u64, _i32, _u8 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int32, ctypes.c_uint8

# name -> (restype, argtypes); every symbol include/keaki_b200.h declares
SIGNATURES = {
    "kb_version": (ctypes.c_char_p, []),
    "kb_ctx_create": (_i32, [_i32, ctypes.POINTER(_vp)]),
    "kb_ctx_create_multi": (_i32, [ctypes.POINTER(_i32), _i32, ctypes.POINTER(_vp)]),
    "kb_ctx_device_count": (_i32, [_vp]),
    "kb_ctx_destroy": (None, [_vp]),
    "kb_last_error": (ctypes.c_char_p, [_vp]),
    "kb_srs_upload": (_i32, [_vp, _vp, _u64, _vp]),
    "kb_srs_generate": (_i32, [_vp, _vp, _u64, _u64, _vp, _vp]),
    "kb_srs_len": (_u64, [_vp]),
    "kb_srs_validate": (_i32, [_vp, ctypes.POINTER(_u64)]),
    "kb_msm_g1": (_i32, [_vp, _vp, _u64, _u64, _vp, _vp]),
    "kb_g1_mul_gen_batch": (_i32, [_vp, _vp, _u64, _vp, _vp]),
    "kb_g1_sum": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp]),
    "kb_open_batch": (_i32, [_vp, _vp, _u64, _vp, _u64, _vp, _vp]),
    "kb_open_all_fk": (_i32, [_vp, _vp, _u64, _vp, _vp]),
    "kb_fr_ntt": (_i32, [_vp, _vp, _u64, _i32]),
    "kb_encrypt_batch": (_i32, [_vp, _vp, _u8, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp]),
    "kb_decrypt_batch": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "kb_pairing_batch": (_i32, [_vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "kb_verify_batch": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "kb_g1_serialize": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp]),
    "kb_g2_serialize": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp]),
    "kb_g1_deserialize": (_i32, [_vp, _vp, _u64, _i32, _i32, _vp, _vp, _vp]),
    "kb_g2_deserialize": (_i32, [_vp, _vp, _u64, _i32, _i32, _vp, _vp, _vp]),
    "kb_debug_fp_op": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _u64]),
    "kb_launch_count": (_u64, [_vp]),
    "kb_last_kernel_ms": (ctypes.c_float, [_vp, _i32]),
}


class KeakiB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"keaki_b200 error {code}: {msg}")
        self.code = code


class PolynomialTooLarge(KeakiB200Error):
    """KZGError::PolynomialTooLarge(len, max) — src/kzg.rs:205-209."""


class InvalidSrsPoint(KeakiB200Error):
    """kb_srs_validate found an SRS element that is not a point of its group (`.index`: G1 power, or srs_len for [tau]_2)."""


_lib = None


def load_library():
    """Loads libkeaki_b200.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C keaki_b200/csrc` — there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"], "buffers must be C-contiguous"
        return x.ctypes.data
    if hasattr(x, "data_ptr"):  # torch tensor (host pinned or CUDA)
        assert x.is_contiguous()
        return x.data_ptr()
    if isinstance(x, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(x)), _vp).value
    raise TypeError(f"unsupported buffer type {type(x)}")


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Context:
    """One GPU.  Mirrors the C ABI one-to-one; higher layers (kzg/kem/enc/vec) build on it."""

    def __init__(self, device=0):
        """device: an index, or a sequence of indices for ONE context over several GPUs of the box (kb_ctx_create_multi:
        commit splits by point range, encrypt / decrypt batches by index, inside the library)."""
        self.lib = load_library()
        h = _vp()
        if isinstance(device, (list, tuple)):
            devs = (_i32 * len(device))(*[int(d) for d in device])
            rc = self.lib.kb_ctx_create_multi(devs, len(device), ctypes.byref(h))
        else:
            rc = self.lib.kb_ctx_create(int(device), ctypes.byref(h))
        if rc != KB_OK:
            raise KeakiB200Error(rc, f"kb_ctx_create(device={device}) failed: no usable sm_100 GPU (there is no CPU fallback)")
        self.h = h
        self.device = device

    def device_count(self):
        return int(self.lib.kb_ctx_device_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.kb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == KB_OK:
            return
        msg = self.lib.kb_last_error(self.h).decode()
        if rc == KB_ERR_POLY_TOO_LARGE:
            raise PolynomialTooLarge(rc, msg)
        raise KeakiB200Error(rc, msg)

    # ---- SRS
    def srs_upload(self, g1_xy, tau_g2_xy):
        n = g1_xy.shape[0] if hasattr(g1_xy, "shape") and len(g1_xy.shape) == 2 else len(g1_xy) // 16
        self._keep = (g1_xy, tau_g2_xy)
        self._check(self.lib.kb_srs_upload(self.h, _ptr(g1_xy), n, _ptr(tau_g2_xy)))

    def srs_generate(self, tau_limbs, n, download=True, first_power=0):
        g1 = np.zeros((n, 16), np.uint32) if download else None
        t2 = np.zeros(32, np.uint32)
        self._check(self.lib.kb_srs_generate(self.h, _ptr(_u32(tau_limbs)), first_power, n, _ptr(g1), _ptr(t2)))
        return g1, t2

    def srs_validate(self):
        """kb_srs_validate: raises InvalidSrsPoint (with .index) unless every resident SRS element is a point of its group."""
        bad = _u64(0)
        rc = self.lib.kb_srs_validate(self.h, ctypes.byref(bad))
        if rc == KB_ERR_INVALID_POINT:
            e = InvalidSrsPoint(rc, self.lib.kb_last_error(self.h).decode())
            e.index = int(bad.value)
            raise e
        self._check(rc)

    def srs_len(self):
        return int(self.lib.kb_srs_len(self.h))

    # ---- KZG
    def msm_g1(self, scalars, n=None, first=0):
        if n is None:
            n = scalars.shape[0]
        out, inf = np.zeros(16, np.uint32), np.zeros(1, np.uint8)
        self._check(self.lib.kb_msm_g1(self.h, _ptr(scalars), first, n, _ptr(out), _ptr(inf)))
        return out, int(inf[0])

    def g1_mul_gen_batch(self, scalars):
        n = scalars.shape[0]
        out, inf = np.zeros((n, 16), np.uint32), np.zeros(n, np.uint8)
        self._check(self.lib.kb_g1_mul_gen_batch(self.h, _ptr(scalars), n, _ptr(out), _ptr(inf)))
        return out, inf

    def g1_sum(self, pts_xy, inf=None):
        n = pts_xy.shape[0]
        out, oi = np.zeros(16, np.uint32), np.zeros(1, np.uint8)
        self._check(self.lib.kb_g1_sum(self.h, _ptr(pts_xy), _ptr(inf), n, _ptr(out), _ptr(oi)))
        return out, int(oi[0])

    def open_batch(self, coeffs, points):
        d, m = coeffs.shape[0], points.shape[0]
        out, inf = np.zeros((m, 16), np.uint32), np.zeros(m, np.uint8)
        self._check(self.lib.kb_open_batch(self.h, _ptr(coeffs), d, _ptr(points), m, _ptr(out), _ptr(inf)))
        return out, inf

    def open_all_fk(self, coeffs):
        d = coeffs.shape[0]
        out, inf = np.zeros((d, 16), np.uint32), np.zeros(d, np.uint8)
        self._check(self.lib.kb_open_all_fk(self.h, _ptr(coeffs), d, _ptr(out), _ptr(inf)))
        return out, inf

    def fr_ntt(self, data, inverse=False):
        """in place"""
        self._check(self.lib.kb_fr_ntt(self.h, _ptr(data), data.shape[0], 1 if inverse else 0))
        return data

    # ---- witness encryption
    def encrypt_batch(self, com_xy, com_inf, points, values, r, msgs, msg_off, out=None):
        n = points.shape[0]
        total = int(msg_off[-1]) if n else 0
        if out is None:
            out = (np.zeros((n, 32), np.uint32), np.zeros(n, np.uint8), np.zeros(max(total, 1), np.uint8))
        ct, ct_inf, msg_ct = out
        self._check(self.lib.kb_encrypt_batch(self.h, _ptr(_u32(com_xy)), int(com_inf), _ptr(points), _ptr(values), _ptr(r),
                                              _ptr(msgs), _ptr(msg_off), n, _ptr(ct), _ptr(ct_inf), _ptr(msg_ct)))
        return ct, ct_inf, msg_ct

    def decrypt_batch(self, proofs_xy, proofs_inf, ct_xy, ct_inf, msg_ct, msg_off, out=None, n=None):
        if n is None:
            n = proofs_xy.shape[0]
        if out is None:
            total = int(msg_off[-1]) if n else 0
            out = np.zeros(max(total, 1), np.uint8)
        self._check(self.lib.kb_decrypt_batch(self.h, _ptr(proofs_xy), _ptr(proofs_inf), _ptr(ct_xy), _ptr(ct_inf),
                                              _ptr(msg_ct), _ptr(msg_off), n, _ptr(out)))
        return out

    def pairing_batch(self, g1_xy, g1_inf, g2_xy, g2_inf):
        n = g1_xy.shape[0]
        out = np.zeros((n, 384), np.uint8)
        self._check(self.lib.kb_pairing_batch(self.h, _ptr(g1_xy), _ptr(g1_inf), _ptr(g2_xy), _ptr(g2_inf), n, _ptr(out)))
        return out

    def verify_batch(self, com_xy, com_inf, points, values, proofs_xy, proofs_inf):
        n = points.shape[0]
        ok = np.zeros(n, np.uint8)
        self._check(self.lib.kb_verify_batch(self.h, _ptr(com_xy), _ptr(com_inf), _ptr(points), _ptr(values),
                                             _ptr(proofs_xy), _ptr(proofs_inf), n, _ptr(ok)))
        return ok

    # ---- diagnostics
    # ---- wire format (ark-serialize bytes of affine points)
    def g1_serialize(self, xy, inf=None, compress=True):
        n = xy.shape[0]
        out = np.zeros((n, 32 if compress else 64), np.uint8)
        self._check(self.lib.kb_g1_serialize(self.h, _ptr(xy), _ptr(inf), n, 1 if compress else 0, _ptr(out)))
        return out

    def g2_serialize(self, xy, inf=None, compress=True):
        n = xy.shape[0]
        out = np.zeros((n, 64 if compress else 128), np.uint8)
        self._check(self.lib.kb_g2_serialize(self.h, _ptr(xy), _ptr(inf), n, 1 if compress else 0, _ptr(out)))
        return out

    def g1_deserialize(self, data, compress=True, validate=True):
        data = np.ascontiguousarray(data, np.uint8).reshape(-1, 32 if compress else 64)
        n = data.shape[0]
        xy, inf, ok = np.zeros((n, 16), np.uint32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self._check(self.lib.kb_g1_deserialize(self.h, _ptr(data), n, 1 if compress else 0, 1 if validate else 0, _ptr(xy), _ptr(inf), _ptr(ok)))
        return xy, inf, ok

    def g2_deserialize(self, data, compress=True, validate=True):
        data = np.ascontiguousarray(data, np.uint8).reshape(-1, 64 if compress else 128)
        n = data.shape[0]
        xy, inf, ok = np.zeros((n, 32), np.uint32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self._check(self.lib.kb_g2_deserialize(self.h, _ptr(data), n, 1 if compress else 0, 1 if validate else 0, _ptr(xy), _ptr(inf), _ptr(ok)))
        return xy, inf, ok

    def debug_fp_op(self, field, op, a, b):
        out = np.zeros_like(a)
        self._check(self.lib.kb_debug_fp_op(self.h, field, op, _ptr(a), _ptr(b), _ptr(out), a.size // 8))
        return out

    def launch_count(self):
        return int(self.lib.kb_launch_count(self.h))

    def last_kernel_ms(self, which=0):
        return float(self.lib.kb_last_kernel_ms(self.h, which))
