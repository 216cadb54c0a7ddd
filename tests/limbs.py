"""Helpers converting oracle integers <-> the C ABI's Montgomery 8 x u32 limb arrays (test side)."""
import numpy as np

from oracle import bn254 as bn


def int_to_limbs(x: int) -> np.ndarray:
    return np.frombuffer(int(x).to_bytes(32, "little"), dtype=np.uint32).copy()


def limbs_to_int(a) -> int:
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint32).tobytes(), "little")


def fq_m(x): return int_to_limbs(bn.to_mont(x % bn.Q, bn.Q))
def fr_m(x): return int_to_limbs(bn.to_mont(x % bn.R, bn.R))
def fq_from(a): return bn.from_mont(limbs_to_int(a), bn.Q)
def fr_from(a): return bn.from_mont(limbs_to_int(a), bn.R)


def fq_vec(xs): return np.concatenate([fq_m(x) for x in xs]) if len(xs) else np.zeros(0, np.uint32)
def fr_vec(xs): return np.concatenate([fr_m(x) for x in xs]) if len(xs) else np.zeros(0, np.uint32)
def fr_vec_from(a): return [fr_from(a[8 * i: 8 * i + 8]) for i in range(len(a) // 8)]


def f2_m(a): return np.concatenate([fq_m(a[0]), fq_m(a[1])])
def f2_from(a): return (fq_from(a[:8]), fq_from(a[8:16]))


def f12_m(a):
    return np.concatenate([f2_m(a[j][i]) for j in range(2) for i in range(3)])


def f12_from(w):
    c = [f2_from(w[16 * k: 16 * k + 16]) for k in range(6)]
    return ((c[0], c[1], c[2]), (c[3], c[4], c[5]))


def g1_m(p):
    """affine G1 -> 16 limbs (zeros for infinity)"""
    if p is None:
        return np.zeros(16, np.uint32)
    return np.concatenate([fq_m(p[0]), fq_m(p[1])])


def g1_from(a):
    x, y = fq_from(a[:8]), fq_from(a[8:16])
    return None if (limbs_to_int(a[:8]) == 0 and limbs_to_int(a[8:16]) == 0) else (x, y)


def g2_m(p):
    if p is None:
        return np.zeros(32, np.uint32)
    return np.concatenate([f2_m(p[0]), f2_m(p[1])])


def g2_from(a):
    if not np.any(a[:32]):
        return None
    return (f2_from(a[:16]), f2_from(a[16:32]))


def g1_vec(ps): return np.concatenate([g1_m(p) for p in ps]) if len(ps) else np.zeros(0, np.uint32)
def g2_vec(ps): return np.concatenate([g2_m(p) for p in ps]) if len(ps) else np.zeros(0, np.uint32)
