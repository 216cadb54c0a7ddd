"""Device arithmetic headers, compiled for the host with emulated carry flags, vs the oracle.

This is how the CUDA algorithm text is validated on a box without a GPU; the GPU tests
(-m gpu) then only have to confirm the PTX primitives behave like their emulation."""
import ctypes
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import keaki_ref as kr
from tests import limbs as L
from tests.hostemu import lib as HE

he = HE.load()
P = HE.ptr
rng = random.Random(0xB200)

EDGE = [0, 1, 2, bn.Q - 1, bn.Q - 2, (1 << 253), (1 << 254) % bn.Q, 0xFFFFFFFF, 1 << 32, (1 << 224) - 1]


def _fp_case(field, mod):
    vals = [v % mod for v in EDGE] + [rng.randrange(mod) for _ in range(200)]
    a = [rng.choice(vals) for _ in range(400)]
    b = [rng.choice(vals) for _ in range(400)]
    n = len(a)
    tm = (lambda x: L.int_to_limbs(bn.to_mont(x, mod)))
    A = np.concatenate([tm(x) for x in a]); B = np.concatenate([tm(x) for x in b])
    out = np.zeros_like(A)
    res = {}
    for op in range(8):
        he.he_fp_op(field, op, P(A), P(B), P(out), n)
        res[op] = [L.limbs_to_int(out[8 * i: 8 * i + 8]) for i in range(n)]
    fm = (lambda x: bn.from_mont(x, mod))
    for i in range(n):
        assert fm(res[0][i]) == (a[i] + b[i]) % mod
        assert fm(res[1][i]) == (a[i] - b[i]) % mod
        assert fm(res[2][i]) == (a[i] * b[i]) % mod
        assert fm(res[3][i]) == (-a[i]) % mod
        assert fm(res[4][i]) == (pow(a[i], -1, mod) if a[i] else 0)
        assert res[5][i] == a[i]                       # from_mont gives the canonical integer
        assert fm(res[6][i]) == bn.to_mont(a[i], mod)  # to_mont of the Montgomery image
        assert fm(res[7][i]) == a[i] * a[i] % mod
        for op in range(8):
            assert res[op][i] < mod                    # always fully reduced


def test_fq_ops():
    _fp_case(0, bn.Q)


def test_fr_ops():
    _fp_case(1, bn.R)


def rand_f2(): return (rng.randrange(bn.Q), rng.randrange(bn.Q))
def rand_f12(): return tuple(tuple(rand_f2() for _ in range(3)) for _ in range(2))


def test_fq2_ops():
    for _ in range(50):
        a, b = rand_f2(), rand_f2()
        out = np.zeros(16, np.uint32)
        A, B = L.f2_m(a), L.f2_m(b)
        he.he_fq2_op(0, P(A), P(B), P(out)); assert L.f2_from(out) == bn.f2_mul(a, b)
        he.he_fq2_op(1, P(A), P(B), P(out)); assert L.f2_from(out) == bn.f2_sqr(a)
        he.he_fq2_op(2, P(A), P(B), P(out)); assert L.f2_from(out) == bn.f2_inv(a)
        he.he_fq2_op(3, P(A), P(B), P(out)); assert L.f2_from(out) == bn.f2_mul_xi(a)


def test_fq12_ops():
    for _ in range(10):
        a, b = rand_f12(), rand_f12()
        A, B = L.f12_m(a), L.f12_m(b)
        out = np.zeros(96, np.uint32)
        he.he_fq12_op(0, P(A), P(B), P(out)); assert L.f12_from(out) == bn.f12_mul(a, b)
        he.he_fq12_op(1, P(A), P(B), P(out)); assert L.f12_from(out) == bn.f12_sqr(a)
        he.he_fq12_op(2, P(A), P(B), P(out)); assert L.f12_from(out) == bn.f12_inv(a)
        he.he_fq12_op(7, P(A), P(B), P(out)); assert L.f12_from(out) == bn.f12_conj(a)
        for k in (1, 2, 3):
            he.he_fq12_op(3 + k, P(A), P(B), P(out)); assert L.f12_from(out) == bn.f12_frobenius(a, k)
        # cyclotomic squaring on an element of the cyclotomic subgroup
        easy = np.zeros(96, np.uint32)
        he.he_fq12_easy(P(A), P(easy))
        e = L.f12_from(easy)
        assert e == bn.f12_pow(a, (bn.Q**6 - 1) * (bn.Q**2 + 1))
        he.he_fq12_op(3, P(easy), P(B), P(out)); assert L.f12_from(out) == bn.f12_sqr(e)


def test_mul_by_line():
    for _ in range(10):
        a = rand_f12()
        l0, l1, l3 = rand_f2(), rand_f2(), rand_f2()
        line = bn._f12_from_w([l0, l1, bn.F2_ZERO, l3, bn.F2_ZERO, bn.F2_ZERO])
        out = np.zeros(96, np.uint32)
        he.he_mul_by_line(P(L.f12_m(a)), P(L.f2_m(l0)), P(L.f2_m(l1)), P(L.f2_m(l3)), P(out))
        assert L.f12_from(out) == bn.f12_mul(a, line)


def test_g1_arith():
    pts = [bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R)) for _ in range(6)]
    out = np.zeros(16, np.uint32)
    for p in pts:
        for k in [0, 1, 2, bn.R - 1, bn.R, rng.randrange(bn.R), rng.randrange(1 << 64)]:
            K = L.int_to_limbs(k)
            he.he_g1_mul_add(P(L.g1_m(p)), P(K), None, P(out))
            assert L.g1_from(out) == bn.g1_mul(p, k)
            q = rng.choice(pts)
            he.he_g1_mul_add(P(L.g1_m(p)), P(K), P(L.g1_m(q)), P(out))
            assert L.g1_from(out) == bn.g1_add(bn.g1_mul(p, k), q)
    # exceptional cases of the mixed and full additions: P + P, P + (-P), inf + P
    p = pts[0]
    one = L.int_to_limbs(1)
    he.he_g1_mul_add(P(L.g1_m(p)), P(one), P(L.g1_m(p)), P(out)); assert L.g1_from(out) == bn.g1_add(p, p)
    he.he_g1_mul_add(P(L.g1_m(p)), P(one), P(L.g1_m(bn.g1_neg(p))), P(out)); assert L.g1_from(out) is None
    he.he_g1_lincomb(P(L.g1_m(p)), P(L.int_to_limbs(5)), P(L.g1_m(p)), P(L.int_to_limbs(5)), P(out))
    assert L.g1_from(out) == bn.g1_mul(p, 10)
    he.he_g1_lincomb(P(L.g1_m(p)), P(L.int_to_limbs(5)), P(L.g1_m(p)), P(L.int_to_limbs(bn.R - 5)), P(out))
    assert L.g1_from(out) is None
    he.he_g1_lincomb(P(L.g1_m(p)), P(L.int_to_limbs(0)), P(L.g1_m(pts[1])), P(L.int_to_limbs(7)), P(out))
    assert L.g1_from(out) == bn.g1_mul(pts[1], 7)


def test_g2_arith():
    out = np.zeros(32, np.uint32)
    q = bn.g2_mul(bn.G2_GEN, rng.randrange(1, bn.R))
    for k in [0, 1, 3, bn.R - 1, rng.randrange(bn.R)]:
        he.he_g2_mul_add(P(L.g2_m(bn.G2_GEN)), P(L.int_to_limbs(k)), P(L.g2_m(q)), P(out))
        assert L.g2_from(out) == bn.g2_add(bn.g2_mul(bn.G2_GEN, k), q)


def test_pairing_bytes_and_key():
    cases = [(bn.G1_GEN, bn.G2_GEN)]
    for _ in range(3):
        cases.append((bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R)), bn.g2_mul(bn.G2_GEN, rng.randrange(1, bn.R))))
    cases += [(None, bn.G2_GEN), (bn.G1_GEN, None)]
    for p, q in cases:
        out = (ctypes.c_uint8 * 384)()
        he.he_pairing_bytes(P(L.g1_m(p)), P(L.g2_m(q)), out)
        want = bn.gt_to_bytes(bn.pairing(p, q))
        assert bytes(out) == want
        for ln in (0, 1, 31, 32, 64, 65, 200):
            msg = bytes(rng.randrange(256) for _ in range(ln))
            key = (ctypes.c_uint8 * max(ln, 1))()
            he.he_gt_key(want, None, key, ctypes.c_uint64(ln))
            assert bytes(key)[:ln] == kr.gt_key(bn.pairing(p, q), ln)
            he.he_gt_key(want, msg, key, ctypes.c_uint64(ln))
            assert bytes(key)[:ln] == bytes(a ^ b for a, b in zip(kr.gt_key(bn.pairing(p, q), ln), msg))


def test_miller_matches_up_to_subfield_factor():
    # Miller values may differ by a subfield factor; they must agree after the final exponentiation
    p, q = bn.g1_mul(bn.G1_GEN, 7), bn.g2_mul(bn.G2_GEN, 11)
    out = np.zeros(96, np.uint32)
    he.he_miller(P(L.g1_m(p)), P(L.g2_m(q)), P(out))
    f = L.f12_from(out)
    assert bn.final_exponentiation(f) == bn.pairing(p, q)
    fe = np.zeros(96, np.uint32)
    he.he_final_exp(P(out), P(fe))
    assert L.f12_from(fe) == bn.pairing(p, q)


@pytest.mark.parametrize("c", [3, 8, 13, 15, 16])
def test_signed_digits(c):
    nwin = (255 + c - 1) // c
    for k in [0, 1, bn.R - 1, (1 << 254) - 1, (1 << (c - 1)), (1 << (c - 1)) + 1, (1 << c) - 1] + [rng.randrange(bn.R) for _ in range(50)]:
        d = (ctypes.c_int32 * nwin)()
        he.he_msm_digits(P(L.int_to_limbs(k)), c, nwin, d)
        assert sum(int(d[w]) << (c * w) for w in range(nwin)) == k
        assert all(-(1 << (c - 1)) < int(x) <= (1 << (c - 1)) for x in d)


def test_binary_inversion_edge_representations():
    """fp_inv is Kaliski's almost-inverse + a power-of-two fix-up that branches on k <= 256 (fp.cuh): exercise Montgomery
    images at the extremes of the iteration count (1, 2, small, p - 1, powers of two) for both fields."""
    for field, mod in ((0, bn.Q), (1, bn.R)):
        rinv = pow(1 << 256, -1, mod)
        images = [1, 2, 3, 4, 5, mod - 1, mod - 2, 1 << 253, (1 << 253) + 1, (mod - 1) // 2, (mod + 1) // 2, 1 << 128, (1 << 200) - 1]
        for am in images:
            a = am * rinv % mod
            A = L.int_to_limbs(am)
            out = np.zeros(8, np.uint32)
            he.he_fp_op(field, 4, P(A), P(A), P(out), 1)
            assert bn.from_mont(L.limbs_to_int(out), mod) == pow(a, -1, mod), (field, am)


def test_glv_scalar_multiplication_on_g1():
    """glv.cuh: the division-free decomposition k = k1 + k2 lambda (|k1|, |k2| < 2^128, congruence mod r) and
    k * P by the shared-table double scalar multiplication, against the oracle's plain double-and-add."""
    import os
    import re
    txt = open(os.path.join(os.path.dirname(__file__), "..", "keaki_b200", "csrc", "glv_gen.cuh")).read()

    def table(name):
        body = re.search(r"KB_LIMB_TABLE\(%s, ([^)]*)\)" % name, txt).group(1)
        return sum(int(x.strip().rstrip("u"), 16) << (32 * i) for i, x in enumerate(body.split(",")))
    lam, beta = table("lambda"), bn.from_mont(table("beta"), bn.Q)
    assert pow(lam, 3, bn.R) == 1 and lam != 1 and pow(beta, 3, bn.Q) == 1 and beta != 1
    a1, a2, b1, b2 = table("a1"), table("a2"), table("b1abs"), table("b2")
    assert (a1 - b1 * lam) % bn.R == 0 and (a2 + b2 * lam) % bn.R == 0 and a1 * b2 + a2 * b1 == bn.R
    assert table("g1c") == (b2 << 256) // bn.R and table("g2c") == (b1 << 256) // bn.R
    P0 = bn.g1_mul(bn.G1_GEN, 0xDEADBEEFCAFE)
    assert ((beta * P0[0]) % bn.Q, P0[1]) == bn.g1_mul(P0, lam)             # phi(P) = lambda P
    ks = [0, 1, 2, 3, bn.R - 1, bn.R - 2, lam, lam - 1, lam + 1, bn.R // 2, 1 << 253, 1 << 128, (1 << 128) - 1, 1 << 127] + [rng.randrange(bn.R) for _ in range(40)]
    A = L.g1_m(P0)
    for k in ks:
        out = np.zeros(16, np.uint32); split = np.zeros(12, np.uint32)
        he.he_g1_mul_glv(P(A), P(L.int_to_limbs(k)), P(out), P(split))
        m1, m2 = L.limbs_to_int(split[:5]), L.limbs_to_int(split[5:10])
        s1 = -m1 if split[10] else m1
        s2 = -m2 if split[11] else m2
        assert m1 < (1 << 128) and m2 < (1 << 128) and (s1 + s2 * lam - k) % bn.R == 0, k
        assert L.g1_from(out) == bn.g1_mul(P0, k), k
