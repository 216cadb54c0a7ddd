"""The compiled single-thread pairing (keaki_b200/csrc/pairing_st.cuh) and the lazy-reduction field routines
(fpl.cuh) vs the oracle, through the TEST-ONLY host build of the device headers (tests/hostemu): the exact
algorithm text the GPU runs, with the PTX carry chains emulated."""
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from tests import limbs as L
from tests.hostemu import lib as HE

rng = random.Random(0x57A7)
he = HE.load()
P = HE.ptr
RINV = pow(1 << 256, -1, bn.Q)


def _rand_fq():
    return rng.randrange(bn.Q)


def _edge_values():
    return [0, 1, 2, bn.Q - 1, bn.Q - 2, (1 << 253) % bn.Q, (1 << 32) - 1, (1 << 224) - 1, bn.Q >> 1]


def test_mul_wide_and_redc():
    vals = _edge_values() + [_rand_fq() for _ in range(40)]
    for a in vals:
        for b in vals[:12] + [_rand_fq()]:
            for xa, xb in ((a, b), (min(a + bn.Q, (1 << 256) - 1), b)):   # unreduced multiplicands are legal inputs
                A, B = HE.u32(L.int_to_limbs(xa)), HE.u32(L.int_to_limbs(xb))
                out = np.zeros(16, np.uint32)
                he.he_lazy_op(0, P(A), P(B), P(out))
                assert int.from_bytes(out.tobytes(), "little") == xa * xb
    # redc: t < q R  ->  t R^-1 mod q, fully reduced
    ts = [0, 1, bn.Q, bn.Q * (1 << 256) - 1, (bn.Q - 1) ** 2, 2 * bn.Q * bn.Q - 1] + [rng.randrange(bn.Q << 256) for _ in range(200)]
    for t in ts:
        T = np.frombuffer(int(t).to_bytes(64, "little"), np.uint32).copy()
        out = np.zeros(8, np.uint32)
        he.he_lazy_op(1, P(T), P(T), P(out))
        assert L.limbs_to_int(out) == t * RINV % bn.Q


def test_fq2_mul_lazy():
    edge = _edge_values()
    cases = [((a, b), (c, d)) for a in edge[:5] for b in edge[:5] for c in (0, 1, bn.Q - 1) for d in (0, bn.Q - 1)]
    cases += [((_rand_fq(), _rand_fq()), (_rand_fq(), _rand_fq())) for _ in range(300)]
    for x, y in cases:
        X, Y = HE.u32(L.f2_m(x)), HE.u32(L.f2_m(y))
        out = np.zeros(16, np.uint32)
        he.he_lazy_op(2, P(X), P(Y), P(out))
        assert L.f2_from(out) == bn.f2_mul(x, y)


def _rand_f12():
    return tuple(tuple((_rand_fq(), _rand_fq()) for _ in range(3)) for _ in range(2))


def _f12_op(op, a, b=None):
    A = HE.u32(L.f12_m(a))
    B = HE.u32(b) if b is not None else A
    out = np.zeros(96, np.uint32)
    he.he_st_f12_op(op, P(A), P(B), P(out))
    return L.f12_from(out)


def test_f12_routines():
    for _ in range(3):
        a, b = _rand_f12(), _rand_f12()
        assert _f12_op(0, a, L.f12_m(b)) == bn.f12_mul(a, b)
        assert _f12_op(1, a, L.f12_m(b)) == bn.f12_mul(a, bn.f12_conj(b))
        assert _f12_op(2, a) == bn.f12_mul(a, a)
        assert bn.f12_mul(_f12_op(8, a), a) == bn.F12_ONE
        for k in (1, 2, 3):
            assert _f12_op(4 + k, a) == bn.f12_frobenius(a, k)
        # sparse line: l0 + l1 w + l3 w^3 = c0 (l0, 0, 0), c1 (l1, l3, 0)
        l0, l1, l3 = [(_rand_fq(), _rand_fq()) for _ in range(3)]
        z = (0, 0)
        line = ((l0, z, z), (l1, l3, z))
        lb = np.concatenate([L.f2_m(l0), L.f2_m(l1), L.f2_m(l3)])
        assert _f12_op(4, a, lb) == bn.f12_mul(a, line)
        # cyclotomic squaring on an element of the cyclotomic subgroup
        e = bn.f12_mul(bn.f12_conj(a), bn.f12_inv(a))
        e = bn.f12_mul(bn.f12_frobenius(e, 2), e)
        assert _f12_op(3, e) == bn.f12_mul(e, e)


@pytest.mark.parametrize("k", range(3))
def test_pairing_bytes_match_oracle(k):
    if k == 0:
        p, q = bn.G1_GEN, bn.G2_GEN
    else:
        p, q = bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R)), bn.g2_mul(bn.G2_GEN, rng.randrange(1, bn.R))
    out = np.zeros(384, np.uint8)
    he.he_st_pairing_bytes(P(HE.u32(L.g1_m(p))), P(HE.u32(L.g2_m(q))), P(out))
    assert bytes(out) == bn.gt_to_bytes(bn.pairing(p, q))
