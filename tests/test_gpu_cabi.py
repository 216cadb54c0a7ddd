"""GPU parity tests proper: every C-ABI entry point vs the oracle on the same seeded inputs
(bit-exact: integer/byte work).  Run on the B200 box with `pytest -m gpu`."""
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import keaki_ref as kr
from tests import limbs as L

pytestmark = pytest.mark.gpu

rng = random.Random(0x6B65616B69)
TAU = rng.randrange(1, bn.R)


@pytest.fixture(scope="module")
def ctx():
    from keaki_b200 import _ffi
    c = _ffi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def srs(ctx):
    """2^13 synthetic powers generated on the device, checked against the oracle below."""
    n = 1 << 13
    g1, tau2 = ctx.srs_generate(L.fr_m(TAU), n)
    return g1, tau2


def trapdoor_commit(scalars, first=0):
    """(sum_i s_i tau^(first+i)) * G1 — the trapdoor oracle of SURVEY.md §7."""
    acc, t = 0, pow(TAU, first, bn.R)
    for s in scalars:
        acc = (acc + s * t) % bn.R
        t = t * TAU % bn.R
    return bn.g1_mul(bn.G1_GEN, acc)


def g1_out(xy, inf):
    return None if inf else L.g1_from(xy)


# ---------------------------------------------------------------------------------------------
def test_field_primitives_on_device(ctx):
    edge = [0, 1, 2, 0xFFFFFFFF, 1 << 32, (1 << 224) - 1, 1 << 253]
    for field, mod in ((0, bn.Q), (1, bn.R)):
        vals = [v % mod for v in edge] + [mod - 1, mod - 2] + [rng.randrange(mod) for _ in range(300)]
        a = [rng.choice(vals) for _ in range(1000)]
        b = [rng.choice(vals) for _ in range(1000)]
        A = np.concatenate([L.int_to_limbs(bn.to_mont(x, mod)) for x in a])
        B = np.concatenate([L.int_to_limbs(bn.to_mont(x, mod)) for x in b])
        fm = lambda x: bn.from_mont(x, mod)  # noqa: E731
        want = {
            0: lambda x, y: (x + y) % mod, 1: lambda x, y: (x - y) % mod, 2: lambda x, y: x * y % mod,
            3: lambda x, y: -x % mod, 4: lambda x, y: pow(x, -1, mod) if x else 0, 7: lambda x, y: x * x % mod,
        }
        for op, f in want.items():
            out = ctx.debug_fp_op(field, op, A, B)
            got = [L.limbs_to_int(out[8 * i: 8 * i + 8]) for i in range(len(a))]
            assert all(g < mod for g in got)
            assert [fm(g) for g in got] == [f(x, y) for x, y in zip(a, b)], f"field {field} op {op}"
        out = ctx.debug_fp_op(field, 5, A, B)
        assert [L.limbs_to_int(out[8 * i: 8 * i + 8]) for i in range(len(a))] == a


def test_srs_generate_matches_setup(ctx, srs):
    g1, tau2 = srs
    ref = kr.KZGSetup.setup(TAU, 6)          # src/kzg.rs:218-239 (test_kzg_setup)
    for i in range(6):
        assert L.g1_from(g1[i]) == ref.g1_aff[i]
    for i in (100, 4095, 8191):
        assert L.g1_from(g1[i]) == bn.g1_mul(bn.G1_GEN, pow(TAU, i, bn.R))
    assert L.g2_from(tau2) == ref.tau_g2
    assert ctx.srs_len() == 1 << 13


@pytest.mark.parametrize("n", [1, 2, 3, 31, 255, 256, 257, 1000, 4096, 8192])
def test_msm_matches_trapdoor(ctx, srs, n):
    scalars = [rng.randrange(bn.R) for _ in range(n)]
    xy, inf = ctx.msm_g1(L.fr_vec(scalars).reshape(n, 8))
    assert g1_out(xy, inf) == trapdoor_commit(scalars)


def test_msm_edge_scalars_and_naive_sum(ctx, srs):
    g1, _ = srs
    # commit == sum coeff_i * g1_pow[i]  (src/kzg.rs:241-258), special scalars 0, 1, r-1
    scalars = [0, 1, bn.R - 1, 2, 0, bn.R - 2, 1 << 253, 5]
    xy, inf = ctx.msm_g1(L.fr_vec(scalars).reshape(-1, 8))
    pts = [L.g1_from(g1[i]) for i in range(len(scalars))]
    assert g1_out(xy, inf) == bn.g1_msm(pts, scalars)
    # all-zero scalars -> identity; empty -> identity
    xy, inf = ctx.msm_g1(L.fr_vec([0] * 5).reshape(-1, 8))
    assert inf == 1
    xy, inf = ctx.msm_g1(np.zeros((0, 8), np.uint32), n=0)
    assert inf == 1
    # cancelling pair: s*P0 + (-s*tau^-1... ) use P0 with s and r-s on the same scalar slot twice is not possible;
    # instead check a result that is a small multiple, exercising doubling inside buckets:
    same = [7] * 300                                   # every point lands in the same bucket per window
    xy, inf = ctx.msm_g1(L.fr_vec(same).reshape(-1, 8))
    assert g1_out(xy, inf) == trapdoor_commit(same)


def test_msm_point_range_shards_sum_to_whole(ctx, srs):
    n = 3000
    scalars = [rng.randrange(bn.R) for _ in range(n)]
    S = L.fr_vec(scalars).reshape(n, 8)
    parts, infs = [], []
    for lo, hi in ((0, 700), (700, 701), (701, 2048), (2048, 3000)):
        xy, inf = ctx.msm_g1(np.ascontiguousarray(S[lo:hi]), first=lo)
        assert g1_out(xy, inf) == trapdoor_commit(scalars[lo:hi], first=lo)
        parts.append(xy); infs.append(inf)
    xy, inf = ctx.g1_sum(np.stack(parts), np.array(infs, np.uint8))
    assert g1_out(xy, inf) == trapdoor_commit(scalars)


def test_msm_too_large_is_an_error(ctx, srs):
    from keaki_b200 import _ffi
    with pytest.raises(_ffi.PolynomialTooLarge):      # src/kzg.rs:260-277
        ctx.msm_g1(np.zeros(((1 << 13) + 1, 8), np.uint32))


def test_g1_sum(ctx):
    pts = [bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R)) for _ in range(9)]
    pts += [None, bn.g1_neg(pts[0]), pts[1]]
    inf = np.array([1 if p is None else 0 for p in pts], np.uint8)
    xy, oi = ctx.g1_sum(L.g1_vec(pts).reshape(-1, 16), inf)
    want = None
    for p in pts:
        want = bn.g1_add(want, p)
    assert g1_out(xy, oi) == want


@pytest.mark.parametrize("logn", [0, 1, 2, 5, 10])
def test_fr_ntt(ctx, logn):
    n = 1 << logn
    vals = [rng.randrange(bn.R) for _ in range(n)]
    dom = bn.Radix2Domain(n)
    a = L.fr_vec(vals).reshape(n, 8)
    ctx.fr_ntt(a, inverse=False)
    assert L.fr_vec_from(a.reshape(-1)) == dom.fft(vals)
    a = L.fr_vec(vals).reshape(n, 8)
    ctx.fr_ntt(a, inverse=True)
    assert L.fr_vec_from(a.reshape(-1)) == dom.ifft(vals)


def test_pairing_batch(ctx):
    cases = [(bn.G1_GEN, bn.G2_GEN), (None, bn.G2_GEN), (bn.G1_GEN, None)]
    for _ in range(5):
        cases.append((bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R)), bn.g2_mul(bn.G2_GEN, rng.randrange(1, bn.R))))
    g1 = L.g1_vec([c[0] for c in cases]).reshape(-1, 16)
    g2 = L.g2_vec([c[1] for c in cases]).reshape(-1, 32)
    i1 = np.array([1 if c[0] is None else 0 for c in cases], np.uint8)
    i2 = np.array([1 if c[1] is None else 0 for c in cases], np.uint8)
    out = ctx.pairing_batch(g1, i1, g2, i2)
    for k, (p, q) in enumerate(cases):
        assert bytes(out[k]) == bn.gt_to_bytes(bn.pairing(p, q)), f"case {k}"


def _setup_small(ctx, d):
    """polynomial of d coefficients committed on the resident SRS; returns oracle-side objects"""
    p = [rng.randrange(bn.R) for _ in range(d)]
    com = trapdoor_commit(p)
    return p, com


def test_open_batch_and_verify(ctx, srs):
    d = 37
    p, com = _setup_small(ctx, d)
    points = [5, 0, 1, bn.R - 1, rng.randrange(bn.R)]
    proofs, inf = ctx.open_batch(L.fr_vec(p).reshape(d, 8), L.fr_vec(points).reshape(-1, 8))
    vals = [kr.poly_eval(p, z) for z in points]
    for j, z in enumerate(points):
        want = bn.g1_mul(bn.G1_GEN, (kr.poly_eval(p, TAU) - vals[j]) * pow(TAU - z, -1, bn.R) % bn.R)
        assert g1_out(proofs[j], inf[j]) == want
    # verify: true for the right tuple, false for wrong point / value / proof / commitment (src/kzg.rs:310-468)
    m = len(points)
    com_xy = np.tile(L.g1_m(com), (m, 1))
    ok = ctx.verify_batch(com_xy, np.zeros(m, np.uint8), L.fr_vec(points).reshape(-1, 8), L.fr_vec(vals).reshape(-1, 8), proofs, inf)
    assert list(ok) == [1] * m
    bad_vals = [(v + 1) % bn.R for v in vals]
    ok = ctx.verify_batch(com_xy, np.zeros(m, np.uint8), L.fr_vec(points).reshape(-1, 8), L.fr_vec(bad_vals).reshape(-1, 8), proofs, inf)
    assert list(ok) == [0] * m
    bad_pts = [(z + 1) % bn.R for z in points]
    ok = ctx.verify_batch(com_xy, np.zeros(m, np.uint8), L.fr_vec(bad_pts).reshape(-1, 8), L.fr_vec(vals).reshape(-1, 8), proofs, inf)
    assert list(ok) == [0] * m
    ok = ctx.verify_batch(com_xy, np.zeros(m, np.uint8), L.fr_vec(points).reshape(-1, 8), L.fr_vec(vals).reshape(-1, 8),
                          np.roll(proofs, 1, axis=0), np.roll(inf, 1))
    assert list(ok) == [0] * m
    other = np.tile(L.g1_m(bn.g1_mul(com, 2)), (m, 1))
    ok = ctx.verify_batch(other, np.zeros(m, np.uint8), L.fr_vec(points).reshape(-1, 8), L.fr_vec(vals).reshape(-1, 8), proofs, inf)
    assert list(ok) == [0] * m
    # constant polynomial: proof is the identity
    proofs, inf = ctx.open_batch(L.fr_vec([42]).reshape(1, 8), L.fr_vec([3]).reshape(1, 8))
    assert inf[0] == 1


@pytest.mark.parametrize("d", [1, 2, 4, 8, 16, 32, 64, 256])
def test_open_all_fk(ctx, srs, d):
    p = [rng.randrange(bn.R) for _ in range(d)]
    proofs, inf = ctx.open_all_fk(L.fr_vec(p).reshape(d, 8))
    ptau = kr.poly_eval(p, TAU)
    for i, w in enumerate(bn.Radix2Domain(d).elements()):
        want = bn.g1_mul(bn.G1_GEN, (ptau - kr.poly_eval(p, w)) * pow(TAU - w, -1, bn.R) % bn.R)
        assert g1_out(proofs[i], inf[i]) == want, f"proof {i}"
    if d == 8:  # open_fk[i] == open(w^i)  (src/kzg.rs:470-505) through the other entry point
        pr2, inf2 = ctx.open_batch(L.fr_vec(p).reshape(d, 8), L.fr_vec(bn.Radix2Domain(d).elements()).reshape(-1, 8))
        assert np.array_equal(pr2, proofs) and np.array_equal(inf2, inf)


def test_encrypt_decrypt_bit_exact(ctx, srs):
    d = 16
    p, com = _setup_small(ctx, d)
    setup = kr.KZGSetup([], bn.g2_mul(bn.G2_GEN, TAU))
    n = 12
    points = [rng.randrange(bn.R) for _ in range(n)]
    points[0] = 0
    values = [kr.poly_eval(p, z) for z in points]
    values[1] = 0                       # wrong value (and the value = 0 edge: beta*G1 = identity)
    values[9] = 1                       # value = 1 takes the second per-commitment table (A / gT)^r, value = 2 the general form
    values[10] = 2
    rs = [rng.randrange(bn.R) for _ in range(n)]
    rs[2] = 0                           # r = 0: ciphertext is the identity, key = H(1)
    rs[3] = 1
    lens = [32, 32, 32, 0, 1, 63, 64, 65, 200, 32, 32, 7]
    msgs = [bytes(rng.randrange(256) for _ in range(k)) for k in lens]
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    flat = np.frombuffer(b"".join(msgs), np.uint8).copy()
    ct, ct_inf, msg_ct = ctx.encrypt_batch(L.g1_m(com), 0, L.fr_vec(points).reshape(n, 8), L.fr_vec(values).reshape(n, 8),
                                           L.fr_vec(rs).reshape(n, 8), flat, off)
    want = kr.vec_encrypt(rs, setup, com, points, values, msgs)
    for i in range(n):
        got_ct = None if ct_inf[i] else L.g2_from(ct[i])
        assert got_ct == want[i][0], f"ct {i}"
        assert bytes(msg_ct[int(off[i]): int(off[i + 1])]) == want[i][1], f"msg_ct {i}"
    assert ct_inf[2] == 1
    # decrypt with true proofs (trapdoor) -> original messages where the value was right
    ptau = kr.poly_eval(p, TAU)
    proofs = [bn.g1_mul(bn.G1_GEN, (ptau - kr.poly_eval(p, z)) * pow(TAU - z, -1, bn.R) % bn.R) for z in points]
    pinf = np.array([1 if q is None else 0 for q in proofs], np.uint8)
    out = ctx.decrypt_batch(L.g1_vec(proofs).reshape(n, 16), pinf, ct, ct_inf, msg_ct, off)
    ref = kr.vec_decrypt(proofs, want)
    for i in range(n):
        got = bytes(out[int(off[i]): int(off[i + 1])])
        assert got == ref[i], f"dec {i}"
        if i not in (1, 9, 10):
            assert got == msgs[i]
    assert bytes(out[int(off[1]): int(off[2])]) != msgs[1]   # wrong value -> garbage (src/enc.rs:99-125)
    # commitment at infinity and commitment == value*G1 (com_beta = identity): key = H(1)
    ct2, ci2, mc2 = ctx.encrypt_batch(np.zeros(16, np.uint32), 1, L.fr_vec([3]).reshape(1, 8), L.fr_vec([0]).reshape(1, 8),
                                      L.fr_vec([9]).reshape(1, 8), np.zeros(32, np.uint8), np.array([0, 32], np.uint64))
    assert bytes(mc2[:32]).hex() == "207d2aaa3257b30b7c371b6804480c9b2a7a04b4f69847270c5aadf5e5bc9454"
    vG = bn.g1_mul(bn.G1_GEN, 77)
    ct3, ci3, mc3 = ctx.encrypt_batch(L.g1_m(vG), 0, L.fr_vec([3]).reshape(1, 8), L.fr_vec([77]).reshape(1, 8),
                                      L.fr_vec([9]).reshape(1, 8), np.zeros(32, np.uint8), np.array([0, 32], np.uint64))
    assert bytes(mc3[:32]).hex() == "207d2aaa3257b30b7c371b6804480c9b2a7a04b4f69847270c5aadf5e5bc9454"


def test_encrypt_decrypt_roundtrip_larger(ctx, srs):
    """size-independent property: dec(enc(m)) == m for every index (2^10 here; bench sizes in test_gpu_full.py)"""
    n = 1 << 10
    d = n
    p = [rng.randrange(bn.R) for _ in range(d)]
    P = L.fr_vec(p).reshape(d, 8)
    com_xy, com_inf = ctx.msm_g1(P)
    proofs, pinf = ctx.open_all_fk(P)
    points = bn.Radix2Domain(n).elements()
    vals = [kr.poly_eval(p, z) for z in points[:8]]
    Pe = P.copy()
    ctx.fr_ntt(Pe)                      # evaluations at the domain = the values being opened
    assert L.fr_vec_from(Pe[:8].reshape(-1)) == vals
    rs = np.stack([L.fr_m(rng.randrange(bn.R)) for _ in range(n)])
    msgs = np.frombuffer(bytes(rng.randrange(256) for _ in range(32 * n)), np.uint8).copy()
    off = (np.arange(n + 1, dtype=np.uint64) * 32)
    ct, ct_inf, msg_ct = ctx.encrypt_batch(com_xy, com_inf, L.fr_vec(points).reshape(n, 8), Pe, rs, msgs, off)
    out = ctx.decrypt_batch(proofs, pinf, ct, ct_inf, msg_ct, off)
    assert np.array_equal(out[: 32 * n], msgs)


def test_msm_skewed_scalars_one_bucket_per_window(ctx, srs):
    """every scalar equal: each window puts all its entries into ONE bucket (the worst case for the one-thread-per-bucket
    accumulation and for the size-ordered schedule); commit(k, k, ..., k) = k * sum_i [tau^i]_1 = k (tau^n - 1)/(tau - 1) G1"""
    n = 1 << 13   # the fixture's SRS length
    for k in (bn.R - 1, 0x0FEDCBA987654321FEDCBA987654321FEDCBA987654321FEDCBA98765432 % bn.R):
        sc = np.tile(L.fr_m(k), (n, 1))
        xy, inf = ctx.msm_g1(np.ascontiguousarray(sc), n=n)
        want = bn.g1_mul(bn.G1_GEN, k * (pow(TAU, n, bn.R) - 1) * pow(TAU - 1, -1, bn.R) % bn.R)
        assert (None if inf else L.g1_from(xy)) == want
