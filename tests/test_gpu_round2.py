"""GPU parity tests added in round 2: the c = 16 window path of the MSM at the BASELINE config-2
size, special scalars at 2^16 / 2^20, SRS with repeated and negated bases (the P = +-Q branches of the mixed addition
inside the bucket accumulation), infinity operands inside a large decrypt batch, the full 2^16 batch of BASELINE
config 4 against the C oracle, SRS validation.  Bit-exact: integer / byte work."""
import os
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import keaki_ref as kr
from tests import limbs as L

pytestmark = pytest.mark.gpu

rng = random.Random(0x726F756E6432)
nprng = np.random.default_rng(0x6B32)
TAU = rng.randrange(1, bn.R)
GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT_GT_ONE = "207d2aaa3257b30b7c371b6804480c9b2a7a04b4f69847270c5aadf5e5bc9454"   # BLAKE3(ser(GT::one))[..32]


@pytest.fixture(scope="module")
def ctx():
    from keaki_b200 import _ffi
    c = _ffi.Context(0)
    yield c
    c.close()


def rand_fr_limbs(n):
    """n uniformly random Montgomery images below 2^252 < r (any value below r is the image of some scalar)"""
    a = nprng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= 0x0FFFFFFF
    return a


def scalars_of(limbs):
    """Montgomery limbs (n, 8) -> python ints"""
    raw = np.ascontiguousarray(limbs, np.uint32).tobytes()
    rinv = pow(1 << 256, -1, bn.R)
    return [int.from_bytes(raw[32 * i: 32 * i + 32], "little") * rinv % bn.R for i in range(limbs.shape[0])]


def trapdoor(scalars, first=0):
    acc = 0
    for s in reversed(scalars):   # Horner
        acc = (acc * TAU + s) % bn.R
    return bn.g1_mul(bn.G1_GEN, acc * pow(TAU, first, bn.R) % bn.R)


def g1_out(xy, inf):
    return None if inf else L.g1_from(xy)


# ---------------------------------------------------------------------------------------------
# MSM: the c = 16 path (2^13 < n <= 2^17), BASELINE config 2 is n = 2^16
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def srs17(ctx):
    n = 1 << 17
    g1, _ = ctx.srs_generate(L.fr_m(TAU), n)
    return g1


@pytest.mark.parametrize("logn", [14, 16, 17])
def test_msm_c16_path_matches_trapdoor(ctx, srs17, logn):
    n = 1 << logn
    S = rand_fr_limbs(n)
    sc = scalars_of(S)
    xy, inf = ctx.msm_g1(S)
    assert g1_out(xy, inf) == trapdoor(sc)
    # n not a power of two inside the same window class, and the point-range shard form (first > 0)
    m = n - 12345 if n > 20000 else n - 3
    xy, inf = ctx.msm_g1(np.ascontiguousarray(S[:m]))
    assert g1_out(xy, inf) == trapdoor(sc[:m])
    first = (1 << 17) - m
    xy, inf = ctx.msm_g1(np.ascontiguousarray(S[:m]), first=first)
    assert g1_out(xy, inf) == trapdoor(sc[:m], first=first)


def test_msm_2_14_matches_c_oracle(ctx, srs17):
    """the C restatement of arkworks' Pippenger on the same bases and scalars (a second, independent summation order)"""
    from oracle import coracle as co
    n = 1 << 14
    S = rand_fr_limbs(n)
    xy, inf = ctx.msm_g1(S)
    want_xy, want_inf = co.msm_g1(np.ascontiguousarray(srs17[:n]), S, threads=4)
    assert inf == want_inf and np.array_equal(np.asarray(xy, np.uint32), np.asarray(want_xy, np.uint32).reshape(16))


@pytest.mark.parametrize("logn", [16, 20])
def test_msm_special_scalars(ctx, logn):
    """scalars drawn from {0, 1, r - 1, random}: zero digits, the carry chain of the signed recoding at its extreme
    (r - 1), and buckets that see the same point in several windows"""
    n = 1 << logn
    ctx.srs_generate(L.fr_m(TAU), n, download=False)
    pool = [0, 1, bn.R - 1, 2, bn.R - 2, (1 << 253) % bn.R]
    pool_l = np.stack([L.fr_m(v) for v in pool])
    pick = nprng.integers(0, len(pool) + 2, size=n)
    S = rand_fr_limbs(n)
    special = pick < len(pool)
    S[special] = pool_l[pick[special]]
    sc = scalars_of(S)
    xy, inf = ctx.msm_g1(S)
    assert g1_out(xy, inf) == trapdoor(sc)
    # only special values (every bucket population is extreme: 1/3 of the entries are zero digits)
    S2 = pool_l[nprng.integers(0, 3, size=n)]
    xy, inf = ctx.msm_g1(np.ascontiguousarray(S2))
    assert g1_out(xy, inf) == trapdoor(scalars_of(S2))


@pytest.mark.parametrize("n", [96, 3000])
def test_msm_repeated_and_negated_bases(ctx, n):
    """an uploaded SRS that repeats and negates points, with equal scalars on them: inside one bucket the accumulator
    meets the point it already holds (doubling branch of the mixed addition) and its negative (infinity branch)"""
    base = [bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R)) for _ in range(8)]
    pts, scalars = [], []
    for i in range(n):
        k = i % 24
        p = base[(k // 3) % 8]
        if k % 3 == 2:
            p = bn.g1_neg(p)
        pts.append(p)
        scalars.append(0)
    s_shared = [rng.randrange(bn.R) for _ in range(n // 3 + 1)]
    for i in range(n):
        # triples (P, P, -P) share one scalar -> contribute s * P in total; every 5th triple uses different scalars
        scalars[i] = s_shared[i // 3] if (i // 3) % 5 else rng.randrange(bn.R)
    tau2 = L.g2_m(bn.g2_mul(bn.G2_GEN, TAU))
    ctx.srs_upload(L.g1_vec(pts).reshape(n, 16), tau2)
    xy, inf = ctx.msm_g1(L.fr_vec(scalars).reshape(n, 8))
    assert g1_out(xy, inf) == bn.g1_msm(pts, scalars)
    # all triples cancel pairwise: (P, -P) with equal scalars only -> identity
    pts2 = [base[i // 2 % 8] if i % 2 == 0 else bn.g1_neg(base[i // 2 % 8]) for i in range(n - n % 2)]
    sc2 = [s_shared[(i // 2) % 7] for i in range(len(pts2))]
    ctx.srs_upload(L.g1_vec(pts2).reshape(len(pts2), 16), tau2)
    xy, inf = ctx.msm_g1(L.fr_vec(sc2).reshape(len(pts2), 8))
    assert inf == 1


# ---------------------------------------------------------------------------------------------
# witness encryption
# ---------------------------------------------------------------------------------------------
def test_decrypt_large_batch_with_infinity_operands(ctx):
    """proof = infinity / ct = infinity inside a batch large enough to fill whole warps: those lanes must give
    key = H(ser(GT::one)) (arkworks skips pairs with an infinity) while their neighbours are untouched"""
    n = 4096 + 37
    ctx.srs_generate(L.fr_m(TAU), 64, download=False)
    d = 64
    p = [rng.randrange(bn.R) for _ in range(d)]
    com = trapdoor(p)
    pts = rand_fr_limbs(n)
    vals = rand_fr_limbs(n)
    rs = rand_fr_limbs(n)
    msgs = nprng.integers(0, 256, size=32 * n, dtype=np.uint8)
    off = (np.arange(n + 1, dtype=np.uint64) * 32)
    ct, ct_inf, msg_ct = ctx.encrypt_batch(L.g1_m(com), 0, pts, vals, rs, msgs, off)
    proofs, pinf = ctx.g1_mul_gen_batch(rand_fr_limbs(n))
    base = ctx.decrypt_batch(proofs, pinf, ct, ct_inf, msg_ct, off).copy()
    pinf2, cinf2 = pinf.copy(), ct_inf.copy()
    special_p = [0, 31, 32, 1000, 4096, n - 1]
    special_c = [5, 63, 2048, n - 2]
    pinf2[special_p] = 1
    cinf2[special_c] = 1
    proofs2 = proofs.copy()
    proofs2[1234] = 0                      # all-zero coordinates without the flag are infinity too
    out = ctx.decrypt_batch(proofs2, pinf2, ct, cinf2, msg_ct, off)
    kat = np.frombuffer(bytes.fromhex(KAT_GT_ONE), np.uint8)
    touched = set(special_p + special_c + [1234])
    for i in touched:
        assert np.array_equal(out[32 * i: 32 * i + 32], msg_ct[32 * i: 32 * i + 32] ^ kat), i
    mask = np.ones(n, bool)
    mask[list(touched)] = False
    assert np.array_equal(out[: 32 * n].reshape(n, 32)[mask], base[: 32 * n].reshape(n, 32)[mask])


def test_config4_full_batch_matches_c_oracle(ctx):
    """BASELINE config 4 at full size: 2^16 messages of 32 B, points = the 2^16-th roots of unity, values in {0, 1};
    EVERY ciphertext and masked message against the C restatement of the reference path, then the GPU decryption of all
    of them with true openings returns every message"""
    from oracle import coracle as co
    from keaki_b200.types import Radix2EvaluationDomain
    logn = int(os.environ.get("KB_CFG4_LOG_N", "16"))
    n = 1 << logn
    ctx.srs_generate(L.fr_m(TAU), n, download=False)
    # polynomial through bit values on the domain: coefficients = ifft(bits)
    bits = nprng.integers(0, 2, size=n)
    one, zero = L.fr_m(1), L.fr_m(0)
    evals = np.ascontiguousarray(np.where(bits[:, None] == 1, one[None, :], zero[None, :]).astype(np.uint32))
    coeffs = evals.copy()
    ctx.fr_ntt(coeffs, inverse=True)
    com_xy, com_inf = ctx.msm_g1(coeffs)
    points = Radix2EvaluationDomain(n).elements_limbs()
    rs = rand_fr_limbs(n)
    msgs = nprng.integers(0, 256, size=32 * n, dtype=np.uint8)
    off = (np.arange(n + 1, dtype=np.uint64) * 32)
    ct, ct_inf, msg_ct = ctx.encrypt_batch(com_xy, com_inf, points, evals, rs, msgs, off)
    tau2 = L.g2_m(bn.g2_mul(bn.G2_GEN, TAU))
    threads = max(1, len(os.sched_getaffinity(0)))
    w_ct, w_inf, w_mc = co.encrypt_batch(np.asarray(com_xy, np.uint32), int(com_inf), tau2, points, evals, rs, msgs, off, threads=threads)
    assert np.array_equal(ct_inf, w_inf)
    assert np.array_equal(ct, np.asarray(w_ct, np.uint32).reshape(n, 32))
    assert np.array_equal(msg_ct[: 32 * n], np.asarray(w_mc, np.uint8)[: 32 * n])
    # decrypt all of them with the FK openings of the committed polynomial
    proofs, pinf = ctx.open_all_fk(coeffs)
    out = ctx.decrypt_batch(proofs, pinf, ct, ct_inf, msg_ct, off)
    assert np.array_equal(out[: 32 * n], msgs)


# ---------------------------------------------------------------------------------------------
# SRS validation
# ---------------------------------------------------------------------------------------------
def test_srs_validate(ctx, tmp_path):
    from keaki_b200 import _ffi, kzg, ptau
    n = 1 << 12
    g1, tau2 = ctx.srs_generate(L.fr_m(TAU), n)
    ctx.srs_validate()                                   # a generated SRS is valid
    bad = g1.copy()
    bad[777, 3] ^= 1                                     # one flipped bit in one coordinate
    bad[3000, 9] ^= 0x100
    ctx.srs_upload(bad, tau2)
    with pytest.raises(_ffi.InvalidSrsPoint) as e:
        ctx.srs_validate()
    assert e.value.index == 777
    inf_pt = g1.copy()
    inf_pt[5] = 0                                        # (0, 0) = the affine encoding of infinity: not an SRS element
    ctx.srs_upload(inf_pt, tau2)
    with pytest.raises(_ffi.InvalidSrsPoint) as e:
        ctx.srs_validate()
    assert e.value.index == 5
    big = g1.copy()
    big[9, :8] = L.int_to_limbs(bn.Q + 5)                # coordinate limbs not below q
    ctx.srs_upload(big, tau2)
    with pytest.raises(_ffi.InvalidSrsPoint) as e:
        ctx.srs_validate()
    assert e.value.index == 9
    # [tau]_2: off the twist, and on the twist but outside the r-torsion (cofactor component)
    t2 = tau2.copy()
    t2[0] ^= 1
    ctx.srs_upload(g1, t2)
    with pytest.raises(_ffi.InvalidSrsPoint) as e:
        ctx.srs_validate()
    assert e.value.index == n
    x = (rng.randrange(bn.Q), rng.randrange(bn.Q))
    while True:                                          # a random point of the twist: in the r-torsion with probability 1/h2
        y = bn.f2_sqrt(bn.f2_add(bn.f2_mul(bn.f2_sqr(x), x), bn.B2))
        if y is not None and not bn.g2_in_subgroup((x, y)):
            break
        x = (x[0] + 1, x[1])
    ctx.srs_upload(g1, L.g2_m((x, y)))
    with pytest.raises(_ffi.InvalidSrsPoint) as e:
        ctx.srs_validate()
    assert e.value.index == n
    # the reference's own fixture: valid when its limbs are taken as the Montgomery limbs they are ...
    f1, f2 = ptau.get_powers_from_file(os.path.join(GOLD, "ppot_0080_01_mini.ptau"))
    ctx.srs_upload(f1, f2[1])
    ctx.srs_validate()
    setup = kzg.KZGSetup.new_from_file(os.path.join(GOLD, "ppot_0080_01_mini.ptau"), ctx=ctx)
    assert len(setup) == 3
    # ... and INVALID when they are read as canonical integers, which is what the reference's
    # `deserialize_uncompressed_unchecked` does (src/kzg/ptau.rs:266,314): its SRS is off the curve
    as_canonical = np.stack([np.concatenate([L.fq_m(L.limbs_to_int(p[:8]) % bn.Q), L.fq_m(L.limbs_to_int(p[8:]) % bn.Q)]) for p in f1])
    ctx.srs_upload(as_canonical, f2[1])
    with pytest.raises(_ffi.InvalidSrsPoint) as e:
        ctx.srs_validate()
    assert e.value.index == 0
    # a file with one corrupted element is a SetupFileError from new_from_file
    corrupt = f1.copy()
    corrupt[2, 0] ^= 4
    path = tmp_path / "corrupt.ptau"
    ptau.write_ptau(str(path), corrupt, f2, power=1)
    with pytest.raises(ptau.SetupFileError):
        kzg.KZGSetup.new_from_file(str(path), ctx=ctx)
    assert len(kzg.KZGSetup.new_from_file(str(path), ctx=ctx, validate=False)) == 3


def test_msm_structured_scalars_do_not_serialise(ctx):
    """all windows share one bucket set, so all-equal / 0-1 / small scalars put O(n) entries into a few
    buckets.  Over-full buckets are summed in segments by many threads: the result is exact and a 2^20 commit of such a
    polynomial stays within a small multiple of the uniform case instead of one thread running 2^20 additions."""
    n = 1 << 20
    ctx.srs_generate(L.fr_m(TAU), n, download=False)
    geo = (pow(TAU, n, bn.R) - 1) * pow(TAU - 1, -1, bn.R) % bn.R          # sum_i tau^i
    uniform = rand_fr_limbs(n)
    ctx.msm_g1(uniform)
    ctx.msm_g1(uniform)
    t_uniform = ctx.last_kernel_ms(0)
    for k in (1, 7, bn.R - 1, 0x0FEDCBA987654321FEDCBA987654321FEDCBA987654321FEDCBA98765432 % bn.R):
        sc = np.ascontiguousarray(np.tile(L.fr_m(k), (n, 1)))
        xy, inf = ctx.msm_g1(sc)
        ms = ctx.last_kernel_ms(0)
        assert g1_out(xy, inf) == bn.g1_mul(bn.G1_GEN, k * geo % bn.R)
        assert ms < 12 * t_uniform + 5.0, f"all-equal scalars {k:#x}: {ms:.1f} ms vs {t_uniform:.1f} ms uniform"
    # 0/1 vector (a bit-vector polynomial, what a laconic-OT receiver commits to before the inverse transform)
    bits = nprng.integers(0, 2, size=n)
    sc = np.ascontiguousarray(np.where(bits[:, None] == 1, L.fr_m(1)[None, :], L.fr_m(0)[None, :]).astype(np.uint32))
    xy, inf = ctx.msm_g1(sc)
    ms = ctx.last_kernel_ms(0)
    acc, t = 0, 1
    tau_pows = None
    # sum over set bits of tau^i: Horner over the bit list
    acc = 0
    for b in reversed(bits.tolist()):
        acc = (acc * TAU + b) % bn.R
    assert g1_out(xy, inf) == bn.g1_mul(bn.G1_GEN, acc)
    assert ms < 12 * t_uniform + 5.0, f"0/1 scalars: {ms:.1f} ms vs {t_uniform:.1f} ms uniform"


# ---------------------------------------------------------------------------------------------
# FK open-all: merged radix-8 passes of the small G1 transforms (poly.cu) against one stage per launch and the trapdoor
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("logd", [3, 9, 13, 14])
def test_open_all_fk_radix8_passes(ctx, logd):
    """2d = 2^(logd+1) points: logd = 13 ends on a radix-4 pass (14 = 3+3+3+3+2), logd = 14 sends the 2d transform down the
    large-domain path and the d transform through radix-8 passes; every proof equals the radix-2 build's, a few equal the
    trapdoor's [(p(tau) - p(w^i)) / (tau - w^i)] G1"""
    from keaki_b200 import _ffi
    from keaki_b200.types import Radix2EvaluationDomain
    d = 1 << logd
    ctx.srs_generate(L.fr_m(TAU), d, download=False)
    coeffs = rand_fr_limbs(d)
    proofs, pinf = ctx.open_all_fk(coeffs)
    os.environ["KB_NTT_RADIX2"] = "1"
    try:
        c2 = _ffi.Context(0)
    finally:
        os.environ.pop("KB_NTT_RADIX2", None)
    try:
        c2.srs_generate(L.fr_m(TAU), d, download=False)
        p2, i2 = c2.open_all_fk(coeffs)
    finally:
        c2.close()
    assert np.array_equal(pinf, i2) and np.array_equal(proofs, p2)
    p = scalars_of(coeffs)
    dom = Radix2EvaluationDomain(d).elements()

    def ev(z):
        acc = 0
        for c in reversed(p):
            acc = (acc * z + c) % bn.R
        return acc
    ptau = ev(TAU)
    for i in sorted({0, 1, d // 2, d - 1, rng.randrange(d)}):
        w = int(dom[i])
        want = bn.g1_mul(bn.G1_GEN, (ptau - ev(w)) * pow(TAU - w, -1, bn.R) % bn.R)
        assert g1_out(proofs[i], pinf[i]) == want, i
