"""The C restatement of the reference CPU path (oracle/c) against the Python big-int oracle."""
import random

import blake3
import numpy as np

from oracle import bn254 as bn
from oracle import coracle as co
from oracle import keaki_ref as kr
from tests import limbs as L

rng = random.Random(2024264)


def test_blake3_matches_reference_implementation():
    for n in (0, 1, 63, 64, 65, 384, 1000, 1024):
        data = bytes(rng.randrange(256) for _ in range(n))
        for out_len in (0, 1, 32, 64, 65, 200):
            assert co.blake3_xof(data, out_len) == blake3.blake3(data).digest(out_len)
    one = bn.gt_to_bytes(bn.F12_ONE)
    assert co.blake3_xof(one, 32).hex() == "207d2aaa3257b30b7c371b6804480c9b2a7a04b4f69847270c5aadf5e5bc9454"


def test_g1_mul_and_msm():
    tau = rng.randrange(1, bn.R)
    for n in (1, 5, 31, 32, 33, 200):
        pts = [bn.g1_mul(bn.G1_GEN, pow(tau, i, bn.R)) for i in range(n)]
        sc = [rng.randrange(bn.R) for _ in range(n)]
        if n >= 5:
            sc[0], sc[1], sc[2] = 0, 1, bn.R - 1
        xy, inf = co.msm_g1(L.g1_vec(pts).reshape(n, 16), L.fr_vec(sc).reshape(n, 8), threads=1)
        want = bn.g1_mul(bn.G1_GEN, sum(s * pow(tau, i, bn.R) for i, s in enumerate(sc)) % bn.R)
        assert (None if inf else L.g1_from(xy)) == want
        xy2, inf2 = co.msm_g1(L.g1_vec(pts).reshape(n, 16), L.fr_vec(sc).reshape(n, 8), threads=4)
        assert np.array_equal(xy, xy2) and inf == inf2
    p = bn.g1_mul(bn.G1_GEN, 12345)
    for k in (0, 1, 2, bn.R - 1, rng.randrange(bn.R)):
        xy, inf = co.g1_mul(L.g1_m(p), 0, L.fr_m(k))
        assert (None if inf else L.g1_from(xy)) == bn.g1_mul(p, k)


def test_pairing_bytes():
    cases = [(bn.G1_GEN, bn.G2_GEN), (None, bn.G2_GEN), (bn.G1_GEN, None)]
    for _ in range(3):
        cases.append((bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R)), bn.g2_mul(bn.G2_GEN, rng.randrange(1, bn.R))))
    g1 = L.g1_vec([c[0] for c in cases]).reshape(-1, 16)
    g2 = L.g2_vec([c[1] for c in cases]).reshape(-1, 32)
    i1 = np.array([c[0] is None for c in cases], np.uint8)
    i2 = np.array([c[1] is None for c in cases], np.uint8)
    out = co.pairing_batch(g1, i1, g2, i2, threads=2)
    for k, (p, q) in enumerate(cases):
        assert bytes(out[k]) == bn.gt_to_bytes(bn.pairing(p, q))


def test_encrypt_decrypt_match_python_oracle():
    tau = rng.randrange(1, bn.R)
    setup = kr.KZGSetup.setup(tau, 4)
    p = [rng.randrange(bn.R) for _ in range(4)]
    com = kr.commit(setup, p)
    n = 5
    points = [rng.randrange(bn.R) for _ in range(n)]
    values = [kr.poly_eval(p, z) for z in points]
    values[1] = 0
    rs = [rng.randrange(bn.R) for _ in range(n)]
    rs[2] = 0
    lens = [32, 0, 65, 32, 7]
    msgs = [bytes(rng.randrange(256) for _ in range(k)) for k in lens]
    off = np.zeros(n + 1, np.uint64); off[1:] = np.cumsum(lens)
    flat = np.frombuffer(b"".join(msgs), np.uint8).copy()
    ct, ct_inf, msg_ct = co.encrypt_batch(L.g1_m(com), 0, L.g2_m(setup.tau_g2), L.fr_vec(points).reshape(n, 8),
                                          L.fr_vec(values).reshape(n, 8), L.fr_vec(rs).reshape(n, 8), flat, off, threads=2)
    want = kr.vec_encrypt(rs, setup, com, points, values, msgs)
    for i in range(n):
        assert (None if ct_inf[i] else L.g2_from(ct[i])) == want[i][0]
        assert bytes(msg_ct[int(off[i]): int(off[i + 1])]) == want[i][1]
    proofs = [kr.open(setup, p, z) for z in points]
    pinf = np.array([q is None for q in proofs], np.uint8)
    out = co.decrypt_batch(L.g1_vec(proofs).reshape(n, 16), pinf, ct, ct_inf, msg_ct, off, threads=2)
    ref = kr.vec_decrypt(proofs, want)
    for i in range(n):
        assert bytes(out[int(off[i]): int(off[i + 1])]) == ref[i]
        if i != 1:
            assert ref[i] == msgs[i]


def test_product_counter_counts_single_thread_work():
    """The instrumentation behind DESIGN.md's exact per-unit work table: counts are deterministic and scale with n."""
    import numpy as np
    rng = np.random.default_rng(3)
    pts = [bn.g1_mul(bn.G1_GEN, k) for k in (1, 2, 3, 5)]
    bases = np.tile(L.g1_vec(pts).reshape(4, 16), (64, 1))
    sc = rng.integers(0, 1 << 32, size=(256, 8), dtype=np.uint64).astype(np.uint32); sc[:, 7] &= 0x0FFFFFFF
    co.count_muls(True)
    co.msm_g1(bases, sc, threads=1)
    c1 = co.count_muls(True)
    co.msm_g1(bases, sc, threads=1)
    c2 = co.count_muls(False)
    co.msm_g1(bases, sc, threads=1)
    assert c1 == c2 > 256 * 50 and co.count_muls(False) == 0
