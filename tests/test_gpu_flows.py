"""GPU tests through the reference-facing mirror (keaki_b200.kzg / kem / enc / vec / laconic_ot): the
reference's own tests on BN254, BASELINE.json configs 1 and 3, the committed golden vectors, and
size-independent properties at the bench sizes."""
import json
import os
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import keaki_ref as kr
from tests import limbs as L

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
rng = random.Random(0xC0FFEE)


@pytest.fixture(scope="module")
def ctx():
    from keaki_b200 import Context
    c = Context(0)
    yield c
    c.close()


def g1o(p):
    return p.to_affine_ints()


def to_g1(p):
    from keaki_b200 import G1
    return G1(None, True) if p is None else G1(L.g1_m(p))


# ------------------------------------------------------------------ reference unit tests on the GPU path
def test_reference_kzg_suite(ctx):
    from keaki_b200 import kzg, FrRng
    secret = FrRng(1).fr()
    s = kzg.KZGSetup.setup(secret, 4, ctx=ctx)                      # src/kzg.rs:218-239
    for i, p in enumerate(s.g1_pow()):
        assert g1o(p) == bn.g1_mul(bn.G1_GEN, pow(secret, i, bn.R))
    assert s.tau_g2().to_affine_ints() == bn.g2_mul(bn.G2_GEN, secret)
    ref = kr.KZGSetup.setup(secret, 4)
    p = [1, 3, 2]
    com = kzg.commit(s, p)                                          # :241-258
    assert g1o(com) == kr.commit(ref, p)
    proof = kzg.open(s, p, 5)                                       # :310-331 (p(5) = 66)
    assert g1o(proof) == kr.open(ref, p, 5)
    assert kzg.verify(s, com, 5, 66, proof)
    assert not kzg.verify(s, com, 6, 66, proof) and not kzg.verify(s, com, 5, 67, proof)
    assert not kzg.verify(s, com, 5, 66, kzg.open(s, p, 6)) and not kzg.verify(s, kzg.commit(s, [1, 2]), 5, 66, proof)
    s2 = kzg.KZGSetup.setup(secret, 2, ctx=ctx)
    with pytest.raises(kzg.KZGError) as e:                          # :260-277
        kzg.commit(s2, [1, 3, 2, 4])
    assert "PolynomialTooLarge(4, 2)" in str(e.value)
    with pytest.raises(kzg.KZGError) as e:                          # :279-308 (quotient.len() = 5)
        kzg.open(s2, [1, 2, 3, 4, 5, 6], 5)
    assert "PolynomialTooLarge(5, 2)" in str(e.value)


def test_reference_open_fk_equals_open(ctx):                         # src/kzg.rs:470-505
    from keaki_b200 import kzg, FrRng, Radix2EvaluationDomain
    r = FrRng(2)
    d = 16
    s = kzg.KZGSetup.setup(r.fr(), d, ctx=ctx)
    p = [r.fr() for _ in range(d)]
    dom = Radix2EvaluationDomain(d)
    fk = kzg.open_fk(s, p, dom)
    assert fk == kzg.open_many(s, p, dom.elements())
    assert fk[3] == kzg.open(s, p, dom.elements()[3])


def test_reference_kem_and_enc_suite(ctx):                           # src/kem.rs:87-224, src/enc.rs:70-125
    from keaki_b200 import kzg, kem, enc, FrRng, G2
    r = FrRng(3)
    s = kzg.KZGSetup.setup(r.fr(), 10, ctx=ctx)
    p = [1, 2, 3, 4, 5, 6, 7]
    com = kzg.commit(s, p)
    point, value = 11, kr.poly_eval(p, 11)
    ct, key = kem.encapsulate(r, s, com, point, value, 32)
    assert len(key) == 32 and kem.decapsulate(kzg.open(s, p, point), ct, 32, ctx=ctx) == key
    assert kem.decapsulate(kzg.open(s, [7, 6, 5, 4, 3, 2, 1], point), ct, 32, ctx=ctx) != key
    assert kem.decapsulate(kzg.open(s, p, 12), ct, 32, ctx=ctx) != key
    ct2 = bn.g2_mul(ct.to_affine_ints(), 2)
    assert kem.decapsulate(kzg.open(s, p, point), G2(L.g2_m(ct2)), 32, ctx=ctx) != key
    c = enc.encrypt(r, s, com, point, value, b"helloworld")
    assert enc.decrypt(kzg.open(s, p, point), c, ctx=ctx) == b"helloworld"
    assert enc.decrypt(kzg.open(s, p, 12), c, ctx=ctx) != b"helloworld"


def test_reference_laconic_ot(ctx):                                   # tests/laconic_ot.rs:127-200
    from keaki_b200 import kzg, FrRng
    from keaki_b200.laconic_ot import Receiver, Sender
    r = FrRng(4)
    n = 8
    s = kzg.KZGSetup.setup(r.fr(), 16, ctx=ctx)
    choices = [r.fr() & 1 for _ in range(n)]
    rcv = Receiver(s, r, choices)
    snd = Sender(s, rcv.commitment)
    values = [[r.bytes(32) for _ in range(n)] for _ in range(2)]
    out = rcv.receive(snd.send(r, values))
    assert out == [values[choices[i]][i] for i in range(n)]


# ------------------------------------------------------------------ BASELINE config 1: laconic OT on the ptau SRS
def test_config1_laconic_ot_on_ptau_srs_bit_exact(ctx):
    from keaki_b200 import kzg, FrRng
    from keaki_b200.laconic_ot import Receiver, Sender
    path = os.path.join(GOLD, "ppot_0080_01_mini.ptau")
    s = kzg.KZGSetup.new_from_file(path, ctx=ctx)
    assert len(s) == 3                                               # src/kzg/ptau.rs:476-494
    ref = kr.KZGSetup.new_from_file(path)
    assert [g1o(p) for p in s.g1_pow()] == ref.g1_aff and s.tau_g2().to_affine_ints() == ref.tau_g2
    for choice in (0, 1):
        r1, r2 = FrRng(50 + choice), FrRng(50 + choice)
        rcv = Receiver(s, r1, [choice])                              # 1 choice + 1 pad -> domain of size 2
        ref_rcv = kr.Receiver(ref, r2.fr(), [choice])
        assert g1o(rcv.commitment) == ref_rcv.commitment
        assert [g1o(p) for p in rcv.proofs] == ref_rcv.proofs
        values = [[r1.bytes(32)], [r1.bytes(32)]]
        r2.bytes(64)
        enc = Sender(s, rcv.commitment).send(r1, values)
        ref_enc = kr.Sender(ref, ref_rcv.commitment).send([r2.fr()], [r2.fr()], values)
        for a in range(2):
            assert enc[a][0][0].to_affine_ints() == ref_enc[a][0][0] and enc[a][0][1] == ref_enc[a][0][1]
        assert rcv.receive(enc) == [values[choice][0]] == ref_rcv.receive(ref_enc)


# ------------------------------------------------------------------ committed golden vectors
def test_golden_vectors_on_gpu(ctx):
    from keaki_b200 import kzg, vec, G1, G2, Radix2EvaluationDomain
    v = json.load(open(os.path.join(GOLD, "oracle_vectors.json")))
    p = [int(c) for c in v["coeffs"]]
    s = kzg.KZGSetup.setup(int(v["tau"]), len(p), ctx=ctx)
    com = kzg.commit(s, p)
    assert [str(c) for c in g1o(com)] == v["commitment"]
    proofs = kzg.open_fk(s, p, Radix2EvaluationDomain(len(p)))
    assert [[str(c) for c in g1o(q)] for q in proofs] == v["proofs"]

    class FixedRng:
        def __init__(self, xs): self.xs = list(xs)
        def fr(self): return self.xs.pop(0)
    msgs = [bytes.fromhex(m) for m in v["messages"]]
    cts = vec.vec_encrypt(FixedRng(int(r) for r in v["r"]), s, com, [int(z) for z in v["points"]], [int(x) for x in v["values"]], msgs)
    for c, g in zip(cts, v["ciphertexts"]):
        a = c[0].to_affine_ints()
        assert [[str(a[0][0]), str(a[0][1])], [str(a[1][0]), str(a[1][1])]] == g["g2"] and c[1].hex() == g["msg_ct"]
    assert vec.vec_decrypt(proofs, cts, ctx=ctx) == msgs
    keys = vec.vec_decrypt(proofs, [(c[0], bytes(32)) for c in cts], ctx=ctx)
    assert [k.hex() for k in keys] == v["keys32"]
    gt = ctx.pairing_batch(L.g1_m(bn.g1_mul(bn.G1_GEN, 5)).reshape(1, 16), None, L.g2_m(bn.g2_mul(bn.G2_GEN, 7)).reshape(1, 32), None)
    assert bytes(gt[0]).hex() == v["gt_5_7"]
    # wire bytes of the proofs and ciphertext points
    pxy = np.stack([q.xy for q in proofs]); pinf = np.array([q.inf for q in proofs], np.uint8)
    cxy = np.stack([c[0].xy for c in cts]); cinf = np.array([c[0].inf for c in cts], np.uint8)
    for compress, kp, kc in ((True, "proofs_compressed", "ct_compressed"), (False, "proofs_uncompressed", "ct_uncompressed")):
        assert [bytes(b).hex() for b in ctx.g1_serialize(pxy, pinf, compress)] == v["wire"][kp]
        assert [bytes(b).hex() for b in ctx.g2_serialize(cxy, cinf, compress)] == v["wire"][kc]
        back, binf, ok = ctx.g2_deserialize(np.frombuffer(bytes.fromhex("".join(v["wire"][kc])), np.uint8), compress, True)
        assert ok.all() and np.array_equal(back, cxy) and np.array_equal(binf, cinf)


# ------------------------------------------------------------------ BASELINE config 3: vec open-all at 2^12
def test_config3_vec_commit_open_all_4096(ctx):
    from keaki_b200 import kzg, vec, FrRng
    r = FrRng(7)
    tau = r.fr()
    d = 1 << 12
    s = kzg.KZGSetup.setup(tau, d, ctx=ctx)
    vals = [r.fr() for _ in range(d - 1)]
    r_pad = FrRng(99)
    com, proofs = vec.vec_commit(r_pad, s, vals)
    padded = vals + [FrRng(99).fr()]
    dom = bn.Radix2Domain(d)
    p = dom.ifft(padded)
    ptau = kr.poly_eval(p, tau)
    assert g1o(com) == bn.g1_mul(bn.G1_GEN, ptau)                     # trapdoor: commit = p(tau) G1
    els = dom.elements()
    for i in [0, 1, 2, 17, 2048, 4094, 4095] + [rng.randrange(d) for _ in range(24)]:
        want = bn.g1_mul(bn.G1_GEN, (ptau - padded[i]) * pow(tau - els[i], -1, bn.R) % bn.R)
        assert g1o(proofs[i]) == want, f"proof {i}"
    # every proof verifies on the GPU (size-independent property), and a shifted one does not
    pxy = np.stack([q.xy for q in proofs]); pinf = np.array([q.inf for q in proofs], np.uint8)
    ok = ctx.verify_batch(np.tile(com.xy, (d, 1)), np.zeros(d, np.uint8), L.fr_vec(els).reshape(d, 8),
                          L.fr_vec(padded).reshape(d, 8), pxy, pinf)
    assert ok.all()
    ok = ctx.verify_batch(np.tile(com.xy, (8, 1)), np.zeros(8, np.uint8), L.fr_vec(els[:8]).reshape(8, 8),
                          L.fr_vec(padded[:8]).reshape(8, 8), np.roll(pxy[:8], 1, axis=0), pinf[:8])
    assert not ok.any()


# ------------------------------------------------------------------ bench sizes: size-independent properties
def test_full_size_msm_linearity_and_trapdoor(ctx):
    """2^20-point commit (BASELINE configs[1] at the metric's size): commit(a) + commit(b) == commit(a + b),
    commit(e_j) == srs[j], and the trapdoor identity on a structured polynomial."""
    n = 1 << 20
    tau = 0x1D2C3B4A5968778695A4B3C2D1E0F1E2D3C4B5A69788796A5B4C3D2E1F001122 % bn.R
    g1, _ = ctx.srs_generate(L.fr_m(tau), n, download=True)
    nprng = np.random.default_rng(5)
    a = nprng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); a[:, 7] &= 0x0FFFFFFF
    b = nprng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); b[:, 7] &= 0x0FFFFFFF
    # a + b limb-wise with carry (values < 2^252, so no modular wrap): do it with Python ints on a view
    ai = a.view(np.uint64).astype(object); bi = b.view(np.uint64).astype(object)
    s = np.zeros((n, 4), dtype=object); carry = np.zeros(n, dtype=object)
    for k in range(4):
        t = ai[:, k] + bi[:, k] + carry
        s[:, k] = t & ((1 << 64) - 1); carry = t >> 64
    ab = s.astype(np.uint64).view(np.uint32).reshape(n, 8)
    ca, cb, cab = ctx.msm_g1(a), ctx.msm_g1(b), ctx.msm_g1(np.ascontiguousarray(ab))
    sxy, sinf = ctx.g1_sum(np.stack([ca[0], cb[0]]), np.array([ca[1], cb[1]], np.uint8))
    assert np.array_equal(sxy, cab[0]) and sinf == cab[1] == 0
    # unit vector: Montgomery image of 1 at position j -> the SRS point itself
    for j in (0, 1, n - 1, 777777):
        e = np.zeros((n, 8), np.uint32); e[j] = L.fr_m(1)
        xy, inf = ctx.msm_g1(e)
        assert inf == 0 and np.array_equal(xy, g1[j])
    # geometric coefficients c_i = x^i: commit = (sum (x tau)^i) G1 = ((x tau)^n - 1)/(x tau - 1) G1
    x = 3
    pw = np.zeros((n, 8), np.uint32)
    acc = 1
    rows = []
    for i in range(n):
        rows.append((acc * bn.MONT_R % bn.R).to_bytes(32, "little")); acc = acc * x % bn.R
    pw = np.frombuffer(b"".join(rows), np.uint32).reshape(n, 8).copy()
    xy, inf = ctx.msm_g1(pw)
    xt = x * tau % bn.R
    want = bn.g1_mul(bn.G1_GEN, (pow(xt, n, bn.R) - 1) * pow(xt - 1, -1, bn.R) % bn.R)
    assert (None if inf else L.g1_from(xy)) == want


def test_full_size_we_roundtrip_65536(ctx):
    """2^16 messages (BASELINE configs[3]): dec(enc(m)) == m for every index with true openings obtained from the
    trapdoor, ciphertexts bit-exact vs the C oracle on a 64-message sample."""
    from oracle import coracle as co
    n = 1 << 16
    tau = 0x0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF % bn.R
    ctx.srs_generate(L.fr_m(tau), 64, download=False)
    p = [rng.randrange(bn.R) for _ in range(64)]
    com_xy, com_inf = ctx.msm_g1(L.fr_vec(p).reshape(64, 8))
    ptau = kr.poly_eval(p, tau)
    dom = bn.Radix2Domain(n)
    pts = dom.elements()
    vals = [kr.poly_eval(p, z) for z in pts]
    k = [(ptau - vals[i]) * pow(tau - pts[i], -1, bn.R) % bn.R for i in range(n)]
    proofs, pinf = ctx.g1_mul_gen_batch(L.fr_vec(k).reshape(n, 8))
    nprng = np.random.default_rng(11)
    rs = nprng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); rs[:, 7] &= 0x0FFFFFFF
    msgs = nprng.integers(0, 256, size=n * 32, dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * 32
    P, V = L.fr_vec(pts).reshape(n, 8), L.fr_vec(vals).reshape(n, 8)
    ct, ci, mc = ctx.encrypt_batch(com_xy, com_inf, P, V, rs, msgs, off)
    out = ctx.decrypt_batch(proofs, pinf, ct, ci, mc, off)
    assert np.array_equal(out[: n * 32], msgs)
    m = 64
    tau2 = L.g2_m(bn.g2_mul(bn.G2_GEN, tau))
    ct_o, ci_o, mc_o = co.encrypt_batch(com_xy, com_inf, tau2, P[:m].copy(), V[:m].copy(), rs[:m].copy(), msgs[: 32 * m].copy(), off[: m + 1].copy(), threads=4)
    assert np.array_equal(ct_o, ct[:m]) and np.array_equal(ci_o, ci[:m]) and np.array_equal(mc_o[: 32 * m], mc[: 32 * m])
    # the commitment has now served 2^16 messages: its GT table is upgraded to 16-bit windows (we.cu); same bytes
    ct2, ci2, mc2 = ctx.encrypt_batch(com_xy, com_inf, P[:4096].copy(), V[:4096].copy(), rs[:4096].copy(), msgs[: 32 * 4096].copy(), off[:4097].copy())
    assert np.array_equal(ct2, ct[:4096]) and np.array_equal(ci2, ci[:4096]) and np.array_equal(mc2[: 32 * 4096], mc[: 32 * 4096])


# ------------------------------------------------------------------ BASELINE config 5: laconic OT at scale
def test_config5_laconic_ot_full_flow_at_scale(ctx):
    """tests/laconic_ot.rs flow with 2^k - 1 receiver bits on one GPU (k = 20, the BASELINE size; KB_OT_LOG_N overrides): vec_commit with open_fk at d = 2^k, 2 x (2^k - 1) encryptions, 2^k - 1
    decryptions; all indices round-trip, the other message set does not decrypt, and commitment / proofs /
    ciphertexts agree with the trapdoor and the C oracle on samples."""
    from oracle import coracle as co
    from tests import ot_flow
    k = int(os.environ.get("KB_OT_LOG_N", "20"))
    res = ot_flow.run(ctx, k, checker=(bn, co, L))
    print(json.dumps(res))
    assert res["roundtrip_all_indices"] and res["other_set_fails"]
    assert res["commit_vs_trapdoor"] and res["proofs_vs_trapdoor"] and res["ciphertexts_vs_c_oracle"]
