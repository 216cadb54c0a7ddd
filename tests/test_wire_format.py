"""Wire format: ark-serialize point encodings.  CPU part: the oracle's encoder / decoder are
mutually consistent and reject what arkworks rejects.  GPU part (marked): kb_g{1,2}_{de,}serialize against the
oracle byte for byte, including infinity, negated points, invalid encodings and points outside the G2 subgroup."""
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from tests import limbs as L

rng = random.Random(0x5EA1)


def _twist_point_outside_g2():
    for i in range(1, 200):
        x = (i, 1)
        y = bn.f2_sqrt(bn.f2_add(bn.f2_mul(bn.f2_sqr(x), x), bn.B2))
        if y is not None and not bn.g2_in_subgroup((x, y)):
            return (x, y)
    raise AssertionError


def test_oracle_roundtrip_and_flags():
    for k in [1, 2, 3, bn.R - 1] + [rng.randrange(bn.R) for _ in range(6)]:
        p, q = bn.g1_mul(bn.G1_GEN, k), bn.g2_mul(bn.G2_GEN, k)
        for c in (True, False):
            for pt in (p, bn.g1_neg(p)):
                b = bn.g1_serialize(pt, c)
                assert len(b) == (32 if c else 64) and bn.g1_deserialize(b, c) == pt
            for pt in (q, bn.g2_neg(q)):
                b = bn.g2_serialize(pt, c)
                assert len(b) == (64 if c else 128) and bn.g2_deserialize(b, c) == pt
        # exactly one of P, -P carries the "y is the larger root" flag
        assert (bn.g1_serialize(p, True)[-1] ^ bn.g1_serialize(bn.g1_neg(p), True)[-1]) == bn.SW_FLAG_Y_NEGATIVE
        assert (bn.g2_serialize(q, True)[-1] ^ bn.g2_serialize(bn.g2_neg(q), True)[-1]) == bn.SW_FLAG_Y_NEGATIVE
    for c in (True, False):
        assert bn.g1_serialize(None, c)[-1] == bn.SW_FLAG_INFINITY and bn.g1_deserialize(bn.g1_serialize(None, c), c) is None
        assert bn.g2_serialize(None, c)[-1] == bn.SW_FLAG_INFINITY and bn.g2_deserialize(bn.g2_serialize(None, c), c) is None
    # generator bytes: G1 = (1, 2) -> x = 1, y = 2 is the smaller root
    assert bn.g1_serialize(bn.G1_GEN, True) == (1).to_bytes(32, "little")
    assert bn.g1_serialize(bn.G1_GEN, False) == (1).to_bytes(32, "little") + (2).to_bytes(32, "little")


def test_oracle_rejects_invalid():
    with pytest.raises(ValueError):
        bn.g1_deserialize(bn.Q.to_bytes(32, "little"), True)                      # x = q is not canonical
    with pytest.raises(ValueError):
        bn.g1_deserialize(bytes(31) + bytes([0xC0]), True)                        # both flags
    bad = next(x for x in range(2, 50) if bn.fq_sqrt((x ** 3 + 3) % bn.Q) is None)
    with pytest.raises(ValueError):
        bn.g1_deserialize(bad.to_bytes(32, "little"), True)                       # no y for this x
    with pytest.raises(ValueError):
        bn.g1_deserialize((1).to_bytes(32, "little") + (3).to_bytes(32, "little"), False)   # (1, 3) is off the curve
    assert bn.g1_deserialize((1).to_bytes(32, "little") + (3).to_bytes(32, "little"), False, validate=False) == (1, 3)
    t = _twist_point_outside_g2()
    for c in (True, False):
        with pytest.raises(ValueError):
            bn.g2_deserialize(bn.g2_serialize(t, c), c)                           # on the twist, outside the r-torsion
        assert bn.g2_deserialize(bn.g2_serialize(t, c), c, validate=False) == t


def test_f2_sqrt():
    for _ in range(20):
        a = (rng.randrange(bn.Q), rng.randrange(bn.Q))
        s = bn.f2_sqrt(bn.f2_sqr(a))
        assert s in (a, bn.f2_neg(a))
    for a in [(5, 0), (bn.Q - 5, 0), (0, 7), (0, 0)]:
        s = bn.f2_sqrt(bn.f2_sqr(a))
        assert s is not None and bn.f2_sqr(s) == bn.f2_sqr(a)


@pytest.fixture(scope="module")
def ctx():
    from keaki_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_wire_format_matches_oracle(ctx):
    ks = [1, 2, 3, bn.R - 1, bn.R - 2] + [rng.randrange(bn.R) for _ in range(40)]
    g1 = [bn.g1_mul(bn.G1_GEN, k) for k in ks] + [None]
    g2 = [bn.g2_mul(bn.G2_GEN, k) for k in ks] + [None]
    g1 += [bn.g1_neg(p) for p in g1[:8]]
    g2 += [bn.g2_neg(p) for p in g2[:8]]
    n = len(g1)
    xy1 = np.stack([L.g1_m(p) if p else np.zeros(16, np.uint32) for p in g1]); inf1 = np.array([p is None for p in g1], np.uint8)
    xy2 = np.stack([L.g2_m(p) if p else np.zeros(32, np.uint32) for p in g2]); inf2 = np.array([p is None for p in g2], np.uint8)
    for c in (True, False):
        b1, b2 = ctx.g1_serialize(xy1, inf1, c), ctx.g2_serialize(xy2, inf2, c)
        for i in range(n):
            assert bytes(b1[i]) == bn.g1_serialize(g1[i], c), (i, c)
            assert bytes(b2[i]) == bn.g2_serialize(g2[i], c), (i, c)
        for validate in (True, False):
            x1, i1, ok1 = ctx.g1_deserialize(b1, c, validate)
            x2, i2, ok2 = ctx.g2_deserialize(b2, c, validate)
            assert ok1.all() and ok2.all()
            assert np.array_equal(i1, inf1) and np.array_equal(i2, inf2)
            assert np.array_equal(x1, xy1) and np.array_equal(x2, xy2)


@pytest.mark.gpu
def test_gpu_wire_format_rejects_what_the_oracle_rejects(ctx):
    bad_x = next(x for x in range(2, 50) if bn.fq_sqrt((x ** 3 + 3) % bn.Q) is None)
    blobs = [bn.Q.to_bytes(32, "little"), bytes(31) + bytes([0xC0]), bad_x.to_bytes(32, "little"), (1).to_bytes(32, "little")]
    data = np.frombuffer(b"".join(blobs), np.uint8).reshape(4, 32)
    _, inf, ok = ctx.g1_deserialize(data, True, True)
    want = []
    for b in blobs:
        try:
            bn.g1_deserialize(b, True); want.append(1)
        except ValueError:
            want.append(0)
    assert list(ok) == want == [0, 0, 0, 1] and list(inf) == [1, 1, 1, 0]
    off = (1).to_bytes(32, "little") + (3).to_bytes(32, "little")
    _, _, ok = ctx.g1_deserialize(np.frombuffer(off, np.uint8).reshape(1, 64), False, True)
    assert ok[0] == 0
    xy, _, ok = ctx.g1_deserialize(np.frombuffer(off, np.uint8).reshape(1, 64), False, False)
    assert ok[0] == 1 and L.g1_from(xy[0]) == (1, 3)
    # uncompressed with a non-canonical y (y = q), and with stray flag bits in the x bytes: InvalidData in arkworks
    bad_y = (1).to_bytes(32, "little") + bn.Q.to_bytes(32, "little")
    stray = bytearray(bn.g1_serialize(bn.G1_GEN, False)); stray[31] |= 0x80
    for blob in (bad_y, bytes(stray)):
        with pytest.raises(ValueError):
            bn.g1_deserialize(blob, False)
        _, inf, ok = ctx.g1_deserialize(np.frombuffer(blob, np.uint8).reshape(1, 64), False, True)
        assert ok[0] == 0 and inf[0] == 1
    t = _twist_point_outside_g2()
    for c in (True, False):
        blob = np.frombuffer(bn.g2_serialize(t, c), np.uint8).reshape(1, -1)
        _, _, ok = ctx.g2_deserialize(blob, c, True)
        assert ok[0] == 0
        xy, _, ok = ctx.g2_deserialize(blob, c, False)
        assert ok[0] == 1 and L.g2_from(xy[0]) == t


@pytest.mark.gpu
def test_gpu_ciphertexts_travel_as_bytes(ctx):
    """sender -> bytes -> receiver: encrypt, serialise the ciphertexts, parse them back, decrypt (src/enc.rs:70-125 flow)."""
    from keaki_b200 import kzg, enc, FrRng
    r = FrRng(21)
    s = kzg.KZGSetup.setup(r.fr(), 8, ctx=ctx)
    p = [r.fr() for _ in range(8)]
    com = kzg.commit(s, p)
    from oracle import keaki_ref as kr
    pts = [r.fr() for _ in range(3)]
    cts, msgs, proofs = [], [], []
    for z in pts:
        m = r.bytes(32)
        cts.append(enc.encrypt(r, s, com, z, kr.poly_eval(p, z), m)); msgs.append(m); proofs.append(kzg.open(s, p, z))
    for compress in (True, False):
        blobs = enc.ciphertexts_to_bytes(cts, ctx, compress)
        assert all(len(b) == (64 if compress else 128) + 8 + 32 for b in blobs)
        back = enc.ciphertexts_from_bytes(blobs, ctx, compress)
        assert [c[0] for c in back] == [c[0] for c in cts] and [c[1] for c in back] == [c[1] for c in cts]
        assert [enc.decrypt(proofs[i], back[i], ctx=ctx) for i in range(3)] == msgs
    with pytest.raises(ValueError):
        enc.ciphertexts_from_bytes([bytes(63) + bytes([0xC0]) + (0).to_bytes(8, "little")], ctx, True)
