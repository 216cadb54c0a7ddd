"""The reference's own unit / integration tests, re-instantiated on BN254 against the oracle
(they are all algebraic: SURVEY.md §4).  Each test cites the reference test it restates."""
import random

import pytest

from oracle import bn254 as bn
from oracle import keaki_ref as kr

rng = random.Random(20240264)
SECRET = rng.randrange(1, bn.R)


def test_kzg_setup():                     # src/kzg.rs:218-239
    s = kr.KZGSetup.setup(SECRET, 4)
    assert len(s.g1_pow) == 4
    for i in range(4):
        assert s.g1_pow[i] == bn.g1_mul(bn.G1_GEN, pow(SECRET, i, bn.R))
    assert s.tau_g2 == bn.g2_mul(bn.G2_GEN, SECRET)


def test_kzg_commit():                    # src/kzg.rs:241-258
    s = kr.KZGSetup.setup(SECRET, 4)
    p = [1, 3, 2]
    expected = None
    for i, c in enumerate(p):
        expected = bn.g1_add(expected, bn.g1_mul(s.g1_pow[i], c))
    assert kr.commit(s, p) == expected


def test_kzg_commit_polynomial_too_large():   # src/kzg.rs:260-277
    s = kr.KZGSetup.setup(SECRET, 2)
    with pytest.raises(kr.PolynomialTooLarge) as e:
        kr.commit(s, [1, 3, 2, 4])
    assert e.value.args_ == (4, 2)


def test_kzg_open_polynomial_too_large():     # src/kzg.rs:279-308
    s = kr.KZGSetup.setup(SECRET, 2)
    with pytest.raises(kr.PolynomialTooLarge) as e:
        kr.open(s, [1, 2, 3, 4, 5, 6], 5)
    assert e.value.args_ == (5, 2)            # quotient.len() = 5


def test_kzg_open_and_verify():               # src/kzg.rs:310-331
    s = kr.KZGSetup.setup(SECRET, 4)
    p = [1, 3, 2]
    com = kr.commit(s, p)
    assert kr.poly_eval(p, 5) == 66
    proof = kr.open(s, p, 5)
    assert kr.verify(s, com, 5, 66, proof)


def test_kzg_verify_negative_cases():         # src/kzg.rs:333-468
    s = kr.KZGSetup.setup(SECRET, 10)
    p = [1, 2, 3, 4, 5, 6, 7]
    com = kr.commit(s, p)
    v = kr.poly_eval(p, 11)
    proof = kr.open(s, p, 11)
    assert kr.verify(s, com, 11, v, proof)
    assert not kr.verify(s, com, 12, v, proof)                     # wrong alpha
    assert not kr.verify(s, com, 11, (v + 1) % bn.R, proof)        # wrong beta
    assert not kr.verify(s, com, 11, v, kr.open(s, p, 6))          # wrong proof
    assert not kr.verify(s, kr.commit(s, [1, 2, 3]), 11, v, proof)  # wrong commitment


def test_open_fk_equals_open_at_roots():      # src/kzg.rs:470-505
    d = 8
    s = kr.KZGSetup.setup(SECRET, d)
    p = [rng.randrange(bn.R) for _ in range(d)]
    dom = bn.Radix2Domain(d)
    fk = kr.open_fk(s, p, dom)
    for i, w in enumerate(dom.elements()):
        assert fk[i] == kr.open(s, p, w)
    assert fk == kr.open_fk_direct(s, p)


def _kem_fixture():
    s = kr.KZGSetup.setup(SECRET, 10)
    p = [1, 2, 3, 4, 5, 6, 7]
    com = kr.commit(s, p)
    point = 11
    value = kr.poly_eval(p, point)
    return s, p, com, point, value


def test_kem_valid_opening_gives_same_key():  # src/kem.rs:87-116
    s, p, com, point, value = _kem_fixture()
    ct, k = kr.encapsulate(rng.randrange(bn.R), s, com, point, value, 32)
    assert kr.decapsulate(kr.open(s, p, point), ct, 32) == k and len(k) == 32


def test_kem_negative_cases():                # src/kem.rs:118-224
    s, p, com, point, value = _kem_fixture()
    r = rng.randrange(1, bn.R)
    ct, k = kr.encapsulate(r, s, com, point, value, 32)
    wrong_poly = [7, 6, 5, 4, 3, 2, 1]
    assert kr.decapsulate(kr.open(s, wrong_poly, point), ct, 32) != k          # wrong polynomial
    assert kr.decapsulate(kr.open(s, p, point), bn.g2_mul(ct, 2), 32) != k     # scaled ciphertext
    assert kr.decapsulate(kr.open(s, p, 12), ct, 32) != k                      # other point


def test_enc_roundtrip_and_wrong_proof():     # src/enc.rs:70-125
    s, p, com, point, value = _kem_fixture()
    msg = b"helloworld"
    ct = kr.encrypt(rng.randrange(bn.R), s, com, point, value, msg)
    assert kr.decrypt(kr.open(s, p, point), ct) == msg
    assert kr.decrypt(kr.open(s, p, 12), ct) != msg


def test_laconic_ot():                        # tests/laconic_ot.rs:127-200 (8 choices, 2 x 8 x 32 B values)
    n = 8
    s = kr.KZGSetup.setup(SECRET, 16)
    choices = [rng.randrange(2) for _ in range(n)]
    rcv = kr.Receiver(s, rng.randrange(bn.R), choices)
    snd = kr.Sender(s, rcv.commitment)
    values = [[bytes(rng.randrange(256) for _ in range(32)) for _ in range(n)] for _ in range(2)]
    enc = snd.send([rng.randrange(bn.R) for _ in range(n)], [rng.randrange(bn.R) for _ in range(n)], values)
    out = rcv.receive(enc)
    for i in range(n):
        assert out[i] == values[choices[i]][i]
