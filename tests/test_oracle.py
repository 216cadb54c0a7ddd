"""Pins the oracle: internal consistency, the reference's ptau fixture, committed golden vectors."""
import json
import os

import blake3

from oracle import bn254 as bn
from oracle import keaki_ref as kr

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_curve_constants_and_generators():
    assert bn.g1_on_curve(bn.G1_GEN) and bn.g2_on_curve(bn.G2_GEN)
    assert bn.g1_mul(bn.G1_GEN, bn.R) is None and bn.g2_mul(bn.G2_GEN, bn.R - 1) == bn.g2_neg(bn.G2_GEN)
    assert bn.MONT_R % bn.Q == 0x0E0A77C19A07DF2F666EA36F7879462C0A78EB28F5C70B3DD35D438DC58F0D9D
    assert (-pow(bn.Q, -1, 1 << 64)) % (1 << 64) == 0x87D20782E4866389
    assert (-pow(bn.R, -1, 1 << 32)) % (1 << 32) == 0xEFFFFFFF
    # 2^28-th root of unity constant of ark-bn254 Fr
    assert pow(5, (bn.R - 1) >> 28, bn.R) == 19103219067921713944291392827692070036145651957329286315305642004821462161904


def test_final_exponent_identity_and_chain():
    # arkworks' hard part is 2z(6z^2+3z+1) times the textbook exponent
    assert bn.HARD_EXP == 2 * bn.Z * (6 * bn.Z**2 + 3 * bn.Z + 1) * ((bn.Q**4 - bn.Q**2 + 1) // bn.R)
    f = bn.miller_loop(bn.G1_GEN, bn.G2_GEN)
    e = bn.final_exponentiation(f)
    assert e == bn.final_exponentiation_naive(f)
    assert e != bn.F12_ONE and bn.f12_pow(e, bn.R) == bn.F12_ONE


def test_bilinearity_and_infinity():
    a, b = 0x1234567, 0x7654321
    e = bn.pairing(bn.G1_GEN, bn.G2_GEN)
    assert bn.pairing(bn.g1_mul(bn.G1_GEN, a), bn.g2_mul(bn.G2_GEN, b)) == bn.f12_pow(e, a * b)
    assert bn.pairing(None, bn.G2_GEN) == bn.F12_ONE and bn.pairing(bn.G1_GEN, None) == bn.F12_ONE


def test_frobenius_matches_power():
    f = bn.miller_loop(bn.g1_mul(bn.G1_GEN, 3), bn.G2_GEN)
    for k in (1, 2, 3):
        assert bn.f12_frobenius(f, k) == bn.f12_pow(f, bn.Q**k)


def test_blake3_kat_of_gt_one():
    # the key the reference derives whenever proof = inf or ct = inf
    ser = bn.gt_to_bytes(bn.F12_ONE)
    assert ser == b"\x01" + bytes(383)
    assert blake3.blake3(ser).digest(32).hex() == "207d2aaa3257b30b7c371b6804480c9b2a7a04b4f69847270c5aadf5e5bc9454"
    assert blake3.blake3(ser).digest(100)[:32] == blake3.blake3(ser).digest(32)  # XOF prefix property


def test_reference_ptau_fixture_decodes_to_generators():
    """The reference's own fixture (copied sections, tests/golden/make_golden.py): after de-Montgomerising,
    TauG1[0] = (1, 2), TauG2[0] = the G2 generator, every point is on its curve and the powers are
    consistent: e(g1[1], G2) == e(G1, tau_2)."""
    g1, g2 = kr.get_powers_from_file(os.path.join(GOLD, "ppot_0080_01_mini.ptau"))
    assert len(g1) == 3 and len(g2) == 2          # src/kzg/ptau.rs:476-514 assert exactly these lengths
    assert g1[0] == bn.G1_GEN and g2[0] == bn.G2_GEN
    assert bn.pairing(g1[1], bn.G2_GEN) == bn.pairing(bn.G1_GEN, g2[1])
    assert bn.pairing(g1[2], bn.G2_GEN) == bn.pairing(g1[1], g2[1])
    gold = json.load(open(os.path.join(GOLD, "ppot_0080_01_powers.json")))
    assert [[str(p[0]), str(p[1])] for p in g1] == gold["g1"]
    # container constants the reference's tests pin (src/kzg/ptau.rs:384-474)
    assert gold["file_len"] == 95634 and gold["power"] == 1 and gold["ceremony_power"] == 28 and int(gold["modulus"]) == bn.Q
    # raw (non de-Montgomerised) coordinates are NOT on the curve: the reference's canonical read is wrong
    raw = open(os.path.join(GOLD, "ppot_0080_01_mini.ptau"), "rb").read()
    off = kr.parse_ptau_sections(raw)[2][0]
    x, y = int.from_bytes(raw[off:off + 32], "little"), int.from_bytes(raw[off + 32:off + 64], "little")
    assert not bn.g1_on_curve((x, y))


def test_golden_vectors_reproduce():
    v = json.load(open(os.path.join(GOLD, "oracle_vectors.json")))
    tau = int(v["tau"]); p = [int(c) for c in v["coeffs"]]
    setup = kr.KZGSetup.setup(tau, len(p))
    com = kr.commit(setup, p)
    assert [str(com[0]), str(com[1])] == v["commitment"]
    proofs = kr.open_fk(setup, p, bn.Radix2Domain(len(p)))
    assert [[str(q[0]), str(q[1])] if q else None for q in proofs] == v["proofs"]
    msgs = [bytes.fromhex(m) for m in v["messages"]]
    cts = kr.vec_encrypt([int(r) for r in v["r"]], setup, com, [int(z) for z in v["points"]], [int(x) for x in v["values"]], msgs)
    for c, g in zip(cts, v["ciphertexts"]):
        assert c[1].hex() == g["msg_ct"] and str(c[0][0][0]) == g["g2"][0][0] and str(c[0][1][1]) == g["g2"][1][1]
    assert kr.vec_decrypt(proofs, cts) == msgs
    assert bn.gt_to_bytes(bn.pairing(bn.g1_mul(bn.G1_GEN, 5), bn.g2_mul(bn.G2_GEN, 7))).hex() == v["gt_5_7"]
    assert v["gt_one_key32"] == "207d2aaa3257b30b7c371b6804480c9b2a7a04b4f69847270c5aadf5e5bc9454"
    assert [bn.g1_serialize(q, True).hex() for q in proofs] == v["wire"]["proofs_compressed"]
    assert [bn.g2_serialize(c[0], False).hex() for c in cts] == v["wire"]["ct_uncompressed"]
    assert [bn.g2_deserialize(bytes.fromhex(h), True) for h in v["wire"]["ct_compressed"]] == [c[0] for c in cts]
