#!/usr/bin/env python3
"""Generates the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the GPU box only reads the generated files).

  ppot_0080_01_mini.ptau   the sections of the reference's own fixture ptau/ppot_0080_01.ptau.test that
                           keaki reads (header, TauG1: 3 points, TauG2: 2 points), byte-for-byte, in a
                           valid 11-section container (other sections empty) — lets
                           KZGSetup.new_from_file run on the GPU box, where /root/reference is absent.
  ppot_0080_01_powers.json the same points de-Montgomerised (canonical integers) + container constants
                           the reference's tests assert (src/kzg/ptau.rs:384-474).
  oracle_vectors.json      known-answer vectors produced by the big-int oracle for fixed seeds
                           (commitment, FK proofs, ciphertexts, keys, GT bytes).  The reference holds
                           no byte-level vectors, so these pin the oracle against
                           regressions and give the GPU path committed bytes to match.
"""
import json
import os
import random
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import bn254 as bn  # noqa: E402
from oracle import keaki_ref as kr  # noqa: E402

REF_PTAU = "/root/reference/ptau/ppot_0080_01.ptau.test"
SECTION_IDS = (1, 2, 3, 4, 5, 6, 7, 12, 13, 14, 15)


def main():
    data = open(REF_PTAU, "rb").read()
    sec = kr.parse_ptau_sections(data)
    keep = {sid: data[sec[sid][0]: sec[sid][0] + sec[sid][1]] for sid in (1, 2, 3)}
    with open(os.path.join(HERE, "ppot_0080_01_mini.ptau"), "wb") as f:
        f.write(b"ptau" + struct.pack("<II", 1, 11))
        for sid in SECTION_IDS:
            body = keep.get(sid, b"")
            f.write(struct.pack("<IQ", sid, len(body)) + body)
    g1, g2 = kr.get_powers_from_file(REF_PTAU)
    hoff = sec[1][0]
    n8 = struct.unpack_from("<I", data, hoff)[0]
    power, ceremony = struct.unpack_from("<II", data, hoff + 4 + n8)
    json.dump({"file_len": len(data), "n_sections": 11, "power": power, "ceremony_power": ceremony,
               "modulus": str(int.from_bytes(data[hoff + 4: hoff + 4 + n8], "little")),
               "section_offsets": {str(k): list(v) for k, v in sec.items()},
               "g1": [[str(p[0]), str(p[1])] for p in g1],
               "g2": [[[str(p[0][0]), str(p[0][1])], [str(p[1][0]), str(p[1][1])]] for p in g2]},
              open(os.path.join(HERE, "ppot_0080_01_powers.json"), "w"), indent=1)

    rng = random.Random(0xB200_0001)
    tau = rng.randrange(1, bn.R)
    d = 8
    setup = kr.KZGSetup.setup(tau, d)
    p = [rng.randrange(bn.R) for _ in range(d)]
    com = kr.commit(setup, p)
    dom = bn.Radix2Domain(d)
    proofs = kr.open_fk(setup, p, dom)
    points = dom.elements()
    values = [kr.poly_eval(p, z) for z in points]
    rs = [rng.randrange(bn.R) for _ in range(d)]
    msgs = [bytes(rng.randrange(256) for _ in range(ln)) for ln in (32, 32, 0, 1, 64, 65, 100, 32)]
    cts = kr.vec_encrypt(rs, setup, com, points, values, msgs)
    gt = bn.pairing(bn.g1_mul(bn.G1_GEN, 5), bn.g2_mul(bn.G2_GEN, 7))
    vec = {
        "tau": str(tau), "coeffs": [str(c) for c in p],
        "commitment": [str(com[0]), str(com[1])],
        "proofs": [[str(q[0]), str(q[1])] if q else None for q in proofs],
        "points": [str(z) for z in points], "values": [str(v) for v in values], "r": [str(r) for r in rs],
        "messages": [m.hex() for m in msgs],
        "ciphertexts": [{"g2": [[str(c[0][0][0]), str(c[0][0][1])], [str(c[0][1][0]), str(c[0][1][1])]], "msg_ct": c[1].hex()} for c in cts],
        "keys32": [kr.gt_key(bn.pairing(proofs[i], cts[i][0]), 32).hex() for i in range(d)],
        "gt_5_7": bn.gt_to_bytes(gt).hex(),
        "gt_one_key32": kr.gt_key(bn.F12_ONE, 32).hex(),
        "tau_g2": [[str(setup.tau_g2[0][0]), str(setup.tau_g2[0][1])], [str(setup.tau_g2[1][0]), str(setup.tau_g2[1][1])]],
        # wire format: ark-serialize bytes of the proofs (G1) and of the ciphertext points (G2)
        "wire": {"proofs_compressed": [bn.g1_serialize(q, True).hex() for q in proofs],
                 "proofs_uncompressed": [bn.g1_serialize(q, False).hex() for q in proofs],
                 "ct_compressed": [bn.g2_serialize(c[0], True).hex() for c in cts],
                 "ct_uncompressed": [bn.g2_serialize(c[0], False).hex() for c in cts]},
    }
    vec["wire"]["commitment_uncompressed"] = bn.g1_serialize(com, False).hex()
    json.dump(vec, open(os.path.join(HERE, "oracle_vectors.json"), "w"), indent=1)
    # the same vectors as a line-based fixture for the C++ host-layer test (tests/cpp): `key hex [hex ...]`, scalars as
    # 32-byte little-endian canonical integers, points as their ark-serialize uncompressed bytes
    le = lambda x: int(x).to_bytes(32, "little").hex()
    with open(os.path.join(HERE, "oracle_vectors.txt"), "w") as f:
        f.write("tau " + le(tau) + "\n")
        f.write("coeffs " + " ".join(le(c) for c in p) + "\n")
        f.write("points " + " ".join(le(z) for z in points) + "\n")
        f.write("values " + " ".join(le(v) for v in values) + "\n")
        f.write("r " + " ".join(le(r) for r in rs) + "\n")
        f.write("messages " + " ".join(m.hex() or "-" for m in msgs) + "\n")
        f.write("commitment " + vec["wire"]["commitment_uncompressed"] + "\n")
        f.write("proofs " + " ".join(vec["wire"]["proofs_uncompressed"]) + "\n")
        f.write("ct " + " ".join(vec["wire"]["ct_uncompressed"]) + "\n")
        f.write("msg_ct " + " ".join(c[1].hex() or "-" for c in cts) + "\n")
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
