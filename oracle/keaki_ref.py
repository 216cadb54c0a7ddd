"""keaki API restated on the big-int oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Function-for-function CPU restatement of the reference's public API on BN254 (E = ark_bn254::Bn254),
each citing the reference file:line it follows.  G1/G2 values are affine tuples (None = infinity),
scalars are ints mod r.  Randomness is passed in explicitly (the reference draws `Fr::rand(rng)`:
src/kem.rs:26, src/vec.rs:32) so that the CUDA path and this oracle see the same r_i.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import builtins
import struct

import blake3 as _blake3

from . import bn254 as bn
from .bn254 import R, Radix2Domain

PADDING_LEN = 1  # src/vec.rs:18


class PolynomialTooLarge(Exception):
    """KZGError::PolynomialTooLarge(p.len(), g1_pow.len()) — src/kzg.rs:205-209."""

    def __init__(self, plen, maxlen):
        super().__init__(f"Can't commit to polynomial: polynomial has degree {plen} but max degree is {maxlen}")
        self.args_ = (plen, maxlen)


class KZGSetup:
    """src/kzg.rs:22-85."""

    def __init__(self, g1_aff, tau_g2):
        self.g1_aff = list(g1_aff)
        self.tau_g2 = tau_g2

    @property
    def g1_pow(self):
        return self.g1_aff

    @classmethod
    def setup(cls, secret: int, max_d: int) -> "KZGSetup":
        """src/kzg.rs:55-70: g1_pow[i] = G1 * secret^i, tau_g2 = G2 * secret."""
        pts, t = [], 1
        for _ in range(max_d):
            pts.append(bn.g1_mul(bn.G1_GEN, t))
            t = t * secret % R
        return cls(pts, bn.g2_mul(bn.G2_GEN, secret))

    @classmethod
    def new_from_file(cls, path: str) -> "KZGSetup":
        """src/kzg.rs:33-52 + src/kzg/ptau.rs:347-358.  DELIBERATE DEVIATION: the
        snarkjs file stores Montgomery limbs; the reference reads them as canonical and gets
        off-curve points.  Here they are de-Montgomerised and validated."""
        g1, g2 = get_powers_from_file(path)
        if len(g2) < 2:
            raise ValueError("EmptySection(3)")
        return cls(g1, g2[1])


def parse_ptau_sections(data: bytes):
    """src/kzg/ptau.rs:12-17,129-200: 'ptau' magic, u32 version, u32 n_sections, then per section
    u32 id + u64 length + payload.  Returns {id: (offset, length)}."""
    if data[:4] != b"ptau":
        raise ValueError("InvalidFileType")
    _version, n_sections = struct.unpack_from("<II", data, 4)
    off, sections = 12, {}
    for _ in range(n_sections):
        sid, slen = struct.unpack_from("<IQ", data, off)
        off += 12
        sections[sid] = (off, slen)
        off += slen
    return sections


def get_powers_from_file(path: str):
    """src/kzg/ptau.rs:347-358 (header :222-262, TauG1 :264-290, TauG2 :292-340)."""
    data = builtins.open(path, "rb").read()
    sec = parse_ptau_sections(data)
    hoff, _ = sec[1]
    n8 = struct.unpack_from("<I", data, hoff)[0]
    modulus = int.from_bytes(data[hoff + 4: hoff + 4 + n8], "little")
    power, _ceremony_power = struct.unpack_from("<II", data, hoff + 4 + n8)
    if modulus != bn.Q or n8 != 32:
        raise ValueError("field modulus is not BN254 q")

    def rd(off):
        return bn.from_mont(int.from_bytes(data[off: off + 32], "little"))

    n_g1 = 2 * (1 << power) - 1
    n_g2 = 1 << power
    o1, _ = sec[2]
    g1 = [(rd(o1 + 64 * i), rd(o1 + 64 * i + 32)) for i in range(n_g1)]
    o2, _ = sec[3]
    g2 = [((rd(o2 + 128 * i), rd(o2 + 128 * i + 32)), (rd(o2 + 128 * i + 64), rd(o2 + 128 * i + 96)))
          for i in range(n_g2)]
    for p in g1:
        if not bn.g1_on_curve(p):
            raise ValueError("TauG1 point off curve")
    for p in g2:
        if not bn.g2_on_curve(p):
            raise ValueError("TauG2 point off curve")
    return g1, g2


# ------------------------------------------------------------------------------------------
# KZG — src/kzg.rs
# ------------------------------------------------------------------------------------------
def commit(setup: KZGSetup, p):
    """src/kzg.rs:89-101: sum p_i * g1_aff[i] (msm_unchecked uses the min(len) prefix)."""
    if len(p) > len(setup.g1_aff):
        raise PolynomialTooLarge(len(p), len(setup.g1_aff))
    return bn.g1_msm(setup.g1_aff[: len(p)], p)


def poly_eval(p, z):
    acc = 0
    for c in reversed(p):
        acc = (acc * z + c) % R
    return acc


def quotient(p, z):
    """(p(x) - p(z)) / (x - z) by synthetic division — src/kzg.rs:104-120."""
    d = len(p)
    if d <= 1:
        return []
    q = [0] * (d - 1)
    q[d - 2] = p[d - 1] % R
    for i in range(d - 2, 0, -1):
        q[i - 1] = (p[i] + z * q[i]) % R
    return q


def open(setup: KZGSetup, p, point):  # noqa: A001 - mirrors the reference name
    """src/kzg.rs:104-124."""
    q = quotient(p, point)
    # DensePolynomial strips leading zeros before commit; the commit size check is on that length
    while q and q[-1] == 0:
        q.pop()
    return commit(setup, q)


def verify(setup: KZGSetup, commitment, point, value, proof) -> bool:
    """src/kzg.rs:127-151."""
    lhs = bn.pairing(bn.g1_add(commitment, bn.g1_neg(bn.g1_mul(bn.G1_GEN, value))), bn.G2_GEN)
    rhs = bn.pairing(proof, bn.g2_add(setup.tau_g2, bn.g2_neg(bn.g2_mul(bn.G2_GEN, point))))
    return lhs == rhs


def open_fk(setup: KZGSetup, p, domain_d: Radix2Domain):
    """src/kzg.rs:157-203, restating the reference's data flow (three G1 FFTs)."""
    d = len(p)
    dom2 = Radix2Domain(2 * d)
    s = list(reversed(setup.g1_pow[:d])) + [None] * d
    a = [0] * d + list(p)
    hat_s = dom2.fft_g1(s)
    hat_a = dom2.fft(a)
    hat_h = [bn.g1_mul(hat_s[i], hat_a[i]) for i in range(2 * d)]
    h = dom2.ifft_g1(hat_h)[:d]
    return domain_d.fft_g1(h)


def open_fk_direct(setup: KZGSetup, p):
    """Closed form of the same result: h_k = sum_m f_{m+k+1} [tau^m],
    pi_i = sum_k h_k w^{ik}.  Used to cross-check open_fk."""
    d = len(p)
    h = []
    for k in range(d):
        h.append(bn.g1_msm(setup.g1_pow[: max(d - 1 - k, 0)], p[k + 1:]))
    return Radix2Domain(d).fft_g1(h)


# ------------------------------------------------------------------------------------------
# KEM / encryption — src/kem.rs, src/enc.rs
# ------------------------------------------------------------------------------------------
def gt_key(gt, msg_len: int) -> bytes:
    """src/kem.rs:31-46: serialize_uncompressed (384 B) -> BLAKE3 -> XOF msg_len bytes."""
    return _blake3.blake3(bn.gt_to_bytes(gt)).digest(msg_len)


def encapsulate(r: int, setup: KZGSetup, commitment, point: int, value: int, msg_len: int):
    """src/kem.rs:13-50 with the random scalar r passed in (drawn at :26 in the reference)."""
    com_beta = bn.g1_add(commitment, bn.g1_neg(bn.g1_mul(bn.G1_GEN, value)))
    secret = bn.pairing(bn.g1_mul(com_beta, r), bn.G2_GEN)
    tau_alpha = bn.g2_add(setup.tau_g2, bn.g2_neg(bn.g2_mul(bn.G2_GEN, point)))
    ct = bn.g2_mul(tau_alpha, r)
    return ct, gt_key(secret, msg_len)


def decapsulate(proof, ciphertext, msg_len: int) -> bytes:
    """src/kem.rs:55-72."""
    return gt_key(bn.pairing(proof, ciphertext), msg_len)


def encrypt(r: int, setup: KZGSetup, com, point: int, value: int, msg: bytes):
    """src/enc.rs:19-40."""
    key_ct, key = encapsulate(r, setup, com, point, value, len(msg))
    return key_ct, bytes(k ^ m for k, m in zip(key, msg))


def decrypt(proof, ct) -> bytes:
    """src/enc.rs:44-55."""
    key = decapsulate(proof, ct[0], len(ct[1]))
    return bytes(k ^ c for k, c in zip(key, ct[1]))


# ------------------------------------------------------------------------------------------
# Vector layer — src/vec.rs
# ------------------------------------------------------------------------------------------
def vec_commit(pad_r: int, setup: KZGSetup, vec):
    """src/vec.rs:22-49 with the padding scalar passed in (drawn at :32)."""
    padded = list(vec) + [pad_r % R]
    domain = Radix2Domain(len(padded))
    p_coeff = domain.ifft(padded)
    proofs = open_fk(setup, p_coeff, domain)
    # DensePolynomial::from_coefficients_vec strips trailing zeros (src/vec.rs:43)
    dense = list(p_coeff)
    while dense and dense[-1] == 0:
        dense.pop()
    return commit(setup, dense), proofs


def vec_encrypt(rs, setup: KZGSetup, com, points, values, messages):
    """src/vec.rs:52-69: sequential map; rs[i] is the i-th Fr::rand draw."""
    return [encrypt(rs[i], setup, com, points[i], values[i], messages[i]) for i in range(len(messages))]


def vec_decrypt(proofs, cts):
    """src/vec.rs:72-81."""
    return [decrypt(proofs[i], cts[i]) for i in range(len(cts))]


# ------------------------------------------------------------------------------------------
# Laconic OT — tests/laconic_ot.rs:15-113
# ------------------------------------------------------------------------------------------
class Receiver:
    def __init__(self, setup: KZGSetup, pad_r: int, choices):
        self.choices = list(choices)
        self.commitment, self.proofs = vec_commit(pad_r, setup, self.choices)

    def receive(self, encrypted_sets):
        n = len(encrypted_sets[0])
        chosen = [encrypted_sets[0][i] if self.choices[i] == 0 else encrypted_sets[1][i] for i in range(n)]
        return vec_decrypt(self.proofs, chosen)


class Sender:
    def __init__(self, setup: KZGSetup, commitment):
        self.setup, self.commitment = setup, commitment

    def send(self, rs0, rs1, private_set):
        n_values = len(private_set[0])
        elements = Radix2Domain(n_values + PADDING_LEN).elements()
        ct0 = vec_encrypt(rs0, self.setup, self.commitment, elements, [0] * n_values, private_set[0])
        ct1 = vec_encrypt(rs1, self.setup, self.commitment, elements, [1] * n_values, private_set[1])
        return [ct0, ct1]
