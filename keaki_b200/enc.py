"""Encryption layer — mirror of src/enc.rs.  `Ciphertext<E> = (E::G2, Vec<u8>)` (src/enc.rs:13)."""
from __future__ import annotations

import numpy as np

from .kzg import KZGSetup
from .types import G1, G2, fr_array


def encrypt(rng, kzg_setup: KZGSetup, com: G1, point: int, value: int, msg: bytes):
    """src/enc.rs:19-40 -> (key_ct: G2, msg_ct: bytes); key XOR message is fused into the kernel."""
    r = rng.fr()
    n = len(msg)
    off = np.array([0, n], np.uint64)
    m = np.frombuffer(bytes(msg), np.uint8).copy() if n else np.zeros(1, np.uint8)
    ct, ct_inf, msg_ct = kzg_setup.ctx.encrypt_batch(com.xy, com.inf, fr_array([point]), fr_array([value]), fr_array([r]), m, off)
    return G2(ct[0], ct_inf[0]), bytes(msg_ct[:n])


def decrypt(proof: G1, ct, ctx=None) -> bytes:
    """src/enc.rs:44-55"""
    key_ct, msg_ct = ct
    n = len(msg_ct)
    off = np.array([0, n], np.uint64)
    m = np.frombuffer(bytes(msg_ct), np.uint8).copy() if n else np.zeros(1, np.uint8)
    out = ctx.decrypt_batch(proof.xy.reshape(1, 16), np.array([proof.inf], np.uint8), key_ct.xy.reshape(1, 32),
                            np.array([key_ct.inf], np.uint8), m, off)
    return bytes(out[:n])
