"""Extractable witness KEM — mirror of src/kem.rs."""
from __future__ import annotations

import numpy as np

from .kzg import KZGSetup, default_context
from .types import G1, G2, fr_array, rng_or_secure


def encapsulate(rng, kzg_setup: KZGSetup, commitment: G1, point: int, value: int, msg_len: int):
    """src/kem.rs:13-50 -> (ciphertext: G2, key: bytes).  Draws one `Fr::rand` from rng (:26).  rng must be
    cryptographically secure; rng=None uses the OS CSPRNG (types.SecureFrRng)."""
    r = rng_or_secure(rng).fr()
    off = np.array([0, msg_len], np.uint64)
    ct, ct_inf, key = kzg_setup.ctx.encrypt_batch(commitment.xy, commitment.inf, fr_array([point]), fr_array([value]),
                                                  fr_array([r]), np.zeros(max(msg_len, 1), np.uint8), off)
    return G2(ct[0], ct_inf[0]), bytes(key[:msg_len])


def decapsulate(proof: G1, ciphertext: G2, msg_len: int, ctx=None) -> bytes:
    """src/kem.rs:55-72.  The reference needs no setup here: ctx=None uses the context of the latest KZGSetup."""
    ctx = ctx or default_context()
    off = np.array([0, msg_len], np.uint64)
    out = ctx.decrypt_batch(proof.xy.reshape(1, 16), np.array([proof.inf], np.uint8), ciphertext.xy.reshape(1, 32),
                            np.array([ciphertext.inf], np.uint8), np.zeros(max(msg_len, 1), np.uint8), off)
    return bytes(out[:msg_len])
