"""Encryption layer — mirror of src/enc.rs.  `Ciphertext<E> = (E::G2, Vec<u8>)` (src/enc.rs:13)."""
from __future__ import annotations

import numpy as np

from .kzg import KZGSetup, default_context
from .types import G1, G2, fr_array, rng_or_secure


def encrypt(rng, kzg_setup: KZGSetup, com: G1, point: int, value: int, msg: bytes):
    """src/enc.rs:19-40 -> (key_ct: G2, msg_ct: bytes); key XOR message is fused into the kernel.  rng must be
    cryptographically secure; rng=None uses the OS CSPRNG (types.SecureFrRng)."""
    r = rng_or_secure(rng).fr()
    n = len(msg)
    off = np.array([0, n], np.uint64)
    m = np.frombuffer(bytes(msg), np.uint8).copy() if n else np.zeros(1, np.uint8)
    ct, ct_inf, msg_ct = kzg_setup.ctx.encrypt_batch(com.xy, com.inf, fr_array([point]), fr_array([value]), fr_array([r]), m, off)
    return G2(ct[0], ct_inf[0]), bytes(msg_ct[:n])


def decrypt(proof: G1, ct, ctx=None) -> bytes:
    """src/enc.rs:44-55 (ctx=None: the context of the latest KZGSetup)"""
    ctx = ctx or default_context()
    key_ct, msg_ct = ct
    n = len(msg_ct)
    off = np.array([0, n], np.uint64)
    m = np.frombuffer(bytes(msg_ct), np.uint8).copy() if n else np.zeros(1, np.uint8)
    out = ctx.decrypt_batch(proof.xy.reshape(1, 16), np.array([proof.inf], np.uint8), key_ct.xy.reshape(1, 32),
                            np.array([key_ct.inf], np.uint8), m, off)
    return bytes(out[:n])


# ---- wire format.  The reference derives no serialisation for `Ciphertext<E>`; as a tuple it would be
# ark-serialize's `(G2, Vec<u8>)`: the point (compressed 64 B / uncompressed 128 B), then the byte vector as a u64
# little-endian length followed by the bytes.  The point bytes are produced / parsed on the GPU (kb_g2_serialize).
def ciphertexts_to_bytes(cts, ctx, compress: bool = True) -> list:
    """[(G2, bytes)] -> [bytes]"""
    if not cts:
        return []
    xy = np.stack([c[0].xy for c in cts]).astype(np.uint32)
    inf = np.array([1 if c[0].inf else 0 for c in cts], np.uint8)
    pts = ctx.g2_serialize(np.ascontiguousarray(xy), inf, compress)
    return [bytes(pts[i]) + len(c[1]).to_bytes(8, "little") + bytes(c[1]) for i, c in enumerate(cts)]


def ciphertexts_from_bytes(blobs, ctx, compress: bool = True, validate: bool = True) -> list:
    """[bytes] -> [(G2, bytes)]; raises ValueError where arkworks returns SerializationError."""
    if not blobs:
        return []
    plen = 64 if compress else 128
    for b in blobs:
        if len(b) < plen + 8 or int.from_bytes(b[plen:plen + 8], "little") != len(b) - plen - 8:
            raise ValueError("InvalidData: ciphertext length")
    pts = np.frombuffer(b"".join(bytes(b[:plen]) for b in blobs), np.uint8).reshape(len(blobs), plen)
    xy, inf, ok = ctx.g2_deserialize(pts, compress, validate)
    if not ok.all():
        raise ValueError("InvalidData: ciphertext %d does not decode to a valid G2 point" % int(np.argmin(ok)))
    return [(G2(xy[i], bool(inf[i])), bytes(b[plen + 8:])) for i, b in enumerate(blobs)]
