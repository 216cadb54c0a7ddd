"""Vector layer — mirror of src/vec.rs.  The reference's sequential `for` loops over
encrypt/decrypt (src/vec.rs:63-66,75-78) are the batch dimension of one GPU launch here."""
from __future__ import annotations

import numpy as np

from .kzg import KZGSetup, commit, default_context
from .types import G1, G2, Radix2EvaluationDomain, fr_array, fr_list, pack_g1, pack_g2, rng_or_secure, unpack_g1

PADDING_LEN = 1  # src/vec.rs:18


def vec_commit(rng, kzg_setup: KZGSetup, vec):
    """src/vec.rs:22-49 -> (commitment, proofs).  One `Fr::rand` draw for the padding (:32): the pad is what hides the
    vector, so rng must be cryptographically secure; rng=None uses the OS CSPRNG (types.SecureFrRng)."""
    d = len(vec) + PADDING_LEN
    padded = [int(v) for v in vec] + [rng_or_secure(rng).fr()]
    domain = Radix2EvaluationDomain(d)
    evals = fr_array(padded + [0] * (domain.size - len(padded)))
    p_coeff = kzg_setup.ctx.fr_ntt(evals, inverse=True)             # domain.ifft  (:37)
    proofs_xy, proofs_inf = kzg_setup.ctx.open_all_fk(p_coeff)      # open_fk      (:40)
    com = commit(kzg_setup, fr_list(p_coeff))                       # commit       (:46)
    return com, unpack_g1(proofs_xy, proofs_inf)


def vec_encrypt(rng, kzg_setup: KZGSetup, com: G1, points, values, messages):
    """src/vec.rs:52-69 -> list of (G2, bytes).  r_i are drawn in index order, one per message, exactly
    as the reference's loop consumes its rng (:63-66 -> src/kem.rs:26).  rng must be cryptographically secure; rng=None
    uses the OS CSPRNG (types.SecureFrRng)."""
    n = len(messages)
    rng = rng_or_secure(rng)
    rs = [rng.fr() for _ in range(n)]
    lens = [len(m) for m in messages]
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    flat = np.frombuffer(b"".join(bytes(m) for m in messages), np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
    # like the reference, indexing points[i] / values[i] beyond their length is an error (panic at :64)
    ct, ct_inf, msg_ct = kzg_setup.ctx.encrypt_batch(com.xy, com.inf, fr_array([points[i] for i in range(n)]),
                                                     fr_array([values[i] for i in range(n)]), fr_array(rs), flat, off)
    return [(G2(ct[i], ct_inf[i]), bytes(msg_ct[int(off[i]): int(off[i + 1])])) for i in range(n)]


def vec_decrypt(proofs, cts, ctx=None):
    """src/vec.rs:72-81 -> list of bytes (ctx=None: the context of the latest KZGSetup)"""
    ctx = ctx or default_context()
    n = len(cts)
    pxy, pinf = pack_g1([proofs[i] for i in range(n)])
    cxy, cinf = pack_g2([c[0] for c in cts])
    lens = [len(c[1]) for c in cts]
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    flat = np.frombuffer(b"".join(bytes(c[1]) for c in cts), np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
    out = ctx.decrypt_batch(pxy, pinf, cxy, cinf, flat, off, n=n)
    return [bytes(out[int(off[i]): int(off[i + 1])]) for i in range(n)]
