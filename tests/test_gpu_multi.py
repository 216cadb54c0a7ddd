"""In-library multi-GPU context (kb_ctx_create_multi): one process, one context over several GPUs of the
box.  `commit` split by point range with the partial sums added on the first device, encrypt / decrypt batches split by
index - results must be bit-identical to the single-device context and to the trapdoor.  Skipped on a one-GPU box."""
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from tests import limbs as L

pytestmark = pytest.mark.gpu

rng = random.Random(0x6D756C7469)
nprng = np.random.default_rng(0x6D32)
TAU = rng.randrange(1, bn.R)


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.fixture(scope="module")
def ctxs():
    from keaki_b200 import _ffi
    nd = _device_count()
    if nd < 2:
        pytest.skip("needs at least 2 GPUs")
    multi = _ffi.Context(list(range(min(nd, 8))))
    single = _ffi.Context(0)
    yield multi, single
    multi.close(); single.close()


def rand_fr_limbs(n):
    a = nprng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= 0x0FFFFFFF
    return a


def scalars_of(limbs):
    raw = np.ascontiguousarray(limbs, np.uint32).tobytes()
    rinv = pow(1 << 256, -1, bn.R)
    return [int.from_bytes(raw[32 * i: 32 * i + 32], "little") * rinv % bn.R for i in range(limbs.shape[0])]


def trapdoor(scalars, first=0):
    acc = 0
    for s in reversed(scalars):
        acc = (acc * TAU + s) % bn.R
    return bn.g1_mul(bn.G1_GEN, acc * pow(TAU, first, bn.R) % bn.R)


def test_multi_commit_matches_trapdoor_and_single(ctxs):
    multi, single = ctxs
    assert multi.device_count() >= 2 and single.device_count() == 1
    n = 1 << 16
    multi.srs_generate(L.fr_m(TAU), n, download=False)
    single.srs_generate(L.fr_m(TAU), n, download=False)
    for m, first in ((n, 0), (n - 777, 0), (50000, 1234), (3000, 0), (17, 5)):   # the last two fall below the split threshold
        S = rand_fr_limbs(m)
        xy, inf = multi.msm_g1(S, first=first)
        xs, infs = single.msm_g1(S, first=first)
        assert inf == infs and np.array_equal(xy, xs)
        assert (None if inf else L.g1_from(xy)) == trapdoor(scalars_of(S), first)
    # uploaded SRS reaches every device too
    g1, tau2 = single.srs_generate(L.fr_m(TAU), 1 << 13)
    multi.srs_upload(g1, tau2)
    S = rand_fr_limbs(1 << 13)
    xy, inf = multi.msm_g1(S)
    assert (None if inf else L.g1_from(xy)) == trapdoor(scalars_of(S))
    multi.srs_validate()


def test_multi_encrypt_decrypt_identical_to_single(ctxs):
    multi, single = ctxs
    n = 9000 + 13
    multi.srs_generate(L.fr_m(TAU), 256, download=False)
    single.srs_generate(L.fr_m(TAU), 256, download=False)
    com = bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R))
    pts, rs = rand_fr_limbs(n), rand_fr_limbs(n)
    bits = nprng.integers(0, 3, size=n)
    vals = rand_fr_limbs(n)
    vals[bits == 0] = L.fr_m(0)
    vals[bits == 1] = L.fr_m(1)
    lens = nprng.integers(0, 70, size=n)            # ragged messages, some empty
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    msgs = nprng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    ct_m, ci_m, mc_m = multi.encrypt_batch(L.g1_m(com), 0, pts, vals, rs, msgs, off)
    ct_s, ci_s, mc_s = single.encrypt_batch(L.g1_m(com), 0, pts, vals, rs, msgs, off)
    assert np.array_equal(ct_m, ct_s) and np.array_equal(ci_m, ci_s) and np.array_equal(mc_m, mc_s)
    proofs, pinf = single.g1_mul_gen_batch(rand_fr_limbs(n))
    pinf[[3, 4500, n - 1]] = 1
    out_m = multi.decrypt_batch(proofs, pinf, ct_m, ci_m, mc_m, off)
    out_s = single.decrypt_batch(proofs, pinf, ct_s, ci_s, mc_s, off)
    assert np.array_equal(out_m, out_s)


def test_multi_large_decrypt_segmented_per_device(ctxs):
    """a batch large enough that EVERY device runs the segmented pairing launches (more than one round of resident warps per
    device): sharded result identical to the single-device context, which runs the same batch as 2+ rounds itself"""
    multi, single = ctxs
    n = multi.device_count() * 38000 + 77
    for c in (multi, single):
        c.srs_generate(L.fr_m(TAU), 256, download=False)
    com = bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R))
    pts, rs = rand_fr_limbs(n), rand_fr_limbs(n)
    vals = np.zeros((n, 8), np.uint32)
    off = np.arange(n + 1, dtype=np.uint64) * 32
    msgs = nprng.integers(0, 256, size=32 * n, dtype=np.uint8)
    ct, ci, mc = single.encrypt_batch(L.g1_m(com), 0, pts, vals, rs, msgs, off)
    proofs, pinf = single.g1_mul_gen_batch(rand_fr_limbs(n))
    pinf[[0, 37999, 38000, n - 1]] = 1
    out_m = multi.decrypt_batch(proofs, pinf, ct, ci, mc, off)
    out_s = single.decrypt_batch(proofs, pinf, ct, ci, mc, off)
    assert np.array_equal(out_m, out_s)
