"""Warp-cooperative pairing kernel (pairing_warp.cu) on the GPU: small batches through the C ABI vs the C oracle, and
against the one-thread-per-pairing kernel on the same inputs (the two are independent implementations of the same
function: compiled tower code vs generated lane-parallel schedule)."""
import os
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from tests import limbs as L

pytestmark = pytest.mark.gpu

rng = random.Random(0x3A7)
nprng = np.random.default_rng(0x3A7)
TAU = rng.randrange(1, bn.R)
KAT_GT_ONE_BYTES = (1).to_bytes(32, "little") + bytes(352)


def rand_fr_limbs(n):
    out = np.zeros((n, 8), np.uint32)
    for i in range(n):
        out[i] = L.int_to_limbs(rng.randrange(bn.R))
    return out


@pytest.fixture(scope="module")
def ctx():
    from keaki_b200 import _ffi
    os.environ.pop("KB_PAIRING_WARP_MAX", None)
    os.environ.pop("KB_ENCRYPT_WARP_MAX", None)
    c = _ffi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_thread():
    """a context that never takes the warp-cooperative kernel"""
    from keaki_b200 import _ffi
    os.environ["KB_PAIRING_WARP_MAX"] = "0"
    os.environ["KB_ENCRYPT_WARP_MAX"] = "0"
    try:
        c = _ffi.Context(0)
    finally:
        os.environ.pop("KB_PAIRING_WARP_MAX", None)
        os.environ.pop("KB_ENCRYPT_WARP_MAX", None)
    yield c
    c.close()


def _points(ctx, n):
    g1, i1 = ctx.g1_mul_gen_batch(rand_fr_limbs(n))
    # G2 points: ciphertext points of an encryption (r * (tau2 - a G2)) - whatever they are, both kernels and the oracle see the same
    ctx.srs_generate(L.fr_m(TAU), 64, download=False)
    com = bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R))
    off = np.arange(n + 1, dtype=np.uint64) * 32
    ct, ct_inf, _ = ctx.encrypt_batch(L.g1_m(com), 0, rand_fr_limbs(n), rand_fr_limbs(n), rand_fr_limbs(n), np.zeros(32 * n, np.uint8), off)
    return g1, i1, ct, ct_inf


@pytest.mark.parametrize("n", [1, 5, 33, 600, 2048])
def test_small_batches_match_c_oracle_and_thread_kernel(ctx, ctx_thread, n):
    from oracle import coracle as co
    g1, i1, g2, i2 = _points(ctx, n)
    if n >= 5:
        i1 = i1.copy(); i2 = i2.copy()
        i1[1] = 1
        i2[3] = 1
    got = ctx.pairing_batch(g1, i1, g2, i2)
    want = co.pairing_batch(g1, i1, g2, i2, threads=co.max_threads())
    assert np.array_equal(got, want)
    if n >= 5:
        assert bytes(got[1]) == KAT_GT_ONE_BYTES and bytes(got[3]) == KAT_GT_ONE_BYTES
    assert np.array_equal(ctx_thread.pairing_batch(g1, i1, g2, i2), got)


def test_generator_pairing_python_oracle(ctx):
    out = ctx.pairing_batch(L.g1_m(bn.G1_GEN).reshape(1, 16), np.zeros(1, np.uint8), L.g2_m(bn.G2_GEN).reshape(1, 32), np.zeros(1, np.uint8))
    assert bytes(out[0]) == bn.gt_to_bytes(bn.pairing(bn.G1_GEN, bn.G2_GEN))


@pytest.mark.parametrize("n", [1, 100, 1500])
def test_small_decrypt_batches_both_kernels(ctx, ctx_thread, n):
    """decapsulate + XOR (mode 1: GT words -> BLAKE3 XOF -> XOR) with ragged message lengths"""
    from oracle import coracle as co
    g1, i1, g2, i2 = _points(ctx, n)
    lens = nprng.integers(0, 100, size=n).astype(np.uint64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    msg_ct = nprng.integers(0, 256, size=max(int(off[-1]), 1), dtype=np.uint8)
    got = ctx.decrypt_batch(g1, i1, g2, i2, msg_ct, off)
    want = co.decrypt_batch(g1, i1, g2, i2, msg_ct, off, threads=co.max_threads())
    assert np.array_equal(got[: int(off[-1])], want[: int(off[-1])])
    assert np.array_equal(ctx_thread.decrypt_batch(g1, i1, g2, i2, msg_ct, off)[: int(off[-1])], got[: int(off[-1])])


def test_fresh_commitment_tables_both_paths(ctx, ctx_thread):
    """kb_encrypt_batch on a new commitment: A = e(com, G2) and its 32 window bases come from the warp kernels in one
    context and from the thread kernels in the other; the ciphertexts must be identical (and decrypt)"""
    n = 64
    for c in (ctx, ctx_thread):
        c.srs_generate(L.fr_m(TAU), 64, download=False)
    com = bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R))
    pts, vals, rs = rand_fr_limbs(n), rand_fr_limbs(n), rand_fr_limbs(n)
    msgs = nprng.integers(0, 256, size=32 * n, dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * 32
    a = ctx.encrypt_batch(L.g1_m(com), 0, pts, vals, rs, msgs, off)
    b = ctx_thread.encrypt_batch(L.g1_m(com), 0, pts, vals, rs, msgs, off)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("n", [1, 33, 700])
def test_small_encrypt_batches_warp_per_message(ctx, ctx_thread, n):
    """kb_encrypt_batch below the small-batch threshold (one warp per message: GT product tree + G2 shuffle tree) against the
    thread-per-message kernels and the C oracle: values 0, 1 and general, zero scalars, ragged message lengths"""
    from oracle import coracle as co
    for c in (ctx, ctx_thread):
        c.srs_generate(L.fr_m(TAU), 64, download=False)
    tau_g2 = L.g2_m(bn.g2_mul(bn.G2_GEN, TAU))
    com = bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R))
    pts, vals, rs = rand_fr_limbs(n), rand_fr_limbs(n), rand_fr_limbs(n)
    one = L.fr_m(1)
    for i in range(n):
        if i % 3 == 0:
            vals[i] = 0
        elif i % 3 == 1:
            vals[i] = one
    if n > 4:
        rs[2] = 0          # r = 0: secret = 1, ct = infinity
        pts[4] = 0         # alpha = 0
    lens = nprng.integers(0, 90, size=n).astype(np.uint64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    msgs = nprng.integers(0, 256, size=max(int(off[-1]), 1), dtype=np.uint8)
    a = ctx.encrypt_batch(L.g1_m(com), 0, pts, vals, rs, msgs, off)
    b = ctx_thread.encrypt_batch(L.g1_m(com), 0, pts, vals, rs, msgs, off)
    w = co.encrypt_batch(L.g1_m(com), 0, tau_g2, pts, vals, rs, msgs, off, threads=co.max_threads())
    total = int(off[-1])
    for x, y, z in zip(a, b, w):
        x, y, z = (np.asarray(t).reshape(-1) for t in (x, y, z))
        k = total if x.dtype == np.uint8 and x.size >= total and x.size != n else x.size
        assert np.array_equal(x[:k], y[:k])
        assert np.array_equal(x[:k], z[:k])


def test_segmented_rounds_match_whole_pairings(ctx):
    """batches of more than one round of resident warps run the pairing as 12 segments packed into rounds (pairing_st.cu);
    the same batch with whole pairings per launch (KB_PAIRING_SEGMENTS=0) must give the same bytes - incl. infinity operands,
    a batch size that is not a multiple of 32, and a sample against the C oracle"""
    from keaki_b200 import _ffi
    from oracle import coracle as co
    n = 40000 + 13
    g1, i1, g2, i2 = _points(ctx, n)
    i1 = i1.copy(); i2 = i2.copy()
    i1[[0, 31, 32, 20000, n - 1]] = 1
    i2[[5, 39999]] = 1
    got = ctx.pairing_batch(g1, i1, g2, i2)
    os.environ["KB_PAIRING_SEGMENTS"] = "0"
    try:
        c0 = _ffi.Context(0)
    finally:
        os.environ.pop("KB_PAIRING_SEGMENTS", None)
    try:
        want = c0.pairing_batch(g1, i1, g2, i2)
    finally:
        c0.close()
    assert np.array_equal(got, want)
    idx = np.array(sorted({0, 1, 5, 31, 32, 33, 12345, 20000, 37887, 37888, 39999, n - 2, n - 1}))
    ref = co.pairing_batch(np.ascontiguousarray(g1[idx]), np.ascontiguousarray(i1[idx]), np.ascontiguousarray(g2[idx]), np.ascontiguousarray(i2[idx]),
                           threads=co.max_threads())
    assert np.array_equal(got[idx], ref)
    assert bytes(got[0]) == KAT_GT_ONE_BYTES and bytes(got[5]) == KAT_GT_ONE_BYTES
