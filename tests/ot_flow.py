#!/usr/bin/env python3
"""Laconic OT at scale on ONE GPU through the C ABI — BASELINE.json configs[4] ("laconic OT with 2^20 receiver
bits", tests/laconic_ot.rs:15-113 of the reference): n - 1 choice bits + PADDING_LEN random element
(src/vec.rs:18,27-36) -> iFFT -> commit + open_fk (Receiver::new), 2 x (n - 1) encryptions (Sender::send),
n - 1 decryptions (Receiver::receive).  Every phase is timed; results are checked by

  * the trapdoor: commitment == p(tau) G1, proof_i == ((p(tau) - v_i) / (tau - w^i)) G1 on a sample,
  * the round trip: receive() returns exactly the chosen messages for ALL indices, and garbage for the other set,
  * the C oracle (oracle/c, arkworks-style algorithms) bit-for-bit on a sample of ciphertexts.

Used by tests/test_gpu_flows.py (the oracle is imported by the test, which passes it in) and as a script:
    python -m tests.ot_flow --log-n 20
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root
sys.path.insert(0, ROOT)


def run(ctx, log_n: int, seed: int = 5, sample: int = 16, checker=None) -> dict:
    """checker: optional (bn, coracle, limbs) modules from the TEST side for the oracle comparisons."""
    from keaki_b200.types import FR_MODULUS as R, fr_to_limbs, fr_from_limbs, Radix2EvaluationDomain

    n = 1 << log_n
    nprng = np.random.default_rng(seed)
    t = {}

    def tick(name, t0):
        t[name] = round(time.perf_counter() - t0, 4)

    tau = 0x2B3C4D5E6F708192A3B4C5D6E7F8091A2B3C4D5E6F708192A3B4C5D6E7F80911 % R
    t0 = time.perf_counter(); ctx.srs_generate(fr_to_limbs(tau), n, download=False); tick("srs_generate_s", t0)

    # Receiver::new — choices as field elements 0 / 1 (Montgomery limbs), one random padding element
    bits = nprng.integers(0, 2, size=n - 1, dtype=np.uint8)
    one, zero = fr_to_limbs(1), fr_to_limbs(0)
    pad = nprng.integers(0, 1 << 32, size=8, dtype=np.uint64).astype(np.uint32); pad[7] &= 0x0FFFFFFF
    evals = np.where(bits[:, None] == 1, one[None, :], zero[None, :]).astype(np.uint32)
    evals = np.ascontiguousarray(np.concatenate([evals, pad[None, :]]))
    t0 = time.perf_counter(); coeffs = ctx.fr_ntt(evals.copy(), inverse=True); tick("ifft_s", t0)
    t0 = time.perf_counter(); com_xy, com_inf = ctx.msm_g1(coeffs); tick("commit_s", t0)
    t0 = time.perf_counter(); proofs, pinf = ctx.open_all_fk(coeffs); tick("open_fk_first_s", t0)
    t0 = time.perf_counter(); proofs2, pinf2 = ctx.open_all_fk(coeffs); tick("open_fk_cached_s", t0)
    assert np.array_equal(proofs, proofs2) and np.array_equal(pinf, pinf2)

    # Sender::send — two message sets, every position encrypted under value 0 and under value 1
    m = n - 1
    points = Radix2EvaluationDomain(n).elements_limbs()[:m]
    msgs = [nprng.integers(0, 256, size=m * 32, dtype=np.uint8) for _ in range(2)]
    rs = []
    for _ in range(2):
        r = nprng.integers(0, 1 << 32, size=(m, 8), dtype=np.uint64).astype(np.uint32); r[:, 7] &= 0x0FFFFFFF
        rs.append(r)
    off = np.arange(m + 1, dtype=np.uint64) * 32
    vals = [np.ascontiguousarray(np.tile(zero, (m, 1))), np.ascontiguousarray(np.tile(one, (m, 1)))]
    enc = []
    t0 = time.perf_counter()
    for v in range(2):
        enc.append(ctx.encrypt_batch(com_xy, com_inf, points, vals[v], rs[v], msgs[v], off))
    tick("send_2x_encrypt_s", t0)

    # Receiver::receive — the ciphertext of the chosen set at every position
    sel = bits.astype(bool)
    ct = np.where(sel[:, None], enc[1][0], enc[0][0]); ci = np.where(sel, enc[1][1], enc[0][1])
    mc = np.where(np.repeat(sel, 32), enc[1][2][: m * 32], enc[0][2][: m * 32])
    t0 = time.perf_counter()
    out = ctx.decrypt_batch(np.ascontiguousarray(proofs[:m]), np.ascontiguousarray(pinf[:m]), np.ascontiguousarray(ct),
                            np.ascontiguousarray(ci), np.ascontiguousarray(mc), off)
    tick("receive_decrypt_s", t0)
    want = np.where(np.repeat(sel, 32), msgs[1], msgs[0])
    ok_roundtrip = bool(np.array_equal(out[: m * 32], want))
    # the other set must NOT decrypt (wrong value for the opening): check a slice
    k = min(m, 4096)
    ct_w = np.where(sel[:k, None], enc[0][0][:k], enc[1][0][:k]); ci_w = np.where(sel[:k], enc[0][1][:k], enc[1][1][:k])
    mc_w = np.where(np.repeat(sel[:k], 32), enc[0][2][: k * 32], enc[1][2][: k * 32])
    out_w = ctx.decrypt_batch(np.ascontiguousarray(proofs[:k]), np.ascontiguousarray(pinf[:k]), np.ascontiguousarray(ct_w),
                              np.ascontiguousarray(ci_w), np.ascontiguousarray(mc_w), off[: k + 1].copy())
    other = np.where(np.repeat(sel[:k], 32), msgs[0][: k * 32], msgs[1][: k * 32])
    wrong_rows = (out_w[: k * 32].reshape(k, 32) != other.reshape(k, 32)).any(axis=1)
    ok_wrong_set_fails = bool(wrong_rows.all())

    res = {"log_n": log_n, "times": t, "roundtrip_all_indices": ok_roundtrip, "other_set_fails": ok_wrong_set_fails,
           "encrypt_per_s": 2 * m / t["send_2x_encrypt_s"], "decrypt_per_s": m / t["receive_decrypt_s"]}

    if checker is not None:
        bn, co, L = checker
        # trapdoor: p(tau) by Horner over the coefficients
        t0 = time.perf_counter()
        raw = coeffs.tobytes()       # Montgomery limbs c_i * 2^256: Horner on them gives p(tau) * 2^256
        ptau = 0
        for i in range(n - 1, -1, -1):
            ptau = (ptau * tau + int.from_bytes(raw[32 * i: 32 * i + 32], "little")) % R
        ptau = ptau * pow(1 << 256, -1, R) % R
        res["commit_vs_trapdoor"] = bool((None if com_inf else L.g1_from(com_xy)) == bn.g1_mul(bn.G1_GEN, ptau))
        dom = bn.Radix2Domain(n)
        idx = [0, 1, n // 2, n - 2, n - 1] + [int(x) for x in nprng.integers(0, n, size=sample)]
        okp = True
        for i in idx:
            w = pow(dom.group_gen, i, R)
            v = fr_from_limbs(evals[i])
            want_p = bn.g1_mul(bn.G1_GEN, (ptau - v) * pow(tau - w, -1, R) % R)
            okp &= (None if pinf[i] else L.g1_from(proofs[i])) == want_p
        res["proofs_vs_trapdoor"] = bool(okp)
        # ciphertexts and masked messages bit-exact vs the C oracle on a sample
        tau2 = L.g2_m(bn.g2_mul(bn.G2_GEN, tau))
        s = min(sample, m)
        oke = True
        for v in range(2):
            ct_o, ci_o, mc_o = co.encrypt_batch(com_xy, com_inf, tau2, points[:s].copy(), vals[v][:s].copy(), rs[v][:s].copy(),
                                                msgs[v][: 32 * s].copy(), off[: s + 1].copy(), threads=4)
            oke &= np.array_equal(ct_o, enc[v][0][:s]) and np.array_equal(ci_o, enc[v][1][:s]) and np.array_equal(mc_o[: 32 * s], enc[v][2][: 32 * s])
        res["ciphertexts_vs_c_oracle"] = bool(oke)
        tick("oracle_checks_s", t0)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    from keaki_b200 import _ffi
    ctx = _ffi.Context(0)
    checker = None
    if not a.no_check:   # script use only: the checker lives under oracle/ + tests/ (test infrastructure)
        from oracle import bn254 as bn
        from oracle import coracle as co
        from tests import limbs as L
        checker = (bn, co, L)
    print(json.dumps(run(ctx, a.log_n, checker=checker)), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
