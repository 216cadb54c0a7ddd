#!/usr/bin/env python3
"""Latency / throughput of the two pairing kernels against batch size: warp-cooperative (pairing_warp.cu) vs one thread per
pairing (pairing_st.cu), and the cold path of kb_encrypt_batch (fresh commitment).  Device milliseconds of the call
(kb_last_kernel_ms).  Usage on the GPU box:  python tools/exp/warp_sweep.py > gpurun_out/warp_sweep.txt"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from keaki_b200 import _ffi  # noqa: E402

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
rng = np.random.default_rng(1)


def rand_fr(n):
    out = np.zeros((n, 8), np.uint32)
    for i in range(n):
        x = int.from_bytes(rng.bytes(40), "little") % R
        for k in range(8):
            out[i, k] = (x >> (32 * k)) & 0xFFFFFFFF
    return out


def main():
    sizes = [1, 32, 148, 592, 1024, 2048, 4096, 8192]
    os.environ["KB_PAIRING_WARP_MAX"] = str(1 << 20)
    cw = _ffi.Context(0)
    os.environ["KB_PAIRING_WARP_MAX"] = "0"
    ct = _ffi.Context(0)
    nmax = max(sizes)
    one = np.zeros(8, np.uint32); one[0] = 7
    cw.srs_generate(one, 64, download=False)
    ct.srs_generate(one, 64, download=False)
    g1, i1 = cw.g1_mul_gen_batch(rand_fr(nmax))
    off = np.arange(nmax + 1, dtype=np.uint64) * 32
    com, _ = cw.g1_mul_gen_batch(rand_fr(1))
    g2, i2, msg_ct = cw.encrypt_batch(com[0], 0, rand_fr(nmax), rand_fr(nmax), rand_fr(nmax), np.zeros(32 * nmax, np.uint8), off)
    print("n      warp_ms  thread_ms   (decrypt, device time of the call)")
    for n in sizes:
        row = []
        for c in (cw, ct):
            best = 1e9
            for _ in range(3):
                c.decrypt_batch(g1[:n], i1[:n], g2[:n], i2[:n], msg_ct, off[: n + 1], n=n)
                best = min(best, c.last_kernel_ms(0))
            row.append(best)
        print("%-6d %8.3f %9.3f" % (n, row[0], row[1]), flush=True)
    print("cold kb_encrypt_batch (fresh commitment each call, 64 messages): device ms / wall ms")
    for name, c in (("warp", cw), ("thread", ct)):
        res = []
        for k in range(4):
            cm, _ = c.g1_mul_gen_batch(rand_fr(1))
            t0 = time.perf_counter()
            c.encrypt_batch(cm[0], 0, rand_fr(64), rand_fr(64), rand_fr(64), np.zeros(32 * 64, np.uint8), off[:65])
            res.append((c.last_kernel_ms(0), (time.perf_counter() - t0) * 1e3))
        print(name, " ".join("%.2f/%.2f" % r for r in res), flush=True)


if __name__ == "__main__":
    main()
