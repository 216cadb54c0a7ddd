#!/usr/bin/env python3
"""One decrypt of N pairings as ONE persistent launch vs back-to-back launches of exactly one round (148 SMs x 8 warps x 32)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from keaki_b200 import _ffi  # noqa: E402

c = _ffi.Context(0)
one = np.zeros(8, np.uint32); one[0] = 7
c.srs_generate(one, 64, download=False)
N = int(os.environ.get("N", str(1 << 18)))
rng = np.random.default_rng(1)
k = rng.integers(0, 2**32, size=(N, 8), dtype=np.uint64).astype(np.uint32); k[:, 7] &= 0x0FFFFFFF
g1, i1 = c.g1_mul_gen_batch(k)
off = np.arange(N + 1, dtype=np.uint64) * 32
com, _ = c.g1_mul_gen_batch(k[:1])
g2, i2, mc = c.encrypt_batch(com[0], 0, k, k, k, np.zeros(32 * N, np.uint8), off)
import torch
dev = torch.device("cuda:0")
tg1, ti1, tg2, ti2, tmc, toff = (torch.from_numpy(x).to(dev) for x in (g1, i1, g2, i2, mc, off.view(np.int64)))
out = torch.zeros(32 * N, dtype=torch.uint8, device=dev)
P = _ffi._ptr


def run(lo, hi):
    n = hi - lo
    o = (toff[lo:hi + 1] - toff[lo]).contiguous()
    c._check(c.lib.kb_decrypt_batch(c.h, P(tg1[lo:hi]), P(ti1[lo:hi]), P(tg2[lo:hi]), P(ti2[lo:hi]), P(tmc[32 * lo:]), P(o), n, P(out[32 * lo:])))
    return c.last_kernel_ms(2)


for chunk in (N, 37888, 2 * 37888, 18944):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dev_ms = sum(run(lo, min(N, lo + chunk)) for lo in range(0, N, chunk))
        torch.cuda.synchronize()
        print("N=%d chunk=%d: kernels %.2f ms, wall %.2f ms" % (N, chunk, dev_ms, (time.perf_counter() - t0) * 1e3), flush=True)
