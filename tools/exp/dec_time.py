#!/usr/bin/env python3
"""Device time of kb_decrypt_batch at 2^16 (and other sizes): KB_PAIRING_NO_CAP=1 for the unrestricted dealing."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from keaki_b200 import _ffi  # noqa: E402

c = _ffi.Context(0)
one = np.zeros(8, np.uint32); one[0] = 7
c.srs_generate(one, 64, download=False)
sizes = [int(x) for x in os.environ.get("SIZES", "65536,49152,40000,131072").split(",")]
nmax = max(sizes)
rng = np.random.default_rng(1)
k = rng.integers(0, 2**32, size=(nmax, 8), dtype=np.uint64).astype(np.uint32); k[:, 7] &= 0x0FFFFFFF
g1, i1 = c.g1_mul_gen_batch(k)
off = np.arange(nmax + 1, dtype=np.uint64) * 32
com, _ = c.g1_mul_gen_batch(k[:1])
g2, i2, mc = c.encrypt_batch(com[0], 0, k, k, k, np.zeros(32 * nmax, np.uint8), off)
for n in sizes:
    t = []
    for _ in range(4):
        c.decrypt_batch(g1[:n], i1[:n], g2[:n], i2[:n], mc, off[: n + 1], n=n)
        t.append(c.last_kernel_ms(2))
    print(n, " ".join("%.2f" % x for x in t), flush=True)
