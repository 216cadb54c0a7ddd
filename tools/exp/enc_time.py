#!/usr/bin/env python3
"""Device time of kb_encrypt_batch at 2^16 (bit values, warm 16-bit tables): total and the split reported by the timers."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from keaki_b200 import _ffi  # noqa: E402

c = _ffi.Context(0)
one = np.zeros(8, np.uint32); one[0] = 7
c.srs_generate(one, 64, download=False)
n = 1 << 16
rng = np.random.default_rng(1)
k = rng.integers(0, 2**32, size=(n, 8), dtype=np.uint64).astype(np.uint32); k[:, 7] &= 0x0FFFFFFF
vals = np.zeros((n, 8), np.uint32)
off = np.arange(n + 1, dtype=np.uint64) * 32
com, _ = c.g1_mul_gen_batch(k[:1])
msgs = np.zeros(32 * n, np.uint8)
for rep in range(6):
    c.encrypt_batch(com[0], 0, k, vals, k, msgs, off)
    print("encrypt 2^16: call %.3f ms, kernels %.3f ms" % (c.last_kernel_ms(0), c.last_kernel_ms(3)), flush=True)
