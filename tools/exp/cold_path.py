#!/usr/bin/env python3
"""One cold kb_encrypt_batch (fresh commitment) + one single decrypt, for a kernel launch list under ncu:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cold_launches.csv python tools/exp/cold_path.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from keaki_b200 import _ffi  # noqa: E402

c = _ffi.Context(0)
one = np.zeros(8, np.uint32); one[0] = 7
c.srs_generate(one, 64, download=False)
k = np.zeros((1, 8), np.uint32); k[0, 0] = 12345
for rep in range(2):
    k[0, 1] = rep + 1
    com, _ = c.g1_mul_gen_batch(k)
    n = 64
    off = np.arange(n + 1, dtype=np.uint64) * 32
    s = np.zeros((n, 8), np.uint32); s[:, 0] = np.arange(n) + 3
    ct, ci, mc = c.encrypt_batch(com[0], 0, s, s, s, np.zeros(32 * n, np.uint8), off)
    g1, i1 = c.g1_mul_gen_batch(s)
    c.decrypt_batch(g1[:1], i1[:1], ct[:1], ci[:1], mc, off[:2], n=1)
print("done")
