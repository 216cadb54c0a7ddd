#!/usr/bin/env python3
"""Device time of kb_msm_g1 at 2^20 (uniform scalars below r, resident), and structured scalars."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from keaki_b200 import _ffi  # noqa: E402
import torch

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
c = _ffi.Context(0)
one = np.zeros(8, np.uint32); one[0] = 7
n = 1 << int(os.environ.get("LOGN", "20"))
c.srs_generate(one, n, download=False)
rng = np.random.default_rng(1)
raw = rng.integers(0, 2**32, size=(n, 10), dtype=np.uint64)
vals = [(int.from_bytes(raw[i].astype(np.uint32).tobytes(), "little") % R) for i in range(n)] if n <= (1 << 16) else None
if vals is None:   # fast path: rejection-free reduction of 320-bit values is slow in Python; use 253-bit values below r/2 plus a random top
    k = rng.integers(0, 2**32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    k[:, 7] = rng.integers(0, 0x30644E72, size=n, dtype=np.uint64).astype(np.uint32)   # uniform top limb below r's
else:
    k = np.array([[(v >> (32 * j)) & 0xFFFFFFFF for j in range(8)] for v in vals], np.uint32)
dev = torch.device("cuda:0")
kd = torch.from_numpy(k).to(dev)
out = torch.zeros(17, dtype=torch.int32, device=dev)
for name, src in (("uniform", kd), ("all equal", kd[:1].repeat(n, 1).contiguous()), ("0/1", torch.from_numpy((rng.integers(0, 2, size=(n, 1)) * np.array([[1, 0, 0, 0, 0, 0, 0, 0]])).astype(np.uint32)).to(dev))):
    t = []
    for _ in range(6):
        c._check(c.lib.kb_msm_g1(c.h, _ffi._ptr(src), 0, n, out.data_ptr(), out.data_ptr() + 64))
        t.append((c.last_kernel_ms(0), c.last_kernel_ms(1)))
    print(name, " ".join("%.3f/%.3f" % x for x in t[2:]), flush=True)
