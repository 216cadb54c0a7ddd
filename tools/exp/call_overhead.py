#!/usr/bin/env python3
"""Host-side overhead of small C-ABI calls: wall time vs device time of the call."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from keaki_b200 import _ffi  # noqa: E402

c = _ffi.Context(0)
one = np.zeros(8, np.uint32); one[0] = 7
c.srs_generate(one, 64, download=False)
k = np.zeros((64, 8), np.uint32); k[:, 0] = np.arange(64) + 3
com, _ = c.g1_mul_gen_batch(k[:1])
off = np.arange(65, dtype=np.uint64) * 32
msgs = np.zeros(32 * 64, np.uint8)


def t(name, f, reps=20):
    f()
    w, dv = [], []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); w.append((time.perf_counter() - t0) * 1e3); dv.append(c.last_kernel_ms(0))
    print("%-34s wall %.3f ms   device %.3f ms" % (name, np.median(w), np.median(dv)), flush=True)


pts = k[:1].copy()
t("g1_sum(1 point)", lambda: c.g1_sum(com[:1]))
t("msm_g1(n=1)", lambda: c.msm_g1(k[:1], n=1))
t("msm_g1(n=64)", lambda: c.msm_g1(k, n=64))
for n in (1, 64):
    ct = c.encrypt_batch(com[0], 0, k[:n], k[:n], k[:n], msgs, off[: n + 1])
    t("encrypt_batch(n=%d)" % n, lambda: c.encrypt_batch(com[0], 0, k[:n], k[:n], k[:n], msgs, off[: n + 1]))
    g1, i1 = c.g1_mul_gen_batch(k[:n])
    t("decrypt_batch(n=%d)" % n, lambda: c.decrypt_batch(g1, i1, ct[0], ct[1], ct[2], off[: n + 1], n=n))
    t("pairing_batch(n=%d)" % n, lambda: c.pairing_batch(g1, i1, ct[0], ct[1]))
