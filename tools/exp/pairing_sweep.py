#!/usr/bin/env python3
"""Times kb_decrypt_batch (one pairing + hash per message) at 2^16 for the pairing kernel variants:
KB_PAIRING_IMPL=vm (two lanes + interpreter) / st (compiled single-thread) and the st launch shapes.
Also checks that every variant returns the same bytes.  Run on the GPU box: python tools/exp/pairing_sweep.py [log_n]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from keaki_b200 import _ffi  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << log_n
rng = np.random.default_rng(7)


def rand_fr(k):
    a = rng.integers(0, 1 << 32, size=(k, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= 0x0FFFFFFF
    return a


configs = [("vm", None), ("st", "0"), ("st", "8"), ("st", "7"), ("st", "6"), ("st", "5")]
if os.environ.get("SWEEP_ONLY"):   # e.g. SWEEP_ONLY=st:0
    impl_, shape_ = os.environ["SWEEP_ONLY"].split(":")
    configs = [(impl_, shape_ if shape_ != "" else None)]
reps = int(os.environ.get("SWEEP_REPS", "4"))
ref = None
inputs = None
for impl, shape in configs:
    os.environ["KB_PAIRING_IMPL"] = impl
    if shape is not None:
        os.environ["KB_PAIRING_ST_SHAPE"] = shape
    ctx = _ffi.Context(0)
    if inputs is None:
        tau = rand_fr(1)[0]
        ctx.srs_generate(tau, 16, download=False)
        g1, g1i = ctx.g1_mul_gen_batch(rand_fr(n))
        # G2 points: ciphertexts of an encryption batch (any valid G2 points do)
        com, ci = ctx.msm_g1(rand_fr(16), n=16)
        off = (np.arange(n + 1, dtype=np.uint64) * 32)
        msgs = rng.integers(0, 256, size=n * 32, dtype=np.uint8)
        ct, cti, mc = ctx.encrypt_batch(com, ci, rand_fr(n), rand_fr(n), rand_fr(n), msgs, off)
        inputs = (g1, g1i, ct, cti, mc, off)
    else:
        ctx.srs_generate(tau, 16, download=False)
    g1, g1i, ct, cti, mc, off = inputs
    times = []
    for rep in range(reps):
        t = time.perf_counter()
        out = ctx.decrypt_batch(g1, g1i, ct, cti, mc, off)
        times.append((time.perf_counter() - t) * 1e3)
        kms = ctx.last_kernel_ms(2)
    if ref is None:
        ref = out.copy()
    same = bool(np.array_equal(ref, out))
    print(f"impl={impl} shape={shape} n=2^{log_n}: pairing kernel {kms:.3f} ms (host wall min {min(times):.2f} ms), same_bytes={same}", flush=True)
    ctx.close()
