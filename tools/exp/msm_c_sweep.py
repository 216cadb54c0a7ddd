#!/usr/bin/env python3
"""Times kb_msm_g1 at n = 2^LOG for window widths c (KB_MSM_C override): which c the window rule should pick per size.
Run on the GPU box: python tools/exp/msm_c_sweep.py 16 14 15 16 17 18 19 20"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from keaki_b200 import _ffi  # noqa: E402

log_n = int(sys.argv[1])
cs = [int(x) for x in sys.argv[2:]] or [0]
n = 1 << log_n
rng = np.random.default_rng(3)
a = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
a[:, 7] &= 0x0FFFFFFF
ref = None
for c in cs:
    if c:
        os.environ["KB_MSM_C"] = str(c)
    else:
        os.environ.pop("KB_MSM_C", None)
    ctx = _ffi.Context(0)
    ctx.srs_generate(a[0], max(n, 1 << int(os.environ.get('SWEEP_SRS_LOG', '0'))), download=False)
    d = torch.from_numpy(a).cuda()
    for _ in range(4):
        out = ctx.msm_g1(d, n=n)
    tot, acc = [], []
    t = time.perf_counter()
    for _ in range(20):
        out = ctx.msm_g1(d, n=n)
        tot.append(ctx.last_kernel_ms(0)); acc.append(ctx.last_kernel_ms(1))
    wall = (time.perf_counter() - t) / 20 * 1e3
    if ref is None:
        ref = out
    same = bool(np.array_equal(ref[0], out[0]))
    print(f"n=2^{log_n} c={c or 'default'}: call {np.mean(tot):.3f} ms (accumulate {np.mean(acc):.3f} ms), host wall {wall:.3f} ms, same={same}", flush=True)
    ctx.close()
