import sys, time
sys.path.insert(0, '.')
import numpy as np
from keaki_b200 import _ffi
from keaki_b200.types import fr_to_limbs, FR_MODULUS as R
ctx = _ffi.Context(0)
tau = 0x123456789ABCDEF123456789ABCDEF % R
ctx.srs_generate(fr_to_limbs(tau), 64, download=False)
rng = np.random.default_rng(1)
def rnd(n):
    a = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); a[:, 7] &= 0x0FFFFFFF; return a
for logn in (12, 14):
    n = 1 << logn
    com, ci = ctx.g1_mul_gen_batch(rnd(n)); pr, pi = ctx.g1_mul_gen_batch(rnd(n))
    pts, vals = rnd(n), rnd(n)
    ctx.verify_batch(com[:16], ci[:16], pts[:16], vals[:16], pr[:16], pi[:16])
    t = time.perf_counter(); ok = ctx.verify_batch(com, ci, pts, vals, pr, pi); dt = time.perf_counter() - t
    print("verify 2^%d: %.1f ms (%.0f/s), kernel %.1f ms, any ok %s" % (logn, dt * 1e3, n / dt, ctx.last_kernel_ms(0), ok.any()))
