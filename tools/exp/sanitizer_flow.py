"""Small pass over every entry point for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): MSM at three window widths
(two-digit reduction, quad-cooperative tail), open-all, encrypt (both per-commitment tables), decrypt on the pairing VM,
verify, wire format."""
import sys
sys.path.insert(0, ".")
import numpy as np
from keaki_b200 import _ffi
from keaki_b200.types import fr_to_limbs, FR_MODULUS as R

ctx = _ffi.Context(0)
rng = np.random.default_rng(3)


def rnd(n):
    a = rng.integers(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); a[:, 7] &= 0x0FFFFFFF; return a


tau = 0x1234567 % R
ctx.srs_generate(fr_to_limbs(tau), 1 << 14, download=False)
for n in (100, 5000, 1 << 14):
    xy, inf = ctx.msm_g1(rnd(n))
coeffs = rnd(64)
com, ci = ctx.msm_g1(coeffs)
proofs, pinf = ctx.open_all_fk(coeffs)
n = 64
vals = rnd(n); vals[:16] = fr_to_limbs(0); vals[16:32] = fr_to_limbs(1)
msgs = rng.integers(0, 256, size=n * 32, dtype=np.uint8); off = np.arange(n + 1, dtype=np.uint64) * 32
pts = rnd(n)
ct, cti, mc = ctx.encrypt_batch(com, ci, pts, vals, rnd(n), msgs, off)
out = ctx.decrypt_batch(proofs, pinf, ct, cti, mc, off)
ok = ctx.verify_batch(np.tile(com, (n, 1)), np.zeros(n, np.uint8), pts, vals, proofs, pinf)
b = ctx.g2_serialize(ct, cti, True)
back = ctx.g2_deserialize(b, True, True)
assert back[2].all()
print("sanitizer flow done", int(ok.sum()))
