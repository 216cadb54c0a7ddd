"""Small pass over every entry point for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): MSM at both window
widths (two-digit reduction, quad-cooperative tail) incl. the over-full-bucket path (all-equal scalars), SRS validation,
open-all (radix-8 passes), encrypt (fresh commitment: the per-commitment pairing and window bases on the warp-cooperative
interpreter, the one-warp-per-message kernel), decrypt on the warp-cooperative kernel, on the compiled thread kernel and on
the pairing VM, the thread-per-message encrypt kernels, verify, wire format."""
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
same = np.ascontiguousarray(np.tile(fr_to_limbs(0x1234567890ABCDEF1234567890ABCDEF % R), (1 << 12, 1)))
xy, inf = ctx.msm_g1(same)          # every window lands in one bucket: segments + both folds
ctx.srs_validate()
coeffs = rnd(64)
com, ci = ctx.msm_g1(coeffs)
proofs, pinf = ctx.open_all_fk(coeffs)
n = 64
vals = rnd(n); vals[:16] = fr_to_limbs(0); vals[16:32] = fr_to_limbs(1)
msgs = rng.integers(0, 256, size=n * 32, dtype=np.uint8); off = np.arange(n + 1, dtype=np.uint64) * 32
pts = rnd(n)
ct, cti, mc = ctx.encrypt_batch(com, ci, pts, vals, rnd(n), msgs, off)
out = ctx.decrypt_batch(proofs, pinf, ct, cti, mc, off)
assert bytes(out[: n * 32]) != b"" 
import os
os.environ["KB_PAIRING_WARP_MAX"] = "0"
os.environ["KB_ENCRYPT_WARP_MAX"] = "0"
th_ctx = _ffi.Context(0)          # the thread-per-pairing / thread-per-message kernels on the same inputs
th_ctx.srs_generate(fr_to_limbs(tau), 1 << 14, download=False)
rs_same = rnd(n)
a1 = ctx.encrypt_batch(com, ci, pts, vals, rs_same, msgs, off)
a2 = th_ctx.encrypt_batch(com, ci, pts, vals, rs_same, msgs, off)
assert all(np.array_equal(x, y) for x, y in zip(a1, a2))
assert np.array_equal(out, th_ctx.decrypt_batch(proofs, pinf, ct, cti, mc, off))
th_ctx.close()
os.environ.pop("KB_PAIRING_WARP_MAX"); os.environ.pop("KB_ENCRYPT_WARP_MAX")
os.environ["KB_PAIRING_IMPL"] = "vm"
vm_ctx = _ffi.Context(0)
out_vm = vm_ctx.decrypt_batch(proofs, pinf, ct, cti, mc, off)
assert np.array_equal(out, out_vm)
vm_ctx.close()
# above the small-batch thresholds: the GT encrypt kernel on the pairing's shared-memory machinery, the thread-per-message G2
# kernel, and the SEGMENTED pairing launches (more than one round of resident warps)
nb = 2100
ptsb, valsb = rnd(nb), rnd(nb); valsb[:700] = fr_to_limbs(0); valsb[700:1400] = fr_to_limbs(1)
offb = np.arange(nb + 1, dtype=np.uint64) * 32
ctb, ctib, mcb = ctx.encrypt_batch(com, ci, ptsb, valsb, rnd(nb), rng.integers(0, 256, size=nb * 32, dtype=np.uint8), offb)
nbig = 148 * 256 + 200
g1b, i1b = ctx.g1_mul_gen_batch(rnd(64))
g1big = np.ascontiguousarray(np.tile(g1b, (nbig // 64 + 1, 1))[:nbig]); i1big = np.zeros(nbig, np.uint8); i1big[[0, 37888, nbig - 1]] = 1
g2big = np.ascontiguousarray(np.tile(ctb[:64], (nbig // 64 + 1, 1))[:nbig]); i2big = np.zeros(nbig, np.uint8)
gt_big = ctx.pairing_batch(g1big, i1big, g2big, i2big)
assert np.array_equal(gt_big[64:128], gt_big[128:192]) and np.array_equal(gt_big[1], gt_big[37888 + 64 - 37888 % 64 + 1])
ok = ctx.verify_batch(np.tile(com, (n, 1)), np.zeros(n, np.uint8), pts, vals, proofs, pinf)
b = ctx.g2_serialize(ct, cti, True)
back = ctx.g2_deserialize(b, True, True)
assert back[2].all()
print("sanitizer flow done", int(ok.sum()))
