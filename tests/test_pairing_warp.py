"""The warp-cooperative pairing schedules (tools/gen_pairing_warp.py -> keaki_b200/csrc/pairing_warp_gen.cuh) vs the oracle.

1. the generator's integer simulator reproduces the oracle's `pairing` and the window-base chain;
2. the header on disk is what the generator emits now;
3. the very interpreter the GPU runs (pairing_warp.cuh on the host, emulated carry flags, lanes one after the other with
   the GPU's read-all-then-store step semantics) gives the oracle's GT for the shipped schedules;
4. the 288-bit reduction at the extremes of its input range."""
import os
import random
import subprocess
import sys

import numpy as np

from oracle import bn254 as bn
from tests import limbs as L
from tests.hostemu import lib as HE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_pairing_warp as gw  # noqa: E402

rng = random.Random(0xA11CE)
he = HE.load()
P = HE.ptr


def _flat(gt):
    return [gt[0][0], gt[0][1], gt[0][2], gt[1][0], gt[1][1], gt[1][2]]


def _mont2(x):
    return list(L.f2_m(x))


def _from_mont2(w):
    return L.f2_from(np.asarray(w, dtype=np.uint32))


def _pair_inputs(p, q):
    return [(p[0], p[1]), q[0], q[1], (p[0], 0), (p[1], 0)]


def test_simulated_schedules_match_oracle():
    prog = gw.build("pairing")
    assert prog["stats"]["hist"]["MUL"] < 700 and prog["nslots"] <= 256
    p, q = bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R)), bn.g2_mul(bn.G2_GEN, rng.randrange(1, bn.R))
    assert gw.simulate(prog, _pair_inputs(p, q)) == _flat(bn.pairing(p, q))
    bases = gw.build("gt_bases")
    a = bn.pairing(bn.G1_GEN, bn.G2_GEN)
    got = gw.simulate(bases, _flat(a))
    x = a
    for w in range(32):
        assert got[6 * w:6 * w + 6] == _flat(x), w
        for _ in range(8):
            x = bn.f12_sqr(x)


def test_header_is_current(tmp_path):
    path = os.path.join(ROOT, "keaki_b200", "csrc", "pairing_warp_gen.cuh")
    before = open(path).read()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_pairing_warp.py")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert open(path).read() == before, "pairing_warp_gen.cuh is stale: run tools/gen_pairing_warp.py"


def test_host_interpreter_runs_shipped_pairing():
    cases = [(bn.G1_GEN, bn.G2_GEN), (bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R)), bn.g2_mul(bn.G2_GEN, rng.randrange(1, bn.R)))]
    for p, q in cases:
        inp = HE.u32([w for v in _pair_inputs(p, q) for w in _mont2(v)])
        out = np.zeros(6 * 16, dtype=np.uint32)
        he.he_wp_run(0, P(inp), P(out))
        got = [_from_mont2(list(out[16 * i:16 * i + 16])) for i in range(6)]
        assert got == _flat(bn.pairing(p, q))


def test_host_interpreter_runs_shipped_window_bases():
    a = bn.pairing(bn.g1_mul(bn.G1_GEN, 7), bn.G2_GEN)
    inp = HE.u32([w for v in _flat(a) for w in _mont2(v)])
    out = np.zeros(32 * 6 * 16, dtype=np.uint32)
    he.he_wp_run(1, P(inp), P(out))
    x = a
    for w in range(32):
        got = [_from_mont2(list(out[16 * (6 * w + i):16 * (6 * w + i) + 16])) for i in range(6)]
        assert got == _flat(x), w
        for _ in range(8):
            x = bn.f12_sqr(x)


def test_gt_product_schedule():
    """GT_PROD: the product tree of 48 Fq12 operands behind the small-batch encrypt kernel (ones included, as for zero digits)"""
    prog = gw.build("gt_prod")
    assert prog["nslots"] <= gw.MAX_SLOTS and len(prog["inputs"]) == 48 * 6
    one = ((((1, 0), (0, 0), (0, 0)), ((0, 0), (0, 0), (0, 0))))
    xs = []
    for k in range(48):
        if k % 5 == 3:
            xs.append(one)
        else:
            xs.append(tuple(tuple((rng.randrange(bn.Q), rng.randrange(bn.Q)) for _ in range(3)) for _ in range(2)))
    want = xs[0]
    for x in xs[1:]:
        want = bn.f12_mul(want, x)
    flat_in = [c for x in xs for c in _flat(x)]
    assert gw.simulate(prog, flat_in) == _flat(want)
    inp = HE.u32([w for v in flat_in for w in _mont2(v)])
    out = np.zeros(6 * 16, dtype=np.uint32)
    he.he_wp_run(2, P(inp), P(out))
    assert [_from_mont2(list(out[16 * i:16 * i + 16])) for i in range(6)] == _flat(want)


def test_reduce9_extremes():
    q = bn.Q
    vals = [0, 1, -1, q - 1, q, q + 1, -q, 127 * q, 128 * q - 1, -128 * q + 1, -127 * q - 1, 64 * q + (q >> 1)]
    # values whose top bits sit right at a multiple of the divisor of the quotient estimate
    for k in range(1, 250, 7):
        vals += [k * q - 128 * q, k * q - 128 * q - 1, k * q - 128 * q + 1]
    vals += [rng.randrange(-128 * q + 1, 128 * q) for _ in range(300)]
    for v in vals:
        if not -128 * q < v < 128 * q:
            continue
        tc = v % (1 << 288)
        v9 = HE.u32([(tc >> (32 * i)) & 0xFFFFFFFF for i in range(9)])
        out = np.zeros(8, dtype=np.uint32)
        he.he_wp_reduce9(P(v9), P(out))
        got = sum(int(out[i]) << (32 * i) for i in range(8))
        assert got == v % q, hex(v)
