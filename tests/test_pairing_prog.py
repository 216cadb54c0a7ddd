"""The pairing VM programs (tools/gen_pairing_prog.py -> keaki_b200/csrc/pairing_prog_gen.cuh) vs the oracle.

1. the generator's own integer simulator reproduces the oracle's `pairing` for every slot budget;
2. the header on disk is what the generator emits now (not stale);
3. the very interpreter the GPU runs (pairing_vm.cuh, compiled for the host with emulated carry
   flags) running the shipped program gives the oracle's 384 GT bytes - including the unreduced
   (< 2p) operand sums the multiplier is fed."""
import ctypes
import os
import random
import sys

import pytest

from oracle import bn254 as bn
from tests import limbs as L
from tests.hostemu import lib as HE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_pairing_prog as gp  # noqa: E402

rng = random.Random(0x5107)
he = HE.load()
P = HE.ptr


def _flat(gt):
    return [gt[0][0], gt[0][1], gt[0][2], gt[1][0], gt[1][1], gt[1][2]]


def _cases(n):
    out = [(bn.G1_GEN, bn.G2_GEN)]
    for _ in range(n):
        out.append((bn.g1_mul(bn.G1_GEN, rng.randrange(1, bn.R)), bn.g2_mul(bn.G2_GEN, rng.randrange(1, bn.R))))
    return out


@pytest.mark.parametrize("slots,macros", [(14, True), (18, True), (28, True), (14, False)])
def test_simulated_program_matches_oracle(slots, macros):
    gp.USE_MACROS = macros
    try:
        words, gcount, outs, st = gp.build("pairing", slots)
    finally:
        gp.USE_MACROS = True
    assert st["fq_muls"] < 18000 and all(s < slots for s in outs)
    for p, q in _cases(1):
        got = gp.simulate(words, slots, outs, [(p[0], p[1]), q[0], q[1]])
        assert got == _flat(bn.pairing(p, q))


def test_generated_header_is_current():
    path = os.path.join(ROOT, "keaki_b200", "csrc", "pairing_prog_gen.cuh")
    text = open(path).read()
    for what, slots in gp.VARIANTS:
        words, gcount, outs, st = gp.build(what, slots)
        name = gp.variant_name(what, slots)
        assert "{%d, %d, %d, {%s}, %s_WORDS}" % (slots, gcount, len(words), ", ".join(map(str, outs)), name) in text
        assert "0x%016xull" % words[len(words) // 2] in text


def test_constants_match_oracle():
    for k in (1, 2, 3):
        for i in range(6):
            assert gp.CONSTS[gp.C_FROB[k - 1][i]] == bn.f2_pow(bn.XI, i * (bn.Q**k - 1) // 6)
    assert sum(d << i for i, d in enumerate(gp.Z_WNAF)) == bn.Z


@pytest.mark.parametrize("slots", [14, 28])
def test_hostemu_interpreter_matches_oracle(slots):
    for p, q in _cases(2):
        out = (ctypes.c_uint8 * 384)()
        n = he.he_vm_pairing_bytes(slots, P(L.g1_m(p)), P(L.g2_m(q)), out)
        assert n > 1000
        assert bytes(out) == bn.gt_to_bytes(bn.pairing(p, q))


def test_mul9_add_quotient_estimate():
    import numpy as np
    cases = [(0, 0), (bn.Q - 1, bn.Q), (bn.Q - 1, 0), (0, bn.Q), (1, bn.Q - 1)]
    # values around every multiple of p that 9x + y can straddle
    for k in range(1, 10):
        for dx in (-1, 0, 1):
            x = min(bn.Q - 1, max(0, (k * bn.Q) // 9 + dx))
            for y in (0, 1, bn.Q - 1, bn.Q, (k * bn.Q - 9 * x) % (bn.Q + 1)):
                cases.append((x, y))
    cases += [(rng.randrange(bn.Q), rng.randrange(bn.Q + 1)) for _ in range(2000)]
    out = np.zeros(8, np.uint32)
    for x, y in cases:
        he.he_mul9_add(P(L.int_to_limbs(x)), P(L.int_to_limbs(y)), P(out))
        assert L.limbs_to_int(out) == (9 * x + y) % bn.Q, (x, y)
