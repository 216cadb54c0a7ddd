"""Host-side logic of the keaki mirror (no GPU): types, domains, ptau container, error behaviour."""
import os
import random

import numpy as np
import pytest

from keaki_b200 import kzg, ptau
from keaki_b200.types import (FQ_MODULUS, FR_MODULUS, FrRng, G1, G2, Radix2EvaluationDomain, fr_array, fr_from_limbs,
                              fr_list, fr_to_limbs, pack_g1, unpack_g1)
from oracle import bn254 as bn
from tests import limbs as L

GOLD = os.path.join(os.path.dirname(__file__), "golden")
rng = random.Random(7)


def test_moduli_and_limb_conversions_match_oracle():
    assert FQ_MODULUS == bn.Q and FR_MODULUS == bn.R
    for x in [0, 1, bn.R - 1, rng.randrange(bn.R)]:
        assert np.array_equal(fr_to_limbs(x), L.fr_m(x)) and fr_from_limbs(fr_to_limbs(x)) == x
    xs = [rng.randrange(bn.R) for _ in range(17)]
    assert fr_list(fr_array(xs)) == xs and fr_array([]).shape == (0, 8)


def test_domain_matches_ark_poly_conventions():
    for n in (1, 2, 3, 8, 9, 1000):
        d, o = Radix2EvaluationDomain(n), bn.Radix2Domain(n)
        assert d.size == o.size and d.group_gen == o.group_gen and d.elements() == o.elements()
    assert Radix2EvaluationDomain((1 << 20) + 1).size == 1 << 21   # "2^20 bits + padding rounds up" 
    with pytest.raises(ValueError):
        Radix2EvaluationDomain((1 << 28) + 1)


def test_rng_is_deterministic_and_in_range():
    a, b = FrRng(5), FrRng(5)
    xs = [a.fr() for _ in range(50)]
    assert xs == [b.fr() for _ in range(50)] and all(0 <= x < FR_MODULUS for x in xs)
    assert FrRng(6).fr() != xs[0]


def test_default_rng_is_the_os_csprng_and_contexts_resolve_clearly():
    """rng=None must never mean a fixed seed: SecureFrRng draws from os.urandom; the seeded generator needs an
    explicit seed; the setup-free calls fail with a clear message when no KZGSetup exists"""
    from keaki_b200 import kem
    from keaki_b200.types import SecureFrRng, SeededFrRng, rng_or_secure
    assert isinstance(rng_or_secure(None), SecureFrRng)
    r = FrRng(1)
    assert rng_or_secure(r) is r and isinstance(r, SeededFrRng)
    xs = {SecureFrRng().fr() for _ in range(20)}
    assert len(xs) == 20 and all(0 <= x < FR_MODULUS for x in xs)
    with pytest.raises(TypeError):
        FrRng()
    kzg._latest_ctx = None
    with pytest.raises(RuntimeError, match="no live KZGSetup"):
        kem.decapsulate(G1.generator(), G2.zero(), 32)


def test_ptau_reader_is_as_strict_as_the_reference(tmp_path):
    """src/kzg/ptau.rs:251-256,299-304 (exact section sizes), :146-150 (missing sections), bounded power"""
    import struct
    g1, g2 = ptau.get_powers_from_file(os.path.join(GOLD, "ppot_0080_01_mini.ptau"))
    good = tmp_path / "good.ptau"
    ptau.write_ptau(str(good), g1, g2, power=1)
    data = good.read_bytes()

    def sections(blob):
        off, out = 12, []
        for _ in range(11):
            sid, slen = struct.unpack_from("<IQ", blob, off)
            out.append((sid, blob[off + 12: off + 12 + slen]))
            off += 12 + slen
        return out

    def build(secs):
        return data[:12] + b"".join(struct.pack("<IQ", sid, len(body)) + body for sid, body in secs)

    secs = sections(data)
    assert build(secs) == data
    # a longer TauG1 section (one extra element) is ElementSizeMismatch, not silently accepted
    longer = [(sid, body + bytes(64) if sid == 2 else body) for sid, body in secs]
    p = tmp_path / "longer.ptau"; p.write_bytes(build(longer))
    with pytest.raises(ptau.SetupFileError, match="ElementSizeMismatch"):
        ptau.get_powers_from_file(str(p))
    shorter = [(sid, body[:-128] if sid == 3 else body) for sid, body in secs]
    p = tmp_path / "shorter.ptau"; p.write_bytes(build(shorter))
    with pytest.raises(ptau.SetupFileError, match="ElementSizeMismatch"):
        ptau.get_powers_from_file(str(p))
    # TauG2 section replaced by a second copy of section 4: section 3 is missing -> EmptySection(3), not a KeyError
    dup = [((4, b"") if sid == 3 else (sid, body)) for sid, body in secs]
    p = tmp_path / "dup.ptau"; p.write_bytes(build(dup))
    with pytest.raises(ptau.SetupFileError, match="EmptySection"):
        ptau.get_powers_from_file(str(p))
    # absurd power in the header
    hdr = secs[0][1]
    big = [(1, hdr[:36] + struct.pack("<I", 200) + hdr[40:])] + secs[1:]
    p = tmp_path / "big.ptau"; p.write_bytes(build(big))
    with pytest.raises(ptau.SetupFileError):
        ptau.get_powers_from_file(str(p))
    short_hdr = [(1, hdr[:20])] + secs[1:]
    p = tmp_path / "hdr.ptau"; p.write_bytes(build(short_hdr))
    with pytest.raises(ptau.SetupFileError):
        ptau.get_powers_from_file(str(p))


def test_points_equality_and_packing():
    g = G1.generator()
    assert g.to_affine_ints() == (1, 2) and G1.zero().inf and G1.zero() == G1(None)
    assert G1(g.xy) == g and G1(g.xy) != G1.zero() and G2.zero().to_affine_ints() is None
    xy, inf = pack_g1([g, G1.zero()])
    assert inf.tolist() == [0, 1] and unpack_g1(xy, inf) == [g, G1.zero()]


def test_ptau_reader_on_reference_fixture_sections(tmp_path):
    g1, g2 = ptau.get_powers_from_file(os.path.join(GOLD, "ppot_0080_01_mini.ptau"))
    assert g1.shape == (3, 16) and g2.shape == (2, 32)               # src/kzg/ptau.rs:476-514
    # the file's limbs ARE Montgomery limbs: passing them through is the correct decoding
    assert L.g1_from(g1[0]) == bn.G1_GEN and L.g2_from(g2[0]) == bn.G2_GEN
    # writer -> reader round trip, and container errors (src/kzg/ptau.rs:360-376)
    out = tmp_path / "w.ptau"
    ptau.write_ptau(str(out), g1, g2, power=1)
    a, b = ptau.get_powers_from_file(str(out))
    assert np.array_equal(a, g1) and np.array_equal(b, g2)
    bad = tmp_path / "bad.ptau"
    bad.write_bytes(b"nope" + out.read_bytes()[4:])
    with pytest.raises(ptau.SetupFileError):
        ptau.get_powers_from_file(str(bad))
    with pytest.raises(ptau.SetupFileError):
        ptau.get_powers_from_file(str(tmp_path / "missing.ptau"))
    trunc = tmp_path / "trunc.ptau"
    trunc.write_bytes(out.read_bytes()[:-5])
    with pytest.raises(ptau.SetupFileError):
        ptau.get_powers_from_file(str(trunc))


def test_commit_size_check_happens_before_any_gpu_work():
    """KZGError::PolynomialTooLarge(p.len(), g1_pow.len()) — src/kzg.rs:93-95; trailing zeros are stripped
    like DensePolynomial::from_coefficients_slice does."""
    setup = kzg.KZGSetup(None, np.zeros((2, 16), np.uint32), G2.zero())
    with pytest.raises(kzg.KZGError) as e:
        kzg.commit(setup, [1, 3, 2, 4])
    assert "PolynomialTooLarge(4, 2)" in str(e.value)
    assert kzg._strip([1, 2, 0, 0]) == [1, 2] and kzg._strip([0, 0]) == []


def test_bench_helpers_sharded_trapdoor_traffic_and_scalars():
    """bench.py's host logic: the combined-commitment check of the N > 1 step (each rank Horner-sums its slice, rank 0
    weighs the parts with tau^(rank n)), the DRAM-traffic reader of the committed ncu export, uniform-below-r scalars"""
    import bench
    nrng = np.random.default_rng(5)
    tau = 0x1234567890ABCDEF % FR_MODULUS
    world, n = 3, 40
    limbs = bench.rand_fr_limbs(nrng, world * n)
    vals = fr_list(limbs)
    assert all(0 <= v < FR_MODULUS for v in vals) and limbs.shape == (world * n, 8)
    whole = sum(v * pow(tau, i, FR_MODULUS) for i, v in enumerate(vals)) % FR_MODULUS
    parts = [bench.horner_mod_r(limbs[k * n:(k + 1) * n], tau, FR_MODULUS) for k in range(world)]
    assert sum(p * pow(tau, k * n, FR_MODULUS) for k, p in enumerate(parts)) % FR_MODULUS == whole
    assert bench.horner_mod_r(limbs, tau, FR_MODULUS) == whole
    # top limbs reach above 2^252 (round 1 drew below 2^252, which hides the real top-window distribution of the MSM)
    big = bench.rand_fr_limbs(nrng, 4000)
    assert (big[:, 7] >= 0x10000000).mean() > 0.5 and (big[:, 7] <= 0x30644e72).all()
    traffic, src = bench.ncu_dram_traffic("ncu_msm_accumulate_r02_raw.csv", "msm_accumulate_kernel")
    assert src == "profiles/ncu_msm_accumulate_r02_raw.csv" and 1.0e9 < traffic < 3.0e9
    assert bench.ncu_dram_traffic("missing.csv", "x") == (None, None)
