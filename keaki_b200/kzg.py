"""KZG polynomial commitments — host-side mirror of src/kzg.rs with E = Bn254.

Same names, argument order and error behaviour as the reference; the arithmetic behind each call
is one C-ABI entry point (include/keaki_b200.h) running on the GPU."""
from __future__ import annotations

import weakref

import numpy as np

from . import ptau as _ptau
from ._ffi import Context, InvalidSrsPoint, PolynomialTooLarge
from .types import G1, G2, FR_MODULUS, Radix2EvaluationDomain, fr_array, fr_to_limbs, unpack_g1

KZGError = PolynomialTooLarge  # the only variant (src/kzg.rs:205-209)

_latest_ctx = None   # weak reference to the context of the most recently created KZGSetup


def default_context() -> Context:
    """Context for the reference calls that take no setup (`decapsulate`, `decrypt`, `vec_decrypt`): the one of the most
    recently created KZGSetup (same device, no second set of tables)."""
    ctx = _latest_ctx() if _latest_ctx is not None else None
    if ctx is None or getattr(ctx, "h", None) is None:
        raise RuntimeError("keaki_b200: no live KZGSetup - create one (KZGSetup.setup / new_from_file) or pass ctx= explicitly")
    return ctx


def _strip(p):
    """DensePolynomial::from_coefficients_* strips trailing zero coefficients."""
    p = [int(c) % FR_MODULUS for c in p]
    while p and p[-1] == 0:
        p.pop()
    return p


class KZGSetup:
    """src/kzg.rs:22-85.  Owns the GPU context; the SRS is uploaded once and stays resident."""

    def __init__(self, ctx: Context, g1_xy: np.ndarray, tau_g2: G2):
        global _latest_ctx
        self.ctx = ctx
        self._g1_xy = g1_xy          # (n, 16) Montgomery affine
        self._tau_g2 = tau_g2
        if ctx is not None:
            _latest_ctx = weakref.ref(ctx)

    @classmethod
    def new_from_file(cls, file: str, device: int = 0, ctx: Context | None = None, validate: bool = True) -> "KZGSetup":
        """src/kzg.rs:33-52.  Unlike the reference (`deserialize_uncompressed_unchecked`, src/kzg/ptau.rs:266,314) the
        uploaded points are validated on the GPU (kb_srs_validate): a file whose elements are not points of the curve
        is a SetupFileError("ParseError(...)") instead of a silently wrong SRS."""
        g1, g2 = _ptau.get_powers_from_file(file)
        if g2.shape[0] < 2:
            raise _ptau.SetupFileError("EmptySection(3)")
        ctx = ctx or Context(device)
        ctx.srs_upload(g1, g2[1])
        if validate:
            try:
                ctx.srs_validate()
            except InvalidSrsPoint as e:
                raise _ptau.SetupFileError(f"ParseError({e})") from e
        return cls(ctx, g1, G2(g2[1]))

    @classmethod
    def setup(cls, secret: int, max_d: int, device: int = 0, ctx: Context | None = None) -> "KZGSetup":
        """src/kzg.rs:55-70 ("Don't use this"): powers generated on the device."""
        ctx = ctx or Context(device)
        g1, t2 = ctx.srs_generate(fr_to_limbs(secret), max_d)
        return cls(ctx, g1, G2(t2))

    def g1_pow(self):
        return unpack_g1(self._g1_xy, np.zeros(len(self._g1_xy), np.uint8))

    def g1_aff(self):
        return self.g1_pow()

    def tau_g2(self) -> G2:
        return self._tau_g2

    def __len__(self):
        return self._g1_xy.shape[0]


def commit(setup: KZGSetup, p) -> G1:
    """src/kzg.rs:89-101"""
    p = _strip(p)
    if len(p) > len(setup):
        raise PolynomialTooLarge(-3, f"PolynomialTooLarge({len(p)}, {len(setup)})")
    xy, inf = setup.ctx.msm_g1(fr_array(p), n=len(p))
    return G1(xy, inf)


def open(setup: KZGSetup, p, point: int) -> G1:  # noqa: A001 - reference name
    """src/kzg.rs:104-124"""
    p = _strip(p)
    proofs, inf = setup.ctx.open_batch(fr_array(p), fr_array([point]))
    return G1(proofs[0], inf[0])


def open_many(setup: KZGSetup, p, points):
    """m independent `open` calls of one polynomial in one launch sequence (kb_open_batch)."""
    p = _strip(p)
    proofs, inf = setup.ctx.open_batch(fr_array(p), fr_array(list(points)))
    return unpack_g1(proofs, inf)


def verify(setup: KZGSetup, commitment: G1, point: int, value: int, proof: G1) -> bool:
    """src/kzg.rs:127-151"""
    ok = setup.ctx.verify_batch(commitment.xy.reshape(1, 16), np.array([commitment.inf], np.uint8), fr_array([point]),
                                fr_array([value]), proof.xy.reshape(1, 16), np.array([proof.inf], np.uint8))
    return bool(ok[0])


def open_fk(setup: KZGSetup, p, domain_d: Radix2EvaluationDomain):
    """src/kzg.rs:157-203.  `p` is a coefficient slice (not stripped), len(p) = d = domain size.
    Like the reference this raises when d exceeds the SRS (slice panic at :169) or 2d > 2^28 (:163)."""
    d = len(p)
    if d != domain_d.size:
        raise ValueError("open_fk: len(p) must equal the domain size")
    proofs, inf = setup.ctx.open_all_fk(fr_array([int(c) for c in p]))
    return unpack_g1(proofs, inf)
