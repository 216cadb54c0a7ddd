"""keaki_b200 — B200-native hot path of brech1/keaki behind the keaki API.

Modules mirror the reference crate root (src/lib.rs:5-8): `enc`, `kem`, `kzg`, `vec`, plus
`laconic_ot` (tests/laconic_ot.rs) and `ptau`.  All arithmetic runs in libkeaki_b200.so on the GPU;
there is no CPU fallback."""
from . import _ffi  # noqa: F401
from ._ffi import Context, InvalidSrsPoint, KeakiB200Error, PolynomialTooLarge  # noqa: F401
from .types import G1, G2, FrRng, SecureFrRng, SeededFrRng, Radix2EvaluationDomain, FR_MODULUS, FQ_MODULUS  # noqa: F401

__all__ = ["Context", "KeakiB200Error", "PolynomialTooLarge", "InvalidSrsPoint", "G1", "G2", "SecureFrRng", "SeededFrRng", "Radix2EvaluationDomain",
           "enc", "kem", "kzg", "vec", "laconic_ot", "ptau"]
