"""Value types of the host-side mirror: field elements and group points in the C ABI's layout.

Scalars (`E::ScalarField`) are Python ints mod r on the host side; they cross the ABI as Montgomery
limbs (x * 2^256 mod r, 8 x u32 little-endian) — the same bytes arkworks keeps in RAM.  Points are
affine Montgomery coordinates plus an infinity flag."""
from __future__ import annotations

import hashlib
import os

import numpy as np

Z = 4965661367192848881
FQ_MODULUS = 36 * Z**4 + 36 * Z**3 + 24 * Z**2 + 6 * Z + 1
FR_MODULUS = 36 * Z**4 + 36 * Z**3 + 18 * Z**2 + 6 * Z + 1
_R256 = 1 << 256
_FR_RINV = pow(_R256, -1, FR_MODULUS)
_FQ_RINV = pow(_R256, -1, FQ_MODULUS)
FR_GENERATOR = 5
FR_TWO_ADICITY = 28


def _limbs(x: int) -> np.ndarray:
    return np.frombuffer(int(x).to_bytes(32, "little"), dtype=np.uint32).copy()


def _int(a) -> int:
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint32).tobytes(), "little")


def fr_to_limbs(x: int) -> np.ndarray:
    return _limbs((x % FR_MODULUS) * _R256 % FR_MODULUS)


def fr_from_limbs(a) -> int:
    return _int(a) * _FR_RINV % FR_MODULUS


def fr_array(xs) -> np.ndarray:
    """list of ints -> (n, 8) uint32 Montgomery limbs"""
    if len(xs) == 0:
        return np.zeros((0, 8), np.uint32)
    buf = b"".join(((x % FR_MODULUS) * _R256 % FR_MODULUS).to_bytes(32, "little") for x in xs)
    return np.frombuffer(buf, dtype=np.uint32).reshape(len(xs), 8).copy()


def fr_list(a) -> list:
    a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, 8)
    raw = a.tobytes()
    return [int.from_bytes(raw[32 * i: 32 * i + 32], "little") * _FR_RINV % FR_MODULUS for i in range(a.shape[0])]


def fq_to_limbs(x: int) -> np.ndarray:
    return _limbs((x % FQ_MODULUS) * _R256 % FQ_MODULUS)


def fq_from_limbs(a) -> int:
    return _int(a) * _FQ_RINV % FQ_MODULUS


class G1:
    """E::G1 — affine Montgomery x||y (16 limbs) + infinity flag.  Equality is point equality."""
    __slots__ = ("xy", "inf")

    def __init__(self, xy=None, inf=False):
        self.inf = bool(inf) or xy is None
        self.xy = np.zeros(16, np.uint32) if self.inf else np.ascontiguousarray(xy, dtype=np.uint32).reshape(16).copy()

    @classmethod
    def zero(cls):
        return cls(None, True)

    @classmethod
    def generator(cls):
        return cls(np.concatenate([fq_to_limbs(1), fq_to_limbs(2)]))

    def __eq__(self, other):
        return isinstance(other, G1) and self.inf == other.inf and (self.inf or np.array_equal(self.xy, other.xy))

    def __hash__(self):
        return hash((self.inf, self.xy.tobytes()))

    def to_affine_ints(self):
        """canonical (x, y) integers or None — for serialisation / comparison with other libraries"""
        return None if self.inf else (fq_from_limbs(self.xy[:8]), fq_from_limbs(self.xy[8:]))

    def __repr__(self):
        return "G1(inf)" if self.inf else "G1(x=%#x…)" % (fq_from_limbs(self.xy[:8]) >> 200)


class G2:
    """E::G2 — affine Montgomery x.c0||x.c1||y.c0||y.c1 (32 limbs) + infinity flag."""
    __slots__ = ("xy", "inf")

    def __init__(self, xy=None, inf=False):
        self.inf = bool(inf) or xy is None
        self.xy = np.zeros(32, np.uint32) if self.inf else np.ascontiguousarray(xy, dtype=np.uint32).reshape(32).copy()

    @classmethod
    def zero(cls):
        return cls(None, True)

    def __eq__(self, other):
        return isinstance(other, G2) and self.inf == other.inf and (self.inf or np.array_equal(self.xy, other.xy))

    def __hash__(self):
        return hash((self.inf, self.xy.tobytes()))

    def to_affine_ints(self):
        if self.inf:
            return None
        c = [fq_from_limbs(self.xy[8 * i: 8 * i + 8]) for i in range(4)]
        return ((c[0], c[1]), (c[2], c[3]))


def pack_g1(points):
    """list[G1] -> ((n,16) uint32, (n,) uint8)"""
    n = len(points)
    xy = np.zeros((n, 16), np.uint32)
    inf = np.zeros(n, np.uint8)
    for i, p in enumerate(points):
        xy[i] = p.xy
        inf[i] = 1 if p.inf else 0
    return xy, inf


def unpack_g1(xy, inf):
    return [G1(xy[i], bool(inf[i])) for i in range(len(inf))]


def pack_g2(points):
    n = len(points)
    xy = np.zeros((n, 32), np.uint32)
    inf = np.zeros(n, np.uint8)
    for i, p in enumerate(points):
        xy[i] = p.xy
        inf[i] = 1 if p.inf else 0
    return xy, inf


class SecureFrRng:
    """The `rng: &mut impl rand::Rng` argument of the reference backed by the operating system's CSPRNG (`os.urandom`):
    `fr()` plays `E::ScalarField::rand(rng)` (src/kem.rs:26, src/vec.rs:32) — 254 uniformly random bits, rejected if
    not below r.  This is what `encapsulate` / `encrypt` / `vec_encrypt` / `vec_commit` use when called with rng=None:
    the encapsulation randomness r and the hiding pad of a vector commitment MUST be unpredictable (a repeated or
    guessable r makes the shared secret A^r gT^(-v r) recoverable; a guessable pad makes the commitment non-hiding)."""

    def fr(self) -> int:
        while True:
            v = int.from_bytes(os.urandom(32), "little") & ((1 << 254) - 1)
            if v < FR_MODULUS:
                return v

    def bytes(self, n: int) -> bytes:
        return os.urandom(n)


class SeededFrRng:
    """DETERMINISTIC stand-in for the reference's rng argument — tests, benchmarks and reproducible fixtures only; never
    for real encryptions or commitments (see SecureFrRng).  The seed is mandatory.  `fr()` = one `Fr::rand` draw per
    call, in call order, from a blake2b counter stream.  (arkworks' exact StdRng stream cannot be reproduced without the
    Rust crates; the C ABI takes the drawn scalars as input, so a Rust shim keeps using arkworks' own `Fr::rand`.)"""

    def __init__(self, seed: int):
        self._seed = int(seed).to_bytes(16, "little", signed=False)
        self._ctr = 0

    def _block(self) -> bytes:
        h = hashlib.blake2b(self._seed + self._ctr.to_bytes(8, "little"), digest_size=32).digest()
        self._ctr += 1
        return h

    def fr(self) -> int:
        while True:  # rejection sampling of 254 bits, like UniformRand for Fp
            v = int.from_bytes(self._block(), "little") & ((1 << 254) - 1)
            if v < FR_MODULUS:
                return v

    def bytes(self, n: int) -> bytes:
        out = b""
        while len(out) < n:
            out += self._block()
        return out[:n]


FrRng = SeededFrRng   # historical name; the seed argument is required


def rng_or_secure(rng):
    """rng=None -> the OS CSPRNG (the reference takes a caller CSPRNG; a deterministic default would be a vulnerability)"""
    return SecureFrRng() if rng is None else rng


class Radix2EvaluationDomain:
    """ark-poly `Radix2EvaluationDomain::<Fr>::new(n)`: size = next power of two >= n, generator
    5^((r-1)/size) (src/vec.rs:36, src/kzg.rs:163, tests/laconic_ot.rs:81-85).  Host-side index math
    only; transforms run on the GPU (Context.fr_ntt)."""

    def __init__(self, n: int):
        size = 1
        while size < n:
            size <<= 1
        if size.bit_length() - 1 > FR_TWO_ADICITY:
            raise ValueError("Radix2EvaluationDomain::new returned None (size > 2^28)")
        self.size = size
        self.group_gen = pow(FR_GENERATOR, (FR_MODULUS - 1) // size, FR_MODULUS)

    def elements(self):
        out, x = [], 1
        for _ in range(self.size):
            out.append(x)
            x = x * self.group_gen % FR_MODULUS
        return out

    def elements_limbs(self) -> np.ndarray:
        return fr_array(self.elements())
