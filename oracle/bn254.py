"""BN254 big-int oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain Python integers) of the arithmetic keaki delegates to arkworks 0.4.x on its
hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

The reference (/root/reference) holds none of this arithmetic in-tree: it lives in the third-party
crates pinned by Cargo.lock — ark-bn254 0.4.0 (Cargo.lock:30-31), ark-ec 0.4.2 (:41-42), ark-ff 0.4.2
(:58-59), ark-poly 0.4.2 (:101-102), ark-serialize 0.4.2 (:114-115), blake3 1.5.4 (:165-166) — none of
which is vendored and none of which can be built here (no cargo/rustc, no network).  This module
restates their *published conventions* (curve constants, tower, optimal-ate pairing, arkworks' BN
final-exponentiation exponent, canonical little-endian serialisation); results that are canonical
bytes (affine coordinates, GT bytes, keys) are algorithm-independent for on-curve inputs.

PARITY PINNING: the reference contains no byte-level golden vector for this path.
What pins this oracle: (i) the reference's own ptau fixture decodes (after de-Montgomerising) to
the generators defined here (tests/test_oracle.py); (ii) the reference's algebraic unit tests,
re-instantiated on BN254 (tests/test_oracle_reference_suite.py); (iii) BLAKE3 KAT from the blake3
wheel.  Byte-level parity against arkworks itself is therefore "parity unpinned".

Call sites restated: src/kzg.rs:57,60,98,135,144,148,190; src/kem.rs:22,30,32,36,37,58,61.
"""
from __future__ import annotations

# ----------------------------------------------------------------------------------------------
# Parameters (ark-bn254 0.4.0: fields/fq.rs, fr.rs, curves/g1.rs, g2.rs, mod.rs)
# ----------------------------------------------------------------------------------------------
Z = 4965661367192848881  # BN parameter x (positive for BN254)
Q = 36 * Z**4 + 36 * Z**3 + 24 * Z**2 + 6 * Z + 1
R = 36 * Z**4 + 36 * Z**3 + 18 * Z**2 + 6 * Z + 1
assert Q == 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
assert R == 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001

MONT_R = 1 << 256  # Montgomery radix used by arkworks (4 x u64 limbs)
FR_GENERATOR = 5  # multiplicative generator of Fr (ark-bn254 FrConfig::GENERATOR)
FR_TWO_ADICITY = 28

ATE_LOOP_COUNT_ABS = 6 * Z + 2
# arkworks' signed-digit table (LSB first) — only its value matters for the result.
ATE_LOOP_COUNT = [0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0,
                  1, 1, 1, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0,
                  0, 1, 0, 1, 1]
assert sum(d << i for i, d in enumerate(ATE_LOOP_COUNT)) == ATE_LOOP_COUNT_ABS

B1 = 3  # G1: y^2 = x^3 + 3


def inv_mod(a: int, m: int) -> int:
    return pow(a, -1, m)


# ----------------------------------------------------------------------------------------------
# Fq2 = Fq[u]/(u^2+1); elements are tuples (c0, c1)
# ----------------------------------------------------------------------------------------------
F2_ZERO = (0, 0)
F2_ONE = (1, 0)
XI = (9, 1)  # non-residue for Fq6: 9 + u


def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_neg(a): return ((-a[0]) % Q, (-a[1]) % Q)
def f2_conj(a): return (a[0], (-a[1]) % Q)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_sqr(a): return ((a[0] + a[1]) * (a[0] - a[1]) % Q, 2 * a[0] * a[1] % Q)
def f2_scale(a, k): return (a[0] * k % Q, a[1] * k % Q)
def f2_mul_xi(a): return ((9 * a[0] - a[1]) % Q, (9 * a[1] + a[0]) % Q)


def f2_inv(a):
    d = inv_mod((a[0] * a[0] + a[1] * a[1]) % Q, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)


def f2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


B2 = f2_mul((3, 0), f2_inv(XI))  # G2 (D-type twist): y^2 = x^3 + 3/(9+u)

# ----------------------------------------------------------------------------------------------
# Fq6 = Fq2[v]/(v^3 - xi); elements are tuples (c0, c1, c2) of Fq2
# ----------------------------------------------------------------------------------------------
F6_ZERO = (F2_ZERO, F2_ZERO, F2_ZERO)
F6_ONE = (F2_ONE, F2_ZERO, F2_ZERO)


def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))
def f6_neg(a): return tuple(f2_neg(x) for x in a)


def f6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    c0 = f2_add(f2_mul(a0, b0), f2_mul_xi(f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
    c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul_xi(f2_mul(a2, b2)))
    c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a1, b1)), f2_mul(a2, b0))
    return (c0, c1, c2)


def f6_mul_v(a):  # multiply by v
    return (f2_mul_xi(a[2]), a[0], a[1])


def f6_inv(a):
    a0, a1, a2 = a
    t0 = f2_sub(f2_sqr(a0), f2_mul_xi(f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul_xi(f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    d = f2_add(f2_mul(a0, t0), f2_mul_xi(f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
    di = f2_inv(d)
    return (f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di))


# ----------------------------------------------------------------------------------------------
# Fq12 = Fq6[w]/(w^2 - v); elements are tuples (c0, c1) of Fq6
# ----------------------------------------------------------------------------------------------
F12_ONE = (F6_ONE, F6_ZERO)


def f12_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    t0 = f6_mul(a0, b0)
    t1 = f6_mul(a1, b1)
    c0 = f6_add(t0, f6_mul_v(t1))
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a0, a1), f6_add(b0, b1)), t0), t1)
    return (c0, c1)


def f12_sqr(a): return f12_mul(a, a)
def f12_conj(a): return (a[0], f6_neg(a[1]))


def f12_inv(a):
    a0, a1 = a
    d = f6_inv(f6_sub(f6_mul(a0, a0), f6_mul_v(f6_mul(a1, a1))))
    return (f6_mul(a0, d), f6_neg(f6_mul(a1, d)))


def f12_pow(a, e):
    if e < 0:
        return f12_pow(f12_inv(a), -e)
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_sqr(a)
        e >>= 1
    return r


# Frobenius: coefficient i of w^i (i = 0..5, Fq2 coefficients) is conjugated and scaled by
# gamma_i^(k) = xi^(i (q^k - 1)/6).  Basis mapping: Fq12 element = sum_{j,i} c[j][i] v^i w^j with
# v = w^2, i.e. coefficient of w^(2i + j).
def _f12_to_w(a):
    out = [None] * 6
    for j in range(2):
        for i in range(3):
            out[2 * i + j] = a[j][i]
    return out


def _f12_from_w(c):
    return ((c[0], c[2], c[4]), (c[1], c[3], c[5]))


_FROB_GAMMA = {}
for _k in (1, 2, 3):
    _FROB_GAMMA[_k] = [f2_pow(XI, i * (Q**_k - 1) // 6) for i in range(6)]


def f12_frobenius(a, k):
    """a^(q^k) for k in 1..3."""
    c = _f12_to_w(a)
    out = []
    for i in range(6):
        x = c[i]
        if k & 1:
            x = f2_conj(x)
        out.append(f2_mul(x, _FROB_GAMMA[k][i]))
    return _f12_from_w(out)


# ----------------------------------------------------------------------------------------------
# Curves.  Affine points are (x, y) tuples or None for infinity.  G1 over Fq, G2 over Fq2.
# ----------------------------------------------------------------------------------------------
G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


def g1_on_curve(p):
    return p is None or (p[1] * p[1] - p[0] * p[0] * p[0] - B1) % Q == 0


def g2_on_curve(p):
    if p is None:
        return True
    x, y = p
    return f2_sub(f2_sqr(y), f2_add(f2_mul(f2_sqr(x), x), B2)) == F2_ZERO


def g1_neg(p): return None if p is None else (p[0], (-p[1]) % Q)


def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % Q == 0:
            return None
        lam = 3 * p[0] * p[0] * inv_mod(2 * p[1], Q) % Q
    else:
        lam = (q[1] - p[1]) * inv_mod(q[0] - p[0], Q) % Q
    x = (lam * lam - p[0] - q[0]) % Q
    return (x, (lam * (p[0] - x) - p[1]) % Q)


def g1_mul(p, k):
    """Scalar multiplication with Jacobian doubling internally (fast enough for 2^12 points)."""
    k %= R
    if p is None or k == 0:
        return None
    # Jacobian double-and-add, mixed additions
    X, Y, Zc = 0, 1, 0
    px, py = p
    for bit in bin(k)[2:]:
        if Zc:
            # dbl-2009-l (a = 0)
            A = X * X % Q; Bv = Y * Y % Q; C = Bv * Bv % Q
            D = 2 * ((X + Bv) * (X + Bv) - A - C) % Q
            E = 3 * A % Q
            X3 = (E * E - 2 * D) % Q
            Y3 = (E * (D - X3) - 8 * C) % Q
            Zc = 2 * Y * Zc % Q
            X, Y = X3, Y3
        if bit == '1':
            if not Zc:
                X, Y, Zc = px, py, 1
            else:
                Z2 = Zc * Zc % Q
                U2 = px * Z2 % Q
                S2 = py * Z2 * Zc % Q
                H = (U2 - X) % Q
                Rr = (S2 - Y) % Q
                if H == 0:
                    if Rr == 0:
                        aff = g1_add(p, p)  # doubling (never hit for prime-order k < R)
                        X, Y, Zc = aff[0], aff[1], 1
                        continue
                    X, Y, Zc = 0, 1, 0
                    continue
                H2 = H * H % Q; H3 = H * H2 % Q; V = X * H2 % Q
                X3 = (Rr * Rr - H3 - 2 * V) % Q
                Y3 = (Rr * (V - X3) - Y * H3) % Q
                Zc = Zc * H % Q
                X, Y = X3, Y3
    if not Zc:
        return None
    zi = inv_mod(Zc, Q)
    zi2 = zi * zi % Q
    return (X * zi2 % Q, Y * zi2 * zi % Q)


def g1_msm(points, scalars):
    """Naive sum of scalar multiples (the definition `commit` must equal, src/kzg.rs:241-258)."""
    acc = None
    for p, s in zip(points, scalars):
        acc = g1_add(acc, g1_mul(p, s))
    return acc


def g2_neg(p): return None if p is None else (p[0], f2_neg(p[1]))


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if f2_add(p[1], q[1]) == F2_ZERO:
            return None
        lam = f2_mul(f2_scale(f2_sqr(p[0]), 3), f2_inv(f2_scale(p[1], 2)))
    else:
        lam = f2_mul(f2_sub(q[1], p[1]), f2_inv(f2_sub(q[0], p[0])))
    x = f2_sub(f2_sub(f2_sqr(lam), p[0]), q[0])
    return (x, f2_sub(f2_mul(lam, f2_sub(p[0], x)), p[1]))


def g2_mul(p, k):
    k %= R
    acc = None
    for bit in bin(k)[2:] if k else '':
        acc = g2_add(acc, acc)
        if bit == '1':
            acc = g2_add(acc, p)
    return acc


# ----------------------------------------------------------------------------------------------
# Optimal-ate pairing (ark-ec 0.4.2 models/bn/mod.rs: multi_miller_loop + final_exponentiation)
# ----------------------------------------------------------------------------------------------
# q-power Frobenius on the twist: (x, y) -> (conj(x) * xi^((q-1)/3), conj(y) * xi^((q-1)/2))
_TW_X = f2_pow(XI, (Q - 1) // 3)
_TW_Y = f2_pow(XI, (Q - 1) // 2)


def g2_frobenius(p):
    return (f2_mul(f2_conj(p[0]), _TW_X), f2_mul(f2_conj(p[1]), _TW_Y))


def _line(t, qpt, p):
    """Line through t and qpt (tangent if equal) on the twist, evaluated at P in G1, as a sparse
    Fq12 element; returns (line, t + qpt).  Untwist psi(x', y') = (x' w^2, y' w^3), so the line
    y - yT - lam (x - xT) at psi-images, multiplied through by subfield-safe factors, is
       yP  -  lam xP w  +  (lam xT - yT) w^3 .
    Any Fq2 scaling of the line dies in the final exponentiation."""
    if t[0] == qpt[0] and t[1] == qpt[1]:
        lam = f2_mul(f2_scale(f2_sqr(t[0]), 3), f2_inv(f2_scale(t[1], 2)))
    else:
        lam = f2_mul(f2_sub(qpt[1], t[1]), f2_inv(f2_sub(qpt[0], t[0])))
    x3 = f2_sub(f2_sub(f2_sqr(lam), t[0]), qpt[0])
    y3 = f2_sub(f2_mul(lam, f2_sub(t[0], x3)), t[1])
    c = [F2_ZERO] * 6
    c[0] = (p[1] % Q, 0)
    c[1] = f2_neg(f2_scale(lam, p[0]))
    c[3] = f2_sub(f2_mul(lam, t[0]), t[1])
    return _f12_from_w(c), (x3, y3)


def miller_loop(p, qpt):
    """f_{6z+2,Q}(P) * l_{[6z+2]Q, pi(Q)}(P) * l_{., -pi^2(Q)}(P); 1 if either input is infinity
    (arkworks skips such pairs)."""
    if p is None or qpt is None:
        return F12_ONE
    f = F12_ONE
    t = qpt
    nq = g2_neg(qpt)
    for i in range(len(ATE_LOOP_COUNT) - 2, -1, -1):
        f = f12_sqr(f)
        ln, t = _line(t, t, p)
        f = f12_mul(f, ln)
        d = ATE_LOOP_COUNT[i]
        if d == 1:
            ln, t = _line(t, qpt, p)
            f = f12_mul(f, ln)
        elif d == -1:
            ln, t = _line(t, nq, p)
            f = f12_mul(f, ln)
    q1 = g2_frobenius(qpt)
    q2 = g2_neg(g2_frobenius(q1))
    ln, t = _line(t, q1, p)
    f = f12_mul(f, ln)
    ln, t = _line(t, q2, p)
    f = f12_mul(f, ln)
    return f


# arkworks' BN hard-part exponent (Fuentes-Castaneda): NOT (q^4-q^2+1)/r but a multiple of it.
HARD_EXP = (Q**3 * (12 * Z**3 + 6 * Z**2 + 4 * Z - 1) + Q**2 * (12 * Z**3 + 6 * Z**2 + 6 * Z)
            + Q * (12 * Z**3 + 6 * Z**2 + 4 * Z) + (12 * Z**3 + 12 * Z**2 + 6 * Z + 1))
assert (Q**4 - Q**2 + 1) % R == 0
assert HARD_EXP == 2 * Z * (6 * Z**2 + 3 * Z + 1) * ((Q**4 - Q**2 + 1) // R)
FINAL_EXP = (Q**6 - 1) * (Q**2 + 1) * HARD_EXP


def final_exponentiation_naive(f):
    """f^FINAL_EXP by square-and-multiply: the definition."""
    return f12_pow(f, FINAL_EXP)


def final_exponentiation(f):
    """Same value, restating the arkworks chain (ark-ec 0.4.2 models/bn/mod.rs
    final_exponentiation; y0..y16).  tests/test_oracle.py checks it equals the naive power."""
    f1 = f12_conj(f)
    f2 = f12_inv(f)
    r = f12_mul(f1, f2)
    f2 = r
    r = f12_frobenius(r, 2)
    r = f12_mul(r, f2)

    def exp_by_neg_x(a):
        return f12_conj(f12_pow(a, Z))

    y0 = exp_by_neg_x(r)
    y1 = f12_sqr(y0)
    y2 = f12_sqr(y1)
    y3 = f12_mul(y2, y1)
    y4 = exp_by_neg_x(y3)
    y5 = f12_sqr(y4)
    y6 = exp_by_neg_x(y5)
    y3 = f12_conj(y3)
    y6 = f12_conj(y6)
    y7 = f12_mul(y6, y4)
    y8 = f12_mul(y7, y3)
    y9 = f12_mul(y8, y1)
    y10 = f12_mul(y8, y4)
    y11 = f12_mul(y10, r)
    y12 = f12_frobenius(y9, 1)
    y13 = f12_mul(y12, y11)
    y8 = f12_frobenius(y8, 2)
    y14 = f12_mul(y8, y13)
    r = f12_conj(r)
    y15 = f12_frobenius(f12_mul(r, y9), 3)
    return f12_mul(y15, y14)


def pairing(p, qpt):
    """E::pairing(P, Q) as arkworks computes it on BN254 (src/kem.rs:30,58; src/kzg.rs:148)."""
    return final_exponentiation(miller_loop(p, qpt))


# ----------------------------------------------------------------------------------------------
# ark-serialize 0.4.2 canonical bytes
# ----------------------------------------------------------------------------------------------
def fq_to_bytes(x: int) -> bytes:
    return int(x % Q).to_bytes(32, "little")


def gt_to_bytes(a) -> bytes:
    """serialize_uncompressed of Fq12: c0.c0.c0, c0.c0.c1, c0.c1.c0, ... c1.c2.c1 — 384 B
    (src/kem.rs:31-32,60-61)."""
    out = bytearray()
    for j in range(2):
        for i in range(3):
            out += fq_to_bytes(a[j][i][0]) + fq_to_bytes(a[j][i][1])
    return bytes(out)


def g1_to_bytes(p) -> bytes:
    """x || y little-endian canonical, 64 B; infinity = zeros (flags are carried separately at the C ABI)."""
    if p is None:
        return bytes(64)
    return fq_to_bytes(p[0]) + fq_to_bytes(p[1])


def g2_to_bytes(p) -> bytes:
    if p is None:
        return bytes(128)
    return fq_to_bytes(p[0][0]) + fq_to_bytes(p[0][1]) + fq_to_bytes(p[1][0]) + fq_to_bytes(p[1][1])


# ----------------------------------------------------------------------------------------------
# ark-serialize 0.4.2 point encodings (`CanonicalSerialize for Affine<P>`, ark-ec 0.4.2
# models/short_weierstrass/{mod.rs, serialization_flags.rs}) — the wire format a `Ciphertext<E>` (src/enc.rs:13)
# would travel in.  [restated from memory of the crate; no byte-level vector exists in the
# reference, so this layer is "parity unpinned" like the rest]
#   compressed   = x                    with SWFlags in the two top bits of the LAST byte
#   uncompressed = x || y               with the same flags in the last byte of y
#   SWFlags: YIsPositive = 0 (y <= -y), PointAtInfinity = 1 << 6, YIsNegative = 1 << 7 (y > -y); infinity is x = y = 0
#   Fq2 is c0 || c1 and orders by (c1, c0) (`Ord for QuadExtField`); Fq orders by its canonical integer.
# ----------------------------------------------------------------------------------------------
SW_FLAG_INFINITY = 0x40
SW_FLAG_Y_NEGATIVE = 0x80


def _fq_is_larger(y):
    return y > (-y) % Q


def _f2_is_larger(y):
    n = f2_neg(y)
    return (y[1], y[0]) > (n[1], n[0])


def g1_serialize(p, compress: bool) -> bytes:
    if p is None:
        out = bytearray(32 if compress else 64)
        out[-1] |= SW_FLAG_INFINITY
        return bytes(out)
    out = bytearray(fq_to_bytes(p[0]) + (b"" if compress else fq_to_bytes(p[1])))
    if _fq_is_larger(p[1]):
        out[-1] |= SW_FLAG_Y_NEGATIVE
    return bytes(out)


def g2_serialize(p, compress: bool) -> bytes:
    if p is None:
        out = bytearray(64 if compress else 128)
        out[-1] |= SW_FLAG_INFINITY
        return bytes(out)
    out = bytearray(fq_to_bytes(p[0][0]) + fq_to_bytes(p[0][1]) + (b"" if compress else fq_to_bytes(p[1][0]) + fq_to_bytes(p[1][1])))
    if _f2_is_larger(p[1]):
        out[-1] |= SW_FLAG_Y_NEGATIVE
    return bytes(out)


def fq_sqrt(a):
    """q = 3 mod 4: a^((q+1)/4) when a is a square, else None."""
    a %= Q
    s = pow(a, (Q + 1) // 4, Q)
    return s if s * s % Q == a else None


def f2_sqrt(a):
    """a square root in Fq2 = Fq[u]/(u^2+1) by the norm method, or None."""
    a0, a1 = a[0] % Q, a[1] % Q
    if a1 == 0:
        s = fq_sqrt(a0)
        if s is not None:
            return (s, 0)
        return (0, fq_sqrt((-a0) % Q))          # -1 is a non-residue, so -a0 is a square
    alpha = fq_sqrt((a0 * a0 + a1 * a1) % Q)
    if alpha is None:
        return None
    half = inv_mod(2, Q)
    c0 = fq_sqrt((a0 + alpha) * half % Q)
    if c0 is None:
        c0 = fq_sqrt((a0 - alpha) * half % Q)
    if c0 is None or c0 == 0:
        return None
    c1 = a1 * inv_mod(2 * c0 % Q, Q) % Q
    r = (c0, c1)
    return r if f2_sqr(r) == (a0, a1) else None


def _read_fq(b):
    x = int.from_bytes(b, "little")
    if x >= Q:
        raise ValueError("InvalidData: field element not below the modulus")
    return x


def _split_flags(b):
    flags = b[-1] & 0xC0
    if flags == 0xC0:
        raise ValueError("InvalidData: both flag bits set")
    return bytes(b[:-1]) + bytes([b[-1] & 0x3F]), flags


def g1_deserialize(b: bytes, compress: bool, validate: bool = True):
    """`deserialize_{compressed,uncompressed}` (validate) / `_unchecked`; raises ValueError like arkworks errors."""
    if len(b) != (32 if compress else 64):
        raise ValueError("InvalidData: length")
    body, flags = _split_flags(b)
    x = _read_fq(body[:32])
    if flags & SW_FLAG_INFINITY:
        return None
    if compress:
        y = fq_sqrt((x * x * x + 3) % Q)
        if y is None:
            raise ValueError("InvalidData: x is not on the curve")
        if _fq_is_larger(y) != bool(flags & SW_FLAG_Y_NEGATIVE):
            y = (-y) % Q
    else:
        y = _read_fq(body[32:64])
    p = (x, y)
    if validate and not g1_on_curve(p):
        raise ValueError("InvalidData: point not on the curve")
    return p


def g2_in_subgroup(p):
    """[r]P == O without reducing the scalar (g2_mul reduces mod r, which is only valid inside the subgroup)."""
    acc = None
    for bit in bin(R)[2:]:
        acc = g2_add(acc, acc)
        if bit == '1':
            acc = g2_add(acc, p)
    return acc is None


def g2_deserialize(b: bytes, compress: bool, validate: bool = True):
    if len(b) != (64 if compress else 128):
        raise ValueError("InvalidData: length")
    body, flags = _split_flags(b)
    x = (_read_fq(body[:32]), _read_fq(body[32:64]))
    if flags & SW_FLAG_INFINITY:
        return None
    if compress:
        y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), B2))
        if y is None:
            raise ValueError("InvalidData: x is not on the curve")
        if _f2_is_larger(y) != bool(flags & SW_FLAG_Y_NEGATIVE):
            y = f2_neg(y)
    else:
        y = (_read_fq(body[64:96]), _read_fq(body[96:128]))
    p = (x, y)
    if validate and not (g2_on_curve(p) and g2_in_subgroup(p)):
        raise ValueError("InvalidData: point not on the curve / not in the r-torsion subgroup")
    return p


# ----------------------------------------------------------------------------------------------
# Montgomery-limb helpers (the C ABI carries arkworks' in-RAM representation: x * 2^256 mod m)
# ----------------------------------------------------------------------------------------------
def to_mont(x: int, m: int = Q) -> int: return x * MONT_R % m
def from_mont(x: int, m: int = Q) -> int: return x * inv_mod(MONT_R, m) % m


# ----------------------------------------------------------------------------------------------
# Radix-2 evaluation domain over Fr (ark-poly 0.4.2 Radix2EvaluationDomain)
# ----------------------------------------------------------------------------------------------
class Radix2Domain:
    """`Radix2EvaluationDomain::new(n)`: size = next power of two >= n, generator
    omega = 5^((r-1)/size) (src/vec.rs:36, src/kzg.rs:163, tests/laconic_ot.rs:81-85)."""

    def __init__(self, n: int):
        size = 1
        while size < n:
            size <<= 1
        self.size = size
        self.log_size = size.bit_length() - 1
        if self.log_size > FR_TWO_ADICITY:
            raise ValueError("domain too large")
        self.group_gen = pow(FR_GENERATOR, (R - 1) // size, R)
        self.group_gen_inv = inv_mod(self.group_gen, R)
        self.size_inv = inv_mod(size, R)

    def elements(self):
        out, x = [], 1
        for _ in range(self.size):
            out.append(x)
            x = x * self.group_gen % R
        return out

    @staticmethod
    def _ntt(vals, omega, add, mul_scalar, zero):
        n = len(vals)
        if n == 1:
            return list(vals)
        w2 = omega * omega % R
        even = Radix2Domain._ntt(vals[0::2], w2, add, mul_scalar, zero)
        odd = Radix2Domain._ntt(vals[1::2], w2, add, mul_scalar, zero)
        out = [zero] * n
        w = 1
        for i in range(n // 2):
            t = mul_scalar(odd[i], w)
            out[i] = add(even[i], t)
            out[i + n // 2] = add(even[i], mul_scalar(t, R - 1))
            w = w * omega % R
        return out

    def fft(self, coeffs):
        v = list(coeffs) + [0] * (self.size - len(coeffs))
        return self._ntt(v, self.group_gen, lambda a, b: (a + b) % R, lambda a, k: a * k % R, 0)

    def ifft(self, evals):
        v = list(evals) + [0] * (self.size - len(evals))
        out = self._ntt(v, self.group_gen_inv, lambda a, b: (a + b) % R, lambda a, k: a * k % R, 0)
        return [x * self.size_inv % R for x in out]

    def fft_g1(self, pts):
        v = list(pts) + [None] * (self.size - len(pts))
        return self._ntt(v, self.group_gen, g1_add, g1_mul, None)

    def ifft_g1(self, pts):
        v = list(pts) + [None] * (self.size - len(pts))
        out = self._ntt(v, self.group_gen_inv, g1_add, g1_mul, None)
        return [g1_mul(x, self.size_inv) for x in out]
