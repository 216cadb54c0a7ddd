// BN254 extension tower: Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v),
// xi = 9 + u  (ark-bn254 0.4.0 fields/fq2.rs, fq6.rs, fq12.rs; Cargo.lock:30-31).
// Replaces the ark-ff tower used underneath `Pairing::pairing` (src/kem.rs:30,58; src/kzg.rs:148).
// All elements are in Montgomery form.  Results are canonical field elements, so the choice of
// multiplication algorithm here (Karatsuba, sparse products, Granger-Scott squaring) is free.
#pragma once
#include "fp.cuh"

namespace kb {

// ------------------------------------------------------------------------------------------ Fq2
struct Fq2 {
  Fq c0, c1;
  static KB_HD Fq2 zero() { Fq2 r; r.c0 = Fq::zero(); r.c1 = Fq::zero(); return r; }
  static KB_HD Fq2 one() { Fq2 r; r.c0 = Fq::one(); r.c1 = Fq::zero(); return r; }
  KB_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  KB_HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
  KB_HD bool operator!=(const Fq2& b) const { return !(*this == b); }
};

KB_HD Fq2 operator+(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
KB_HD Fq2 operator-(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
KB_HD Fq2 operator-(const Fq2& a) { Fq2 r; r.c0 = -a.c0; r.c1 = -a.c1; return r; }
KB_HD Fq2 dbl(const Fq2& a) { Fq2 r; r.c0 = dbl(a.c0); r.c1 = dbl(a.c1); return r; }
KB_HD Fq2 conj(const Fq2& a) { Fq2 r; r.c0 = a.c0; r.c1 = -a.c1; return r; }

// Karatsuba: 3 base multiplications
KB_FN Fq2 operator*(const Fq2& a, const Fq2& b) {
  Fq t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
  Fq s = (a.c0 + a.c1) * (b.c0 + b.c1);
  Fq2 r; r.c0 = t0 - t1; r.c1 = s - t0 - t1; return r;
}
// complex squaring: 2 base multiplications
KB_FN Fq2 sqr(const Fq2& a) {
  Fq t = a.c0 * a.c1;
  Fq2 r; r.c0 = (a.c0 + a.c1) * (a.c0 - a.c1); r.c1 = dbl(t); return r;
}
KB_HD Fq2 mul_fq(const Fq2& a, const Fq& k) { Fq2 r; r.c0 = a.c0 * k; r.c1 = a.c1 * k; return r; }
// multiply by xi = 9 + u:  (9 a0 - a1) + (9 a1 + a0) u
KB_HD Fq2 mul_xi(const Fq2& a) {
  Fq t0 = dbl(dbl(dbl(a.c0))) + a.c0;  // 9 a0
  Fq t1 = dbl(dbl(dbl(a.c1))) + a.c1;  // 9 a1
  Fq2 r; r.c0 = t0 - a.c1; r.c1 = t1 + a.c0; return r;
}
KB_FN Fq2 inv(const Fq2& a) {
  Fq d = inv(sqr(a.c0) + sqr(a.c1));
  Fq2 r; r.c0 = a.c0 * d; r.c1 = -(a.c1 * d); return r;
}

// ------------------------------------------------------------------------------------------ Fq6
struct Fq6 {
  Fq2 c0, c1, c2;
  static KB_HD Fq6 zero() { Fq6 r; r.c0 = Fq2::zero(); r.c1 = Fq2::zero(); r.c2 = Fq2::zero(); return r; }
  static KB_HD Fq6 one() { Fq6 r; r.c0 = Fq2::one(); r.c1 = Fq2::zero(); r.c2 = Fq2::zero(); return r; }
  KB_HD bool operator==(const Fq6& b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
};
KB_HD Fq6 operator+(const Fq6& a, const Fq6& b) { Fq6 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; r.c2 = a.c2 + b.c2; return r; }
KB_HD Fq6 operator-(const Fq6& a, const Fq6& b) { Fq6 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; r.c2 = a.c2 - b.c2; return r; }
KB_HD Fq6 operator-(const Fq6& a) { Fq6 r; r.c0 = -a.c0; r.c1 = -a.c1; r.c2 = -a.c2; return r; }
KB_HD Fq6 mul_v(const Fq6& a) { Fq6 r; r.c0 = mul_xi(a.c2); r.c1 = a.c0; r.c2 = a.c1; return r; }

// Karatsuba over Fq2: 6 Fq2 products
KB_HD_NOINLINE Fq6 operator*(const Fq6& a, const Fq6& b) {
  Fq2 v0 = a.c0 * b.c0, v1 = a.c1 * b.c1, v2 = a.c2 * b.c2;
  Fq6 r;
  r.c0 = v0 + mul_xi((a.c1 + a.c2) * (b.c1 + b.c2) - v1 - v2);
  r.c1 = (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1 + mul_xi(v2);
  r.c2 = (a.c0 + a.c2) * (b.c0 + b.c2) - v0 - v2 + v1;
  return r;
}
KB_HD_NOINLINE Fq6 inv(const Fq6& a) {
  Fq2 t0 = sqr(a.c0) - mul_xi(a.c1 * a.c2);
  Fq2 t1 = mul_xi(sqr(a.c2)) - a.c0 * a.c1;
  Fq2 t2 = sqr(a.c1) - a.c0 * a.c2;
  Fq2 d = inv(a.c0 * t0 + mul_xi(a.c2 * t1 + a.c1 * t2));
  Fq6 r; r.c0 = t0 * d; r.c1 = t1 * d; r.c2 = t2 * d; return r;
}

// ------------------------------------------------------------------------------------------ Fq12
struct Fq12 {
  Fq6 c0, c1;
  static KB_HD Fq12 one() { Fq12 r; r.c0 = Fq6::one(); r.c1 = Fq6::zero(); return r; }
  KB_HD bool operator==(const Fq12& b) const { return c0 == b.c0 && c1 == b.c1; }
  // coefficient of w^i, i = 0..5 (Fq12 = Fq2[w]/(w^6 - xi), v = w^2)
  KB_HD Fq2& w(int i) { Fq6& h = (i & 1) ? c1 : c0; return (i >> 1) == 0 ? h.c0 : ((i >> 1) == 1 ? h.c1 : h.c2); }
  KB_HD const Fq2& w(int i) const { const Fq6& h = (i & 1) ? c1 : c0; return (i >> 1) == 0 ? h.c0 : ((i >> 1) == 1 ? h.c1 : h.c2); }
};
KB_HD Fq12 conj(const Fq12& a) { Fq12 r; r.c0 = a.c0; r.c1 = -a.c1; return r; }

KB_HD_NOINLINE Fq12 operator*(const Fq12& a, const Fq12& b) {
  Fq6 t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
  Fq12 r;
  r.c1 = (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1;
  r.c0 = t0 + mul_v(t1);
  return r;
}
// complex squaring: 2 Fq6 products
KB_HD_NOINLINE Fq12 sqr(const Fq12& a) {
  Fq6 t = a.c0 * a.c1;
  Fq12 r;
  r.c0 = (a.c0 + a.c1) * (a.c0 + mul_v(a.c1)) - t - mul_v(t);
  r.c1 = t + t;
  return r;
}
KB_HD_NOINLINE Fq12 inv(const Fq12& a) {
  Fq6 d = inv(a.c0 * a.c0 - mul_v(a.c1 * a.c1));
  Fq12 r; r.c0 = a.c0 * d; r.c1 = -(a.c1 * d); return r;
}

// Granger-Scott squaring for elements of the cyclotomic subgroup (after the easy part of the
// final exponentiation).  With g = sum g_i w^i viewed as three Fq4 pairs (g0,g3), (g1,g4), (g2,g5)
// [Fq4 = Fq2[s]/(s^2 - xi)]:  A = (g0 + g3 s)^2, B = (g1 + g4 s)^2 ... giving
//   h0 = 3 A0 - 2 g0, h3 = 3 A1 + 2 g3, etc.   (formulas derived in DESIGN.md; checked against
// the generic square in tests/hostemu).
KB_HD void fq4_sqr(const Fq2& a, const Fq2& b, Fq2& r0, Fq2& r1) {
  Fq2 a2 = sqr(a), b2 = sqr(b);
  r0 = a2 + mul_xi(b2);
  r1 = sqr(a + b) - a2 - b2;
}
KB_HD_NOINLINE Fq12 cyclotomic_sqr(const Fq12& g) {
  // w-basis coefficients: g = (g0 + g3 w^3) + (g1 + g4 w^3) w + (g2 + g5 w^3) w^2, with (w^3)^2 = xi
  // square in Fq12 = Fq4[w]/(w^3 - s), s = w^3, for unitary elements:
  //   h = 3 * (A + C*s*w ... ) pattern below follows from g^(q^6) = g^{-1}.
  const Fq2 &g0 = g.w(0), &g1 = g.w(1), &g2 = g.w(2), &g3 = g.w(3), &g4 = g.w(4), &g5 = g.w(5);
  Fq2 A0, A1, B0, B1, C0, C1;
  fq4_sqr(g0, g3, A0, A1);  // (g0 + g3 s)^2
  fq4_sqr(g1, g4, B0, B1);  // (g1 + g4 s)^2
  fq4_sqr(g2, g5, C0, C1);  // (g2 + g5 s)^2
  // Fq12 over Fq4 with t = w, t^3 = s:  g = a + b t + c t^2,  a=(g0,g3), b=(g1,g4), c=(g2,g5).
  // For unitary g:  g^2 = (3a^2 - 2 conj(a)) + (3 s c^2 + 2 conj(b)) t + (3 b^2 - 2 conj(c)) t^2
  // where conj on Fq4 negates the s-coefficient.
  Fq12 h;
  // a' = 3 A - 2 conj(a)
  h.w(0) = dbl(A0 - g0) + A0;
  h.w(3) = dbl(A1 + g3) + A1;
  // b' = 3 s C + 2 conj(b);  s * (C0 + C1 s) = xi C1 + C0 s
  Fq2 sC0 = mul_xi(C1), sC1 = C0;
  h.w(1) = dbl(sC0 + g1) + sC0;
  h.w(4) = dbl(sC1 - g4) + sC1;
  // c' = 3 B - 2 conj(c)
  h.w(2) = dbl(B0 - g2) + B0;
  h.w(5) = dbl(B1 + g5) + B1;
  return h;
}

// Sparse product used by the Miller loop: a * (l0 + l1 w + l3 w^3) with l0 in Fq2.
KB_HD_NOINLINE Fq12 mul_by_line(const Fq12& a, const Fq2& l0, const Fq2& l1, const Fq2& l3) {
  // c_k = sum_{i+j=k} a_i l_j with w^6 = xi wrap-around
  Fq12 r;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    Fq2 t = a.w(k) * l0;
    int i1 = (k + 5) % 6, i3 = (k + 3) % 6;  // a_{k-1} * l1, a_{k-3} * l3
    Fq2 p1 = a.w(i1) * l1, p3 = a.w(i3) * l3;
    if (k < 1) p1 = mul_xi(p1);
    if (k < 3) p3 = mul_xi(p3);
    r.w(k) = t + p1 + p3;
  }
  return r;
}

// Frobenius constants gamma_{k,i} = xi^(i (q^k - 1)/6) are passed in by the caller (device
// constant memory / host table); see pairing.cuh.
struct FrobTable { Fq2 g[3][6]; };

KB_HD_NOINLINE Fq12 frobenius(const Fq12& a, int k, const FrobTable& tab) {
  Fq12 r;
  for (int i = 0; i < 6; i++) {
    Fq2 x = a.w(i);
    if (k & 1) x = conj(x);
    r.w(i) = x * tab.g[k - 1][i];
  }
  return r;
}

}  // namespace kb
