// Optimal-ate pairing on BN254: replaces `E::pairing` (ark-ec 0.4.2 models/bn/mod.rs) at
// src/kem.rs:30,58 and src/kzg.rs:148.
//
// Bit-exactness contract: the GT value must equal arkworks', i.e.
//   GT = miller(P, Q) ^ ((q^6-1)(q^2+1) * lambda),
//   lambda = q^3(12z^3+6z^2+4z-1) + q^2(12z^3+6z^2+6z) + q(12z^3+6z^2+4z) + (12z^3+12z^2+6z+1)
// (= 2z(6z^2+3z+1) times the textbook hard exponent).  Line functions may be scaled by any
// element of a proper subfield, so the Jacobian line formulas below (own derivation)
// need not match arkworks' homogeneous ones.
#pragma once
#include "ec.cuh"

namespace kb {

// Precomputed constants living in constant memory on the device / a static table on the host.
struct PairingConsts {
  FrobTable frob;   // gamma_{k,i} = xi^(i (q^k - 1)/6), k = 1..3, Montgomery form
  Fq2 tw_x, tw_y;   // xi^((q-1)/3), xi^((q-1)/2): q-power Frobenius on the twist
};

// BN parameter z = 4965661367192848881 and the signed digits of 6z+2 (LSB first) as used by
// ark-bn254 (only the value 6z+2 matters for the result).
#define KB_BN_Z 0x44E992B44A6909F1ull
KB_HD int ate_digit(int i) {
  // 65 digits, 2 bits each: 0 -> 0, 1 -> +1, 2 -> -1
  constexpr signed char d[65] = {0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0,
                                 1, 1, 1, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0,
                                 0, 1, 0, 1, 1};
  return d[i];
}

// Jacobian point on the twist used as the Miller-loop accumulator.
struct G2Jac { Fq2 x, y, z; };

// Line coefficients before evaluation at P: the line is  (l0 * yP) + (l1 * xP) w + l3 w^3.
struct Line { Fq2 l0, l1, l3; };

// Tangent line at T and T <- 2T.
//   lambda = 3X^2 / (2YZ);  scaled by 2YZ^3 = Z3 Z^2:
//   l0 = Z3 Z^2,  l1 = -3X^2 Z^2,  l3 = 3X^3 - 2Y^2.
KB_HD_NOINLINE Line line_dbl(G2Jac& t) {
  Fq2 a = sqr(t.x), b = sqr(t.y), c = sqr(b);
  Fq2 d = sqr(t.x + b) - a - c; d = dbl(d);
  Fq2 e = dbl(a) + a;
  Fq2 zz = sqr(t.z);
  Fq2 z3 = dbl(t.y * t.z);
  Line l;
  l.l3 = e * t.x - dbl(b);
  l.l1 = -(e * zz);
  l.l0 = z3 * zz;
  Fq2 x3 = sqr(e) - dbl(d);
  Fq2 c8 = dbl(dbl(dbl(c)));
  t.y = e * (d - x3) - c8;
  t.x = x3;
  t.z = z3;
  return l;
}

// Chord through T and affine Q, T <- T + Q.
//   H = x2 Z^2 - X, r = y2 Z^3 - Y, Z3 = Z H;  scaled by Z3:
//   l0 = Z3,  l1 = -r,  l3 = r x2 - Z3 y2.
KB_HD_NOINLINE Line line_add(G2Jac& t, const Fq2& x2, const Fq2& y2) {
  Fq2 zz = sqr(t.z);
  Fq2 h = x2 * zz - t.x;
  Fq2 r = y2 * (t.z * zz) - t.y;
  Fq2 z3 = t.z * h;
  Line l;
  l.l0 = z3;
  l.l1 = -r;
  l.l3 = r * x2 - z3 * y2;
  Fq2 h2 = sqr(h), h3 = h * h2, v = t.x * h2;
  Fq2 x3 = sqr(r) - h3 - dbl(v);
  t.y = r * (v - x3) - t.y * h3;
  t.x = x3;
  t.z = z3;
  return l;
}

KB_HD Fq12 apply_line(const Fq12& f, const Line& l, const Fq& xp, const Fq& yp) {
  return mul_by_line(f, mul_fq(l.l0, yp), mul_fq(l.l1, xp), l.l3);
}

// f_{6z+2,Q}(P) * l_{pi(Q)} * l_{-pi^2(Q)}; returns 1 when either input is infinity (arkworks
// filters such pairs out of the multi-Miller loop).
KB_HD_NOINLINE Fq12 miller_loop(const G1Affine& p, const G2Affine& q, const PairingConsts& pc) {
  Fq12 f = Fq12::one();
  if (p.is_inf() || q.is_inf()) return f;
  G2Jac t; t.x = q.x; t.y = q.y; t.z = Fq2::one();
  Fq2 nqy = -q.y;
  for (int i = 63; i >= 0; i--) {
    if (i != 63) f = sqr(f);
    Line l = line_dbl(t);
    f = apply_line(f, l, p.x, p.y);
    int d = ate_digit(i);
    if (d != 0) {
      l = line_add(t, q.x, d > 0 ? q.y : nqy);
      f = apply_line(f, l, p.x, p.y);
    }
  }
  // Q1 = pi(Q), Q2 = -pi^2(Q)
  Fq2 q1x = conj(q.x) * pc.tw_x, q1y = conj(q.y) * pc.tw_y;
  Fq2 q2x = conj(q1x) * pc.tw_x, q2y = -(conj(q1y) * pc.tw_y);
  Line l = line_add(t, q1x, q1y);
  f = apply_line(f, l, p.x, p.y);
  l = line_add(t, q2x, q2y);
  f = apply_line(f, l, p.x, p.y);
  return f;
}

// a^z for a in the cyclotomic subgroup
KB_HD_NOINLINE Fq12 cyclotomic_exp_z(const Fq12& a) {
  Fq12 r = a;
  for (int i = 61; i >= 0; i--) {  // z has 63 bits; top bit consumed by r = a
    r = cyclotomic_sqr(r);
    if ((KB_BN_Z >> i) & 1ull) r = r * a;
  }
  return r;
}

// f^((q^6-1)(q^2+1) lambda): easy part, then the Fuentes-Castaneda chain for lambda in the
// arrangement arkworks uses (y0..y16), with Granger-Scott squarings.
KB_HD_NOINLINE Fq12 final_exponentiation(const Fq12& f, const PairingConsts& pc) {
  Fq12 f1 = conj(f);
  Fq12 f2 = inv(f);
  Fq12 r = f1 * f2;                       // f^(q^6 - 1)
  f2 = r;
  r = frobenius(r, 2, pc.frob) * f2;      // ^(q^2 + 1): now unitary / cyclotomic
  Fq12 y0 = conj(cyclotomic_exp_z(r));    // r^-z
  Fq12 y1 = cyclotomic_sqr(y0);
  Fq12 y2 = cyclotomic_sqr(y1);
  Fq12 y3 = y2 * y1;
  Fq12 y4 = conj(cyclotomic_exp_z(y3));
  Fq12 y5 = cyclotomic_sqr(y4);
  Fq12 y6 = cyclotomic_exp_z(y5);         // (y5^-z)^-1
  y3 = conj(y3);
  Fq12 y7 = y6 * y4;
  Fq12 y8 = y7 * y3;
  Fq12 y9 = y8 * y1;
  Fq12 y10 = y8 * y4;
  Fq12 y11 = y10 * r;
  Fq12 y12 = frobenius(y9, 1, pc.frob);
  Fq12 y13 = y12 * y11;
  y8 = frobenius(y8, 2, pc.frob);
  Fq12 y14 = y8 * y13;
  Fq12 y15 = frobenius(conj(r) * y9, 3, pc.frob);
  return y15 * y14;
}

// ark-serialize `serialize_uncompressed` of Fq12 (src/kem.rs:31-32,60-61): 12 x 32 B little-endian
// canonical integers in tower order c0.c0.c0, c0.c0.c1, c0.c1.c0, ..., c1.c2.c1.
KB_HD void gt_to_words(const Fq12& a, uint32_t out[96]) {
  const Fq6* h[2] = {&a.c0, &a.c1};
  for (int j = 0; j < 2; j++) {
    const Fq2* c[3] = {&h[j]->c0, &h[j]->c1, &h[j]->c2};
    for (int i = 0; i < 3; i++) {
      Fq lo = fp_from_mont<FqParams>(c[i]->c0), hi = fp_from_mont<FqParams>(c[i]->c1);
      for (int k = 0; k < 8; k++) { out[(j * 3 + i) * 16 + k] = lo.v[k]; out[(j * 3 + i) * 16 + 8 + k] = hi.v[k]; }
    }
  }
}

}  // namespace kb
