// GLV scalar multiplication on BN254 G1 — `Mul<ScalarField>` on G1 at src/kzg.rs:57,135,190 and src/kem.rs:22 and the
// scalar-multiplication butterflies of the G1 transforms in `open_fk` (src/kzg.rs:182-200).
//
// G1 (y^2 = x^3 + 3) has the endomorphism phi(x, y) = (beta x, y) = lambda (x, y) with beta^3 = 1 in Fq and lambda^3 = 1
// in Fr.  A scalar k < r splits as k = k1 + k2 lambda (mod r) with |k1|, |k2| < 2^128, so k P = k1 P + k2 phi(P) needs
// half the doublings: 128 doublings + <= 66 additions instead of 254 + 64.  The result is the same group element, so the
// canonical affine bytes that leave the device do not change.
//
// Decomposition without division (constants in glv_gen.cuh, derived in tools/gen_consts.py): with the lattice basis
// (a1, -|b1|), (a2, b2) of {(x, y): x + y lambda = 0 mod r} and the scaled reciprocals g1c = floor(2^256 b2 / r),
// g2c = floor(2^256 |b1| / r):  c1 = (k g1c) >> 256, c2 = (k g2c) >> 256, k1 = k - c1 a1 - c2 a2, k2 = c1 |b1| - c2 b2,
// evaluated mod 2^256 and read as signed numbers (tests/test_hostemu_arith.py checks |k1|, |k2| < 2^128 and the congruence).
// Compiles for the host too (tests/hostemu, TEST ONLY).
#pragma once
#include "ec.cuh"
#include "glv_gen.cuh"

namespace kb {

// out[0..na+nb) = a * b, schoolbook on 32-bit limbs (cold code: a few hundred multiply-adds per scalar)
template <int NA, int NB>
KB_HD void glv_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
#pragma unroll
  for (int i = 0; i < NA + NB; i++) out[i] = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) {
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
      uint64_t t = (uint64_t)a[i] * b[j] + out[i + j] + carry;
      out[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    out[i + NB] = (uint32_t)carry;
  }
}
// x -= y over 8 limbs (mod 2^256)
KB_HD void glv_sub8(uint32_t* x, const uint32_t* y) {
  x[0] = sub_cc(x[0], y[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) x[i] = subc_cc(x[i], y[i]);
  x[7] = subc(x[7], y[7]);
}
// two's complement magnitude: returns the sign (1 = negative) and leaves |x|
KB_HD uint32_t glv_abs8(uint32_t* x) {
  const uint32_t neg = x[7] >> 31;
  if (neg) {
    uint32_t z[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { z[i] = x[i]; x[i] = 0; }
    glv_sub8(x, z);
  }
  return neg;
}

struct GlvSplit { uint32_t m1[5], m2[5]; uint32_t neg1, neg2; };   // k = (-1)^neg1 m1 + (-1)^neg2 m2 lambda (mod r), m < 2^128

// k: canonical (non-Montgomery) scalar below r
KB_HD GlvSplit glv_decompose(const uint32_t k[8]) {
  uint32_t g1c[3], g2c[5], a1[2], a2[4], b1[4], b2[2];
#pragma unroll
  for (int i = 0; i < 3; i++) g1c[i] = GlvParams::g1c(i);
#pragma unroll
  for (int i = 0; i < 5; i++) g2c[i] = GlvParams::g2c(i);
#pragma unroll
  for (int i = 0; i < 2; i++) { a1[i] = GlvParams::a1(i); b2[i] = GlvParams::b2(i); }
#pragma unroll
  for (int i = 0; i < 4; i++) { a2[i] = GlvParams::a2(i); b1[i] = GlvParams::b1abs(i); }
  uint32_t t1[11], t2[13];
  glv_mul<8, 3>(k, g1c, t1);          // c1 = t1[8..10]  (< 2^64: two limbs, the third is zero)
  glv_mul<8, 5>(k, g2c, t2);          // c2 = t2[8..12]  (< 2^127: four limbs)
  const uint32_t* c1 = t1 + 8;
  const uint32_t* c2 = t2 + 8;
  uint32_t p11[4], p22[8], q11[6], q22[6];
  glv_mul<2, 2>(c1, a1, p11);         // c1 a1 < 2^128
  glv_mul<4, 4>(c2, a2, p22);         // c2 a2 < 2^254
  glv_mul<2, 4>(c1, b1, q11);         // c1 |b1| < 2^191
  glv_mul<4, 2>(c2, b2, q22);         // c2 b2 < 2^191
  uint32_t k1[8], k2[8], w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) k1[i] = k[i];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = i < 4 ? p11[i] : 0u;
  glv_sub8(k1, w);
  glv_sub8(k1, p22);
#pragma unroll
  for (int i = 0; i < 8; i++) { k2[i] = i < 6 ? q11[i] : 0u; w[i] = i < 6 ? q22[i] : 0u; }
  glv_sub8(k2, w);
  GlvSplit s;
  s.neg1 = glv_abs8(k1);
  s.neg2 = glv_abs8(k2);
#pragma unroll
  for (int i = 0; i < 5; i++) { s.m1[i] = k1[i]; s.m2[i] = k2[i]; }
  return s;
}

// k * P for a canonical scalar k < r; 4-bit windows over both half-scalars, one shared table:
// d2 * (+-phi(P)) = phi(+-T[d2]) costs one product by beta and a conditional negation.
KB_HD_NOINLINE G1 g1_mul_glv(const G1& p, const uint32_t k[8]) {
  if (p.is_inf()) return p;
  const GlvSplit s = glv_decompose(k);
  Fq beta;
#pragma unroll
  for (int i = 0; i < 8; i++) beta.v[i] = GlvParams::beta(i);
  G1 tab[16];
  tab[0] = G1::infinity();
  tab[1] = p;
  if (s.neg1) tab[1].y = -tab[1].y;
  for (int i = 2; i < 16; i++) tab[i] = (i & 1) ? ec_add(tab[i - 1], tab[1]) : ec_dbl(tab[i >> 1]);
  const bool flip = s.neg1 != s.neg2;
  G1 acc = G1::infinity();
  for (int w = 32; w >= 0; w--) {   // 33 windows = 132 bits >= the 128-bit bound on |k1|, |k2|
    acc = ec_dbl(ec_dbl(ec_dbl(ec_dbl(acc))));
    const uint32_t d1 = (s.m1[w >> 3] >> ((w & 7) * 4)) & 15u;
    const uint32_t d2 = (s.m2[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d1) acc = ec_add(acc, tab[d1]);
    if (d2) {
      G1 t = tab[d2];
      t.x = t.x * beta;
      if (flip) t.y = -t.y;
      acc = ec_add(acc, t);
    }
  }
  return acc;
}

}  // namespace kb
