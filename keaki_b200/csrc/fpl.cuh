// Lazy-reduction primitives for the BN254 base field: separated 256x256 -> 512-bit integer product and Montgomery
// reduction, and the Fq2 product built on them (Karatsuba on the integer products, ONE reduction per coordinate:
// 3 x 64 + 2 x 64 IMAD.WIDE instead of the 4 x 64 + 2 x 64 of the fused two-product form and the 3 x 128 of three
// full Montgomery products).  Replaces ark-ff 0.4.2's `Fp2::mul` / `QuadExtField::mul_assign` underneath
// `E::pairing` (src/kem.rs:30,58; src/kzg.rs:148).  Same limb layout and Montgomery constant as fp.cuh, so values
// move freely between the fused multiplier and these routines.
//
// Also home of the small unreduced-arithmetic helpers shared by the pairing kernels (pairing_vm.cuh, pairing_st.cuh).
#pragma once
#include "tower.cuh"

namespace kb {
namespace vm {

// a + b without reduction (callers keep the sum below 2^256)
KB_HD Fq add_nr(const Fq& a, const Fq& b) {
  Fq r;
  r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
  r.v[7] = addc(a.v[7], b.v[7]);
  return r;
}
// x >= p ? x - p : x
KB_HD Fq csub(const Fq& x) { Fq r = x; fp_reduce_once<FqParams>(r.v); return r; }
// p - a without the zero test: result in [1, p] (a valid multiplier input)
KB_HD Fq neg_nz(const Fq& a) {
  Fq r;
  r.v[0] = sub_cc(FqParams::mod(0), a.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = subc_cc(FqParams::mod(i), a.v[i]);
  r.v[7] = subc(FqParams::mod(7), a.v[7]);
  return r;
}
KB_HD Fq sel(bool c, const Fq& a, const Fq& b) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}

// (9 x + y) mod p for x < p, y <= p: the 9-limb value w < 10 p is reduced with a quotient estimate from its top
// 32 bits, q = floor(top * floor(2^59 / (floor(p / 2^226) + 1)) / 2^59) in {floor(w / p) - 1, floor(w / p)}
// (checked exhaustively at the extremes in tests/test_pairing_prog.py), then one conditional subtraction.
KB_HD Fq mul9_add(const Fq& x, const Fq& y) {
  uint32_t w[9];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (uint64_t)x.v[i] * 9u + y.v[i]; w[i] = (uint32_t)c; c >>= 32; }
  w[8] = (uint32_t)c;
  const uint32_t top = (w[8] << 30) | (w[7] >> 2);
  const uint32_t q = mul_hi(top, 0xa948e8c0u) >> 27;
  uint32_t qp[9];
  c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (uint64_t)FqParams::mod(i) * q; qp[i] = (uint32_t)c; c >>= 32; }
  Fq r;
  r.v[0] = sub_cc(w[0], qp[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = subc_cc(w[i], qp[i]);
  r.v[7] = subc(w[7], qp[7]);   // w - q p < 2 p < 2^255: the ninth limb cancels
  return csub(r);
}

}  // namespace vm

namespace lz {

struct W { uint32_t v[16]; };   // 512-bit non-negative integer, little-endian limbs

// 2 q^2 as 16 limbs (an offset that is 0 mod q: keeps a0 b0 - a1 b1 non-negative for a1 < 2q, b1 < q)
KB_LIMB_TABLE(q2lo, 0x4ebad362u, 0x76a8b144u, 0x13d58202u, 0x4c040e5au, 0xdb2d95b9u, 0x94a03138u, 0xf4248590u, 0x08d13d2au)
KB_LIMB_TABLE(q2hi, 0x698d671au, 0x4ddbf4b8u, 0x2c6eac0cu, 0x60170aa2u, 0x0691a439u, 0xb334def8u, 0xec797f38u, 0x124b8970u)

// t = a * b, plain integers.  Even/odd accumulator split as in fp.cuh: E holds the partial products whose limb
// position i + j is even (aligned 64-bit pairs at positions 0, 2, ...), O those at odd positions (O[k] is position
// k + 1), so every row of four products is one IMAD.WIDE carry chain.  After row j the running sum is
// a * (b mod 2^(32 (j + 1))) < 2^(32 (j + 9)): a chain that ends at position j + 7 carries into position j + 8 (one
// addc, which cannot overflow), a chain that ends at position j + 8 cannot carry out at all.
KB_HD W mul_wide(const Fq& a, const Fq& b) {
  uint32_t E[16], O[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { E[i] = 0; O[i] = 0; }
  row_mul(E, a.v, b.v[0]);
  row_mul(O, a.v + 1, b.v[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) {
    if (j & 1) {
      row_mad(O + j - 1, a.v, b.v[j]);       // positions j .. j + 7
      O[j + 7] = addc(O[j + 7], 0);
      row_mad(E + j + 1, a.v + 1, b.v[j]);   // positions j + 1 .. j + 8
    } else {
      row_mad(E + j, a.v, b.v[j]);
      E[j + 8] = addc(E[j + 8], 0);
      row_mad(O + j, a.v + 1, b.v[j]);
    }
  }
  W t;
  t.v[0] = E[0];
  t.v[1] = add_cc(E[1], O[0]);
#pragma unroll
  for (int i = 2; i < 15; i++) t.v[i] = addc_cc(E[i], O[i - 1]);
  t.v[15] = addc(E[15], O[14]);
  return t;
}

KB_HD W wsub(const W& a, const W& b) {   // a - b, a >= b
  W r;
  r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 15; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
  r.v[15] = subc(a.v[15], b.v[15]);
  return r;
}
KB_HD W wadd_q2(const W& a) {   // a + 2 q^2
  W r;
  r.v[0] = add_cc(a.v[0], q2lo(0));
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = addc_cc(a.v[i], q2lo(i));
#pragma unroll
  for (int i = 0; i < 7; i++) r.v[8 + i] = addc_cc(a.v[8 + i], q2hi(i));
  r.v[15] = addc(a.v[15], q2hi(7));
  return r;
}

// one word of the reduction: the running value is E + O 2^32 (E at positions 0..7, O at 1..8); adds m q with
// m = -E[0] / q mod 2^32, which clears position 0.
template <class P>
KB_HD void redc_first(uint32_t* ev, uint32_t* od) {
  const uint32_t m = mul_lo(ev[0], P::inv);
#pragma unroll
  for (int k = 0; k < 8; k += 2) { od[k] = mul_lo(P::mod(k + 1), m); od[k + 1] = mul_hi(P::mod(k + 1), m); }
  row_mad_mod<P, 0>(ev, m);
  od[7] = addc(od[7], 0);
}
// next word: `ev` is the previous step's odd vector (now even-aligned after the division by 2^32), `od` the previous
// even vector whose limb 1 is folded into ev[0] and whose limbs 2..7 become the new odd vector (shifted by one pair).
template <class P>
KB_HD void redc_next(uint32_t* ev, uint32_t* od) {
  ev[0] = add_cc(ev[0], od[1]);
  const uint32_t m = mul_lo(ev[0], P::inv);   // mul.lo does not touch the carry flag
#pragma unroll
  for (int k = 0; k < 6; k += 2) { od[k] = madc_lo_cc(P::mod(k + 1), m, od[k + 2]); od[k + 1] = madc_hi_cc(P::mod(k + 1), m, od[k + 3]); }
  od[6] = madc_lo_cc(P::mod(7), m, 0);
  od[7] = madc_hi_cc(P::mod(7), m, 0);        // cannot carry out
  row_mad_mod<P, 0>(ev, m);
  od[7] = addc(od[7], 0);
}

// t R^-1 mod q, fully reduced, for t < q R:  (t_lo + m q) / R <= q, plus t_hi < q, then one conditional subtraction.
template <class P>
KB_HD Fp<P> redc(const W& t) {
  uint32_t ev[8], od[8];
#pragma unroll
  for (int i = 0; i < 8; i++) ev[i] = t.v[i];
  redc_first<P>(ev, od);
  redc_next<P>(od, ev);
  redc_next<P>(ev, od);
  redc_next<P>(od, ev);
  redc_next<P>(ev, od);
  redc_next<P>(od, ev);
  redc_next<P>(ev, od);
  redc_next<P>(od, ev);
  Fp<P> r;
  r.v[0] = add_cc(ev[0], od[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(ev[i], od[i + 1]);
  r.v[7] = addc(ev[7], 0);
  r.v[0] = add_cc(r.v[0], t.v[8]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(r.v[i], t.v[8 + i]);
  r.v[7] = addc(r.v[7], t.v[15]);
  fp_reduce_once<P>(r.v);
  return r;
}

// Fq2 product.  b fully reduced (< q per coordinate), a below 2q per coordinate (an unreduced sum of two reduced values
// is a legal a):  c0 = a0 b0 - a1 b1 + 2 q^2 in (0, 4 q^2),  c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1 = a0 b1 + a1 b0
// < 4 q^2 (exact as integers because the operand sums are NOT reduced); both are below q R = 5.29 q^2.
KB_HD Fq2 fq2_mul_lazy(const Fq2& a, const Fq2& b) {
  const W t0 = mul_wide(a.c0, b.c0), t1 = mul_wide(a.c1, b.c1);
  const W t2 = mul_wide(vm::add_nr(a.c0, a.c1), vm::add_nr(b.c0, b.c1));
  Fq2 r;
  r.c1 = redc<FqParams>(wsub(wsub(t2, t0), t1));
  r.c0 = redc<FqParams>(wsub(wadd_q2(t0), t1));
  return r;
}

}  // namespace lz
}  // namespace kb
