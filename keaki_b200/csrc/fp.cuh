// BN254 prime-field arithmetic on 8 x 32-bit limbs, Montgomery form with R = 2^256.
//
// The in-memory representation is bit-identical to arkworks' `Fp256<MontBackend<_, 4>>` (4 x u64
// little-endian limbs, same R), so the Rust side can hand `fe.0.0` to the C ABI unchanged
// (SURVEY.md §8b).  Replaces ark-ff 0.4.2 (Cargo.lock:58-59) at the call sites listed in
// SURVEY.md §8b.
//
// Multiplier: word-serial Montgomery multiplication with the accumulator split into an
// even-aligned and an odd-aligned limb vector, so every 32x32 partial product lands on an aligned
// 64-bit register pair and a whole row is ONE carry chain of `mad.lo.cc / madc.hi.cc` pairs that
// ptxas fuses into IMAD.WIDE.U32[.X] (verified in the committed SASS listing, profiles/).
// 128 IMAD.WIDE + 8 IMAD per product; IADD3 fix-ups ride the otherwise idle ALU pipe.
//
// The file compiles in two modes:
//   * nvcc, device code: primitives are inline PTX (the product).
//   * any C++ compiler, host code: primitives are emulated with an explicit carry flag, so the
//     very same algorithm text is unit-tested on a CPU-only box (tests/hostemu, TEST ONLY — the
//     shipped library never calls host arithmetic).
#pragma once
#include <stdint.h>

// Inlining policy.  ptxas time explodes when the ~200-instruction multiplier is force-inlined into
// whole curve / tower formulas, so by default (cold code: table builds, reductions, the generic
// pairing) KB_FN functions are real calls with operands passed by value in registers.  A hot
// kernel lives in its own translation unit and defines KB_INLINE_ALL before including this file.
#if defined(__CUDACC__)
#define KB_HD __host__ __device__ __forceinline__
#define KB_D __device__ __forceinline__
#define KB_HD_NOINLINE inline __host__ __device__ __noinline__
#ifdef KB_INLINE_ALL
#define KB_FN __host__ __device__ __forceinline__
#else
#define KB_FN inline __host__ __device__ __noinline__
#endif
#else
#define KB_HD inline
#define KB_D inline
#define KB_HD_NOINLINE inline
#define KB_FN inline
#endif

namespace kb {

// ------------------------------------------------------------------------------------------
// Carry-chain primitives
// ------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
KB_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KB_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KB_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KB_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KB_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KB_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KB_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KB_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KB_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
KB_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
KB_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
KB_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
KB_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// Host emulation of the PTX condition-code flag (test harness only).
struct CarryFlag { uint32_t cf; };
inline CarryFlag& kb_cf() { static thread_local CarryFlag f{0}; return f; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; kb_cf().cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + kb_cf().cf; kb_cf().cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + kb_cf().cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; kb_cf().cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - kb_cf().cf; kb_cf().cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - kb_cf().cf; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
#endif

// ------------------------------------------------------------------------------------------
// Field parameters (ark-bn254 0.4.0 fields/fq.rs, fr.rs; values cross-checked in tests against
// the oracle).  Limb getters are constexpr functions so fully unrolled code sees immediates.
// ------------------------------------------------------------------------------------------
#define KB_LIMB_TABLE(NAME, ...) \
  static constexpr KB_HD uint32_t NAME(int i) { constexpr uint32_t t[8] = {__VA_ARGS__}; return t[i]; }

struct FqParams {
  KB_LIMB_TABLE(mod, 0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u)
  KB_LIMB_TABLE(one, 0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u)
  KB_LIMB_TABLE(r2, 0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u)
  static constexpr uint32_t inv = 0xe4866389u;  // -q^{-1} mod 2^32
};
struct FrParams {
  KB_LIMB_TABLE(mod, 0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u)
  KB_LIMB_TABLE(one, 0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u)
  KB_LIMB_TABLE(r2, 0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u)
  static constexpr uint32_t inv = 0xefffffffu;  // -r^{-1} mod 2^32
};

// ------------------------------------------------------------------------------------------
// Row operations of the even/odd Montgomery multiplier.  `x` points at a[0] (even rows) or a[1]
// (odd rows); a row uses x[0], x[2], x[4], x[6].
// ------------------------------------------------------------------------------------------
// acc[2k], acc[2k+1] = x[2k] * s                       (no carries: 4 independent wide products)
KB_HD void row_mul(uint32_t* acc, const uint32_t* x, uint32_t s) {
#pragma unroll
  for (int k = 0; k < 8; k += 2) { acc[k] = mul_lo(x[k], s); acc[k + 1] = mul_hi(x[k], s); }
}
// acc += x_row * s as one 8-limb carry chain; the carry out of limb 7 is left in CF
KB_HD void row_mad(uint32_t* acc, const uint32_t* x, uint32_t s) {
  acc[0] = mad_lo_cc(x[0], s, acc[0]);
  acc[1] = madc_hi_cc(x[0], s, acc[1]);
#pragma unroll
  for (int k = 2; k < 8; k += 2) { acc[k] = madc_lo_cc(x[k], s, acc[k]); acc[k + 1] = madc_hi_cc(x[k], s, acc[k + 1]); }
}
// acc[k] = x_row * s + acc[k+2] (accumulator shifted down by one 64-bit pair), carry-in from CF;
// the top pair has no addend.  No carry out is possible (hi word of a product <= 2^32 - 2).
KB_HD void row_mad_shift(uint32_t* acc, const uint32_t* x, uint32_t s) {
#pragma unroll
  for (int k = 0; k < 6; k += 2) { acc[k] = madc_lo_cc(x[k], s, acc[k + 2]); acc[k + 1] = madc_hi_cc(x[k], s, acc[k + 3]); }
  acc[6] = madc_lo_cc(x[6], s, 0);
  acc[7] = madc_hi_cc(x[6], s, 0);  // carry-out is always 0; the .cc form lets ptxas fuse the pair into IMAD.WIDE.X
}
// same rows with the modulus as the (immediate) multiplicand
template <class P, int ODD>
KB_HD void row_mad_mod(uint32_t* acc, uint32_t s) {
  acc[0] = mad_lo_cc(P::mod(ODD), s, acc[0]);
  acc[1] = madc_hi_cc(P::mod(ODD), s, acc[1]);
#pragma unroll
  for (int k = 2; k < 8; k += 2) { acc[k] = madc_lo_cc(P::mod(k + ODD), s, acc[k]); acc[k + 1] = madc_hi_cc(P::mod(k + ODD), s, acc[k + 1]); }
}

// One word of the multiplier.  On entry V = E + O * 2^32 (E = `ev`, limbs at positions 0..7; O =
// `od`, positions 1..8), except that for !FIRST the previous step left E shifted: its limb 1 still
// has to be folded into O[0] and its limbs 2..7 are the new odd vector.  On exit the roles of the
// two arrays are swapped (callers alternate the arguments).
template <class P, bool FIRST>
KB_HD void mont_step(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi) {
  if (FIRST) {
    row_mul(od, a + 1, bi);
    row_mul(ev, a, bi);
  } else {
    ev[0] = add_cc(ev[0], od[1]);    // fold the stray limb; carry feeds the odd row (position 1)
    row_mad_shift(od, a + 1, bi);    // od = a_odd * bi + (od >> 64)
    row_mad(ev, a, bi);              // ev += a_even * bi
    od[7] = addc(od[7], 0);          // carry out of position 7 lands on position 8
  }
  uint32_t m = mul_lo(ev[0], P::inv);
  row_mad_mod<P, 1>(od, m);          // cannot carry out: V < 2^288
  row_mad_mod<P, 0>(ev, m);          // ev[0] becomes 0
  od[7] = addc(od[7], 0);
}

template <class P>
struct Fp {
  uint32_t v[8];

  static KB_HD Fp zero() { Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
  static KB_HD Fp one() { Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::one(i); return r; }
  static KB_HD Fp r2() { Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::r2(i); return r; }

  KB_HD bool is_zero() const { uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i]; return o == 0; }
  KB_HD bool operator==(const Fp& b) const { uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i]; return o == 0; }
  KB_HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// r = x - mod if x >= mod else x   (x < 2 * mod)
template <class P>
KB_HD void fp_reduce_once(uint32_t* x) {
  uint32_t t[8];
  t[0] = sub_cc(x[0], P::mod(0));
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = subc_cc(x[i], P::mod(i));
  uint32_t borrow = subc(0, 0);  // 0xffffffff if x < mod
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = borrow ? x[i] : t[i];
}

template <class P>
KB_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
  r.v[7] = addc(a.v[7], b.v[7]);  // < 2^255: no carry out
  fp_reduce_once<P>(r.v);
  return r;
}

template <class P>
KB_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r;
  r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = subc(0, 0);  // all-ones if a < b
  r.v[0] = add_cc(r.v[0], P::mod(0) & borrow);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(r.v[i], P::mod(i) & borrow);
  r.v[7] = addc(r.v[7], P::mod(7) & borrow);
  return r;
}

template <class P>
KB_HD Fp<P> fp_neg(const Fp<P>& a) {
  if (a.is_zero()) return a;
  Fp<P> r;
  r.v[0] = sub_cc(P::mod(0), a.v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = subc_cc(P::mod(i), a.v[i]);
  r.v[7] = subc(P::mod(7), a.v[7]);
  return r;
}

template <class P>
KB_HD Fp<P> fp_dbl(const Fp<P>& a) { return fp_add<P>(a, a); }

// Montgomery product a * b * R^{-1} mod p, fully reduced.
template <class P>
KB_HD Fp<P> fp_mul_inl(const Fp<P>& a, const Fp<P>& b) {
  uint32_t ev[8], od[8];
  mont_step<P, true>(ev, od, a.v, b.v[0]);
  mont_step<P, false>(od, ev, a.v, b.v[1]);
  mont_step<P, false>(ev, od, a.v, b.v[2]);
  mont_step<P, false>(od, ev, a.v, b.v[3]);
  mont_step<P, false>(ev, od, a.v, b.v[4]);
  mont_step<P, false>(od, ev, a.v, b.v[5]);
  mont_step<P, false>(ev, od, a.v, b.v[6]);
  mont_step<P, false>(od, ev, a.v, b.v[7]);
  // merge: result = ev + (od >> 32)
  Fp<P> r;
  r.v[0] = add_cc(ev[0], od[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(ev[i], od[i + 1]);
  r.v[7] = addc(ev[7], 0);
  fp_reduce_once<P>(r.v);
  return r;
}

template <class P>
KB_FN Fp<P> fp_mul(Fp<P> a, Fp<P> b) { return fp_mul_inl<P>(a, b); }
template <class P>
KB_FN Fp<P> fp_sqr(Fp<P> a) { return fp_mul_inl<P>(a, a); }

// Montgomery form -> canonical integer (one Montgomery product by 1)
template <class P>
KB_HD Fp<P> fp_from_mont(const Fp<P>& a) {
  Fp<P> o = Fp<P>::zero();
  o.v[0] = 1;
  return fp_mul<P>(a, o);
}
template <class P>
KB_HD Fp<P> fp_to_mont(const Fp<P>& a) { return fp_mul<P>(a, Fp<P>::r2()); }

// 256-bit helpers of the inversion below (plain integers, no modular meaning)
KB_HD void u256_shr1(uint32_t* x) {
#pragma unroll
  for (int i = 0; i < 7; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
  x[7] >>= 1;
}
KB_HD void u256_shl1(uint32_t* x) {
#pragma unroll
  for (int i = 7; i > 0; i--) x[i] = (x[i] << 1) | (x[i - 1] >> 31);
  x[0] <<= 1;
}
KB_HD void u256_add(uint32_t* x, const uint32_t* y) {
  x[0] = add_cc(x[0], y[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) x[i] = addc_cc(x[i], y[i]);
  x[7] = addc(x[7], y[7]);
}
// x -= y; returns all-ones if the subtraction borrowed (x < y)
KB_HD uint32_t u256_sub(uint32_t* x, const uint32_t* y) {
  x[0] = sub_cc(x[0], y[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) x[i] = subc_cc(x[i], y[i]);
  return subc(0, 0);
}

// Inverse by Kaliski's binary "almost Montgomery inverse" (shifts, additions and subtractions on the ALU pipe: about a
// quarter of the issue slots of the Fermat ladder a^(p-2), and none of them on the multiplier pipe).  a = 0 -> 0.
// Phase 1 keeps u s + v r = p with u = p, v = a, r = 0, s = 1 and ends with v = 0, u = 1, r = -a^-1 2^k (mod p),
// r < 2p, bitlen(p) <= k <= 2 bitlen(p).  Phase 2 removes 2^k and restores the Montgomery factor: the input is
// a R, so the wanted a^-1 R equals (a R)^-1 R^2 = x 2^(512 - k) for x = (a R)^-1 2^k.
template <class P>
KB_HD_NOINLINE Fp<P> fp_inv(const Fp<P>& a) {
  if (a.is_zero()) return a;
  uint32_t u[8], v[8], r[8], s[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { u[i] = P::mod(i); v[i] = a.v[i]; r[i] = 0; s[i] = 0; }
  s[0] = 1;
  uint32_t k = 0;
  for (;;) {
    uint32_t vz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) vz |= v[i];
    if (vz == 0) break;
    if ((u[0] & 1u) == 0) { u256_shr1(u); u256_shl1(s); }
    else if ((v[0] & 1u) == 0) { u256_shr1(v); u256_shl1(r); }
    else {
      uint32_t d[8];
#pragma unroll
      for (int i = 0; i < 8; i++) d[i] = v[i];
      if (u256_sub(d, u) != 0) {          // v < u (both odd, so the difference is even)
        u256_sub(u, v);
        u256_shr1(u); u256_add(r, s); u256_shl1(s);
      } else {                            // v >= u; v = u = 1 at the last step, which ends the loop with v = 0
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = d[i];
        u256_shr1(v); u256_add(s, r); u256_shl1(r);
      }
    }
    k++;
  }
  Fp<P> x;
  {
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = r[i];
    fp_reduce_once<P>(t);               // r < 2p
    x.v[0] = sub_cc(P::mod(0), t[0]);    // x = p - r = (a R)^-1 2^k mod p, in [1, p]
#pragma unroll
    for (int i = 1; i < 7; i++) x.v[i] = subc_cc(P::mod(i), t[i]);
    x.v[7] = subc(P::mod(7), t[7]);
  }
  // x 2^(512 - k):  k > 256: mont(mont(x, R^2), 2^(512-k));  k <= 256: mont(mont(mont(x, R^2), R^2), 2^(256-k))
  x = fp_mul<P>(x, Fp<P>::r2());
  uint32_t e = 512u - k;
  if (e > 255u) { x = fp_mul<P>(x, Fp<P>::r2()); e -= 256u; }
  Fp<P> pw = Fp<P>::zero();
#pragma unroll
  for (int i = 0; i < 8; i++) pw.v[i] = ((e >> 5) == (uint32_t)i) ? (1u << (e & 31u)) : 0u;
  return fp_mul<P>(x, pw);
}

typedef Fp<FqParams> Fq;
typedef Fp<FrParams> Fr;

KB_HD Fq operator+(const Fq& a, const Fq& b) { return fp_add<FqParams>(a, b); }
KB_HD Fq operator-(const Fq& a, const Fq& b) { return fp_sub<FqParams>(a, b); }
KB_HD Fq operator*(const Fq& a, const Fq& b) { return fp_mul<FqParams>(a, b); }
KB_HD Fq operator-(const Fq& a) { return fp_neg<FqParams>(a); }
KB_HD Fq sqr(const Fq& a) { return fp_sqr<FqParams>(a); }
KB_HD Fq dbl(const Fq& a) { return fp_dbl<FqParams>(a); }
KB_HD Fq inv(const Fq& a) { return fp_inv<FqParams>(a); }

KB_HD Fr operator+(const Fr& a, const Fr& b) { return fp_add<FrParams>(a, b); }
KB_HD Fr operator-(const Fr& a, const Fr& b) { return fp_sub<FrParams>(a, b); }
KB_HD Fr operator*(const Fr& a, const Fr& b) { return fp_mul<FrParams>(a, b); }
KB_HD Fr operator-(const Fr& a) { return fp_neg<FrParams>(a); }
KB_HD Fr sqr(const Fr& a) { return fp_sqr<FrParams>(a); }
KB_HD Fr inv(const Fr& a) { return fp_inv<FrParams>(a); }

// 16-byte vectorised global-memory access (two 128-bit transactions per element)
#if defined(__CUDACC__)
template <class P>
KB_D Fp<P> fp_load(const uint32_t* p) {
  Fp<P> r;
  uint4 lo = *reinterpret_cast<const uint4*>(p), hi = *reinterpret_cast<const uint4*>(p + 4);
  r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
  r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
  return r;
}
template <class P>
KB_D void fp_store(uint32_t* p, const Fp<P>& a) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  *reinterpret_cast<uint4*>(p + 4) = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
#else
template <class P>
inline Fp<P> fp_load(const uint32_t* p) { Fp<P> r; for (int i = 0; i < 8; i++) r.v[i] = p[i]; return r; }
template <class P>
inline void fp_store(uint32_t* p, const Fp<P>& a) { for (int i = 0; i < 8; i++) p[i] = a.v[i]; }
#endif

}  // namespace kb
