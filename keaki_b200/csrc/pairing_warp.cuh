// Warp-cooperative pairing: ONE WARP per pairing, the 32 lanes executing independent Fq2 operations of that pairing
// side by side.  The low-latency form of `E::pairing` (ark-ec 0.4.2 models/bn/mod.rs; src/kem.rs:30,58, src/kzg.rs:148):
// a lone thread of the throughput kernel (pairing_st) needs about 9 ms for a pairing, however small the batch; here the
// dependent chain is ~660 Fq2-product levels deep instead of ~6000 products long.  Used for small batches, for the one
// pairing per fresh commitment of kb_encrypt_batch and for the 248 chained cyclotomic squarings of its window bases.
//
// tools/gen_pairing_warp.py traces the tower formulas into a DAG of Fq2 nodes, keeps linear combinations lazy so that
// exactly one combination level separates two product levels, and list-schedules the DAG into STEPS of one kind:
//   MUL   d = (a1 + a2) (b1 + b2)       lazily reduced Fq2 product (fpl.cuh); the a-side sum is not reduced
//   LIN   d = sum c_i x_i + xi sum c'_j y_j   small signed integer coefficients: accumulated as 288-bit integers
//                                       (one IMAD.WIDE row + one carry chain per term), reduced ONCE; two lanes per
//                                       node, one per coordinate of the Fq2 result
//   CONJ  d = conj(x)        INV  d = 1 / x   (once per pairing, one lane)
// Every lane reads its 32-byte descriptor (prefetched one step ahead), computes from the warp's slot file in shared
// memory, and after a warp barrier stores its result, so results may reuse slots whose last reader is in the same step.
//
// Compiles for the host too (tests/hostemu runs the shipped schedules lane by lane, TEST ONLY).
#pragma once
#include "pairing_st.cuh"

namespace kb {
namespace wp {

enum : uint32_t { K_NOP = 0, K_MUL = 1, K_LIN = 2, K_CONJ = 3, K_INV = 4 };

// acc (288-bit two's complement) += (m ? -1 : 1) * c * x;  m = 0 or 0xffffffff
KB_ST_INL void mac9(uint32_t* acc, const Fq& x, uint32_t c, uint32_t m) {
  uint32_t y[9];
  uint64_t t = (uint64_t)x.v[0] * c;
  y[0] = (uint32_t)t ^ m;
#pragma unroll
  for (int i = 1; i < 8; i++) { t = (uint64_t)x.v[i] * c + (t >> 32); y[i] = (uint32_t)t ^ m; }
  y[8] = (uint32_t)(t >> 32) ^ m;
  (void)add_cc(m, m);                    // carry = 1 when negating: -v = ~v + 1
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = addc_cc(acc[i], y[i]);
  acc[8] = addc(acc[8], y[8]);
}
KB_ST_INL void add9(uint32_t* a, const uint32_t* b) {
  a[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) a[i] = addc_cc(a[i], b[i]);
  a[8] = addc(a[8], b[8]);
}
KB_ST_INL void sub9(uint32_t* a, const uint32_t* b) {
  a[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) a[i] = subc_cc(a[i], b[i]);
  a[8] = subc(a[8], b[8]);
}
// a += 9 b (mod 2^288: two's complement values multiply consistently)
KB_ST_INL void mad9x9(uint32_t* a, const uint32_t* b) {
  uint32_t y[9];
  uint64_t t = (uint64_t)b[0] * 9u;
  y[0] = (uint32_t)t;
#pragma unroll
  for (int i = 1; i < 9; i++) { t = (uint64_t)b[i] * 9u + (t >> 32); y[i] = (uint32_t)t; }
  add9(a, y);
}
// v in (-128 q, 128 q) as a 288-bit two's complement integer  ->  v mod q.  128 q is added first; the quotient of the
// positive value V < 256 q is estimated from its bits 232.. against floor(q / 2^232) + 1 = 3171407 (never too large, at
// most 2 too small: checked at the extremes in tests/test_pairing_warp.py), then two conditional subtractions.
KB_ST_INL Fq reduce9(uint32_t* v) {
  uint32_t off[9];
  off[0] = FqParams::mod(0) << 7;
#pragma unroll
  for (int i = 1; i < 8; i++) off[i] = (FqParams::mod(i) << 7) | (FqParams::mod(i - 1) >> 25);
  off[8] = FqParams::mod(7) >> 25;
  add9(v, off);
  const uint32_t vt = (v[8] << 24) | (v[7] >> 8);
  const uint32_t k = mul_hi(vt, 2840127191u) >> 21;     // floor(2^53 / 3171407) = 2840127191
  uint32_t kq[8];
  uint64_t t = (uint64_t)FqParams::mod(0) * k;
  kq[0] = (uint32_t)t;
#pragma unroll
  for (int i = 1; i < 8; i++) { t = (uint64_t)FqParams::mod(i) * k + (t >> 32); kq[i] = (uint32_t)t; }
  Fq r;
  r.v[0] = sub_cc(v[0], kq[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = subc_cc(v[i], kq[i]);
  r.v[7] = subc(v[7], kq[7]);            // V - k q < 3 q < 2^256: the ninth limb cancels
  fp_reduce_once<FqParams>(r.v);
  fp_reduce_once<FqParams>(r.v);
  return r;
}

// One coordinate of a LIN node: lane `half` computes c_half of  sum c_i x_i + xi sum c'_j y_j, where
// xi (y0 + y1 u) = (9 y0 - y1) + (9 y1 + y0) u.  The 14 16-bit terms sit in a 7-register queue that is shifted as they
// are consumed, so that the (step-uniform) term counts drive two small loops instead of 28 unrolled bodies: the warp is
// alone on its scheduler, and instruction fetch of a long straight-line body was a fifth of its stalls.
template <class M>
KB_ST_INL Fq lin_half(const M& m, const uint32_t* d, uint32_t nu, uint32_t nw, uint32_t half) {
  uint32_t acc[9];
#pragma unroll
  for (int i = 0; i < 9; i++) acc[i] = 0;
  uint32_t q[7];
#pragma unroll
  for (int i = 0; i < 7; i++) q[i] = d[1 + i];
#define KB_WP_NEXT_TERM(tm)                                                           \
  const uint32_t tm = q[0] & 0xffffu;                                                 \
  _Pragma("unroll") for (int i = 0; i < 6; i++) q[i] = (q[i] >> 16) | (q[i + 1] << 16); \
  q[6] >>= 16;
#pragma unroll 1
  for (uint32_t k = 0; k < nu; k++) {
    KB_WP_NEXT_TERM(tm)
    mac9(acc, m.ld_half(tm & 511u, half), ((tm >> 10) & 63u) + 1u, 0u - ((tm >> 9) & 1u));
  }
  const uint32_t flip = half ? 0u : 0xffffffffu;   // the cross term enters c0 with a minus sign
#pragma unroll 1
  for (uint32_t k = 0; k < nw; k++) {
    KB_WP_NEXT_TERM(tm)
    const uint32_t c = ((tm >> 10) & 63u) + 1u, neg = 0u - ((tm >> 9) & 1u);
    mac9(acc, m.ld_half(tm & 511u, half), 9u * c, neg);
    mac9(acc, m.ld_half(tm & 511u, half ^ 1u), c, neg ^ flip);
  }
#undef KB_WP_NEXT_TERM
  return reduce9(acc);
}

// One lane's share of a step: the value it will store (lane_store).  Reads only.
template <class M>
KB_ST_INL Fq2 lane_compute(const M& m, const uint32_t* d) {
  const uint32_t kind = d[0] & 15u, active = (d[0] >> 4) & 1u;
  Fq2 r = Fq2::zero();
  if (kind == K_MUL) {
    const Fq2 a = st::add_loose(m.ld(d[1] & 0xffffu), m.ld(d[1] >> 16));
    const Fq2 b = m.ld(d[2] & 0xffffu) + m.ld(d[2] >> 16);
    r = st::f2mul(a, b);
  } else if (kind == K_LIN) {
    r.c0 = lin_half(m, d, (d[0] >> 20) & 15u, (d[0] >> 24) & 15u, (d[0] >> 17) & 1u);
  } else if (kind == K_CONJ) {
    r = m.ld(d[1] & 0xffffu);
    r.c1 = -r.c1;
  } else if (kind == K_INV) {
    if (active) r = st::f2inv(m.ld(d[1] & 0xffffu));
  }
  return r;
}
// ... and the store, after every lane of the step has computed
template <class M>
KB_ST_INL void lane_store(const M& m, const uint32_t* d, const Fq2& r) {
  if (!((d[0] >> 4) & 1u)) return;
  const uint32_t dst = (d[0] >> 8) & 511u;
  if ((d[0] & 15u) == K_LIN) m.st_half(dst, (d[0] >> 17) & 1u, r.c0);
  else m.st(dst, r);
}

}  // namespace wp
}  // namespace kb
