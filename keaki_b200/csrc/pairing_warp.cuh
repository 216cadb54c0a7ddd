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
//                                       (one IMAD.WIDE row + one carry chain per term and coordinate), reduced ONCE
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
static constexpr int MAX_TERMS = 14;

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

KB_ST_INL uint32_t term_of(const uint32_t* d, int k) { return (d[1 + (k >> 1)] >> (16 * (k & 1))) & 0xffffu; }

template <class M>
KB_ST_INL Fq2 lin(const M& m, const uint32_t* d, uint32_t nu, uint32_t nw) {
  uint32_t u0[9], u1[9];
#pragma unroll
  for (int i = 0; i < 9; i++) u0[i] = u1[i] = 0;
#pragma unroll
  for (int k = 0; k < MAX_TERMS; k++) {
    if ((uint32_t)k < nu) {
      const uint32_t tm = term_of(d, k);
      const Fq2 x = m.ld(tm & 511u);
      const uint32_t c = ((tm >> 10) & 63u) + 1u, neg = 0u - ((tm >> 9) & 1u);
      mac9(u0, x.c0, c, neg);
      mac9(u1, x.c1, c, neg);
    }
  }
  if (nw) {
    uint32_t w0[9], w1[9];
#pragma unroll
    for (int i = 0; i < 9; i++) w0[i] = w1[i] = 0;
#pragma unroll
    for (int k = 0; k < MAX_TERMS; k++) {
      if ((uint32_t)k < nw) {
        const uint32_t tm = term_of(d, MAX_TERMS - 1 - k);
        const Fq2 x = m.ld(tm & 511u);
        const uint32_t c = ((tm >> 10) & 63u) + 1u, neg = 0u - ((tm >> 9) & 1u);
        mac9(w0, x.c0, c, neg);
        mac9(w1, x.c1, c, neg);
      }
    }
    // (9 + u)(w0 + w1 u) = (9 w0 - w1) + (9 w1 + w0) u
    mad9x9(u0, w0); sub9(u0, w1);
    mad9x9(u1, w1); add9(u1, w0);
  }
  Fq2 r;
  r.c0 = reduce9(u0);
  r.c1 = reduce9(u1);
  return r;
}

// One lane's share of a step: the value it will store (`dst`, when `active`).  Reads only.
template <class M>
KB_ST_INL Fq2 lane_compute(const M& m, const uint32_t* d) {
  const uint32_t kind = d[0] & 15u, active = (d[0] >> 4) & 1u;
  Fq2 r = Fq2::zero();
  if (kind == K_MUL) {
    const Fq2 a = st::add_loose(m.ld(d[1] & 0xffffu), m.ld(d[1] >> 16));
    const Fq2 b = m.ld(d[2] & 0xffffu) + m.ld(d[2] >> 16);
    r = st::f2mul(a, b);
  } else if (kind == K_LIN) {
    r = lin(m, d, (d[0] >> 20) & 15u, (d[0] >> 24) & 15u);
  } else if (kind == K_CONJ) {
    r = m.ld(d[1] & 0xffffu);
    r.c1 = -r.c1;
  } else if (kind == K_INV) {
    if (active) r = st::f2inv(m.ld(d[1] & 0xffffu));
  }
  return r;
}

}  // namespace wp
}  // namespace kb
