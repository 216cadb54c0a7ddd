// Quad-cooperative G1 group law (shared by the MSM reduction tail, msm_acc.cu, and the small-domain G1 transforms, poly.cu).
#pragma once
#include "ctx.cuh"

namespace kb {

// ------------------------------------------------------------------------------------------
// Quad-cooperative group law for the dependent tails of the reduction.  Once only a few thousand points are left the
// MSM reduction is a chain of ~45 dependent group operations, and a lone warp needs ~0.43 us per field product however few
// of its lanes are busy.  Four adjacent lanes (a quad) therefore share one group operation: all four hold the same
// operands, each computes ONE of the independent products of a level of the formula, and the products are exchanged
// inside the quad by shuffles (quad-wide masks: quads of a warp may diverge from one another).  An addition is 4
// product levels instead of 14 products in sequence, a doubling 3 instead of 9.
// ------------------------------------------------------------------------------------------
struct QuadLane { uint32_t q, mask; };   // lane index inside the quad, shuffle mask of the quad
__device__ __forceinline__ QuadLane quad_lane() {
  QuadLane ql;
  const uint32_t lane = threadIdx.x & 31u;
  ql.q = lane & 3u;
  ql.mask = 0xFu << (lane & 28u);
  return ql;
}
__device__ __forceinline__ Fq quad_get(const Fq& v, int src, const QuadLane& ql) {   // v as held by lane `src` of the quad
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(ql.mask, v.v[i], src, 4);
  return r;
}
__device__ __forceinline__ Fq sel4(uint32_t q, const Fq& a0, const Fq& a1, const Fq& a2, const Fq& a3) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = q == 0 ? a0.v[i] : q == 1 ? a1.v[i] : q == 2 ? a2.v[i] : a3.v[i];
  return r;
}
static __device__ __noinline__ Fq quad_mul(Fq a, Fq b) { return fp_mul_inl<FqParams>(a, b); }

// add-2008-s over a quad; p and q (and the result) are replicated in the four lanes
static __device__ __noinline__ G1 g1_add4(G1 p, G1 q, QuadLane ql) {
  Fq m = quad_mul(sel4(ql.q, p.x, q.x, p.y, q.y), sel4(ql.q, q.zz, p.zz, q.zzz, p.zzz));
  const Fq u1 = quad_get(m, 0, ql), u2 = quad_get(m, 1, ql), s1 = quad_get(m, 2, ql), s2 = quad_get(m, 3, ql);
  const Fq pp_ = u2 - u1, r_ = s2 - s1;
  m = quad_mul(sel4(ql.q, pp_, r_, p.zz, p.zzz), sel4(ql.q, pp_, r_, q.zz, q.zzz));
  const Fq pp = quad_get(m, 0, ql), rr = quad_get(m, 1, ql), zz12 = quad_get(m, 2, ql), zzz12 = quad_get(m, 3, ql);
  m = quad_mul(sel4(ql.q, pp_, u1, zz12, pp_), pp);
  const Fq ppp = quad_get(m, 0, ql), qv = quad_get(m, 1, ql), zz3 = quad_get(m, 2, ql);
  G1 r;
  r.x = rr - ppp - dbl(qv);
  m = quad_mul(sel4(ql.q, r_, s1, zzz12, r_), sel4(ql.q, qv - r.x, ppp, ppp, ppp));
  r.y = quad_get(m, 0, ql) - quad_get(m, 1, ql);
  r.zz = zz3;
  r.zzz = quad_get(m, 2, ql);
  // exceptional cases: every lane decides alike (replicated operands); no shuffles below this line
  if (q.is_inf()) return p;
  if (p.is_inf()) return q;
  if (pp_.is_zero()) return r_.is_zero() ? ec_dbl(p) : G1::infinity();
  return r;
}
// dbl-2008-s-1 over a quad
static __device__ __noinline__ G1 g1_dbl4(G1 p, QuadLane ql) {
  const Fq u = dbl(p.y);
  Fq m = quad_mul(sel4(ql.q, u, p.x, u, p.x), sel4(ql.q, u, p.x, u, p.x));
  const Fq v = quad_get(m, 0, ql), xx = quad_get(m, 1, ql);
  const Fq mm_ = dbl(xx) + xx;
  m = quad_mul(sel4(ql.q, u, p.x, mm_, v), sel4(ql.q, v, v, mm_, p.zz));
  const Fq w = quad_get(m, 0, ql), sv = quad_get(m, 1, ql), msq = quad_get(m, 2, ql), zz3 = quad_get(m, 3, ql);
  G1 r;
  r.x = msq - dbl(sv);
  m = quad_mul(sel4(ql.q, mm_, w, w, w), sel4(ql.q, sv - r.x, p.y, p.zzz, p.y));
  r.y = quad_get(m, 0, ql) - quad_get(m, 1, ql);
  r.zz = zz3;
  r.zzz = quad_get(m, 2, ql);
  if (p.is_inf()) return p;
  return r;
}


// k * P over a quad (GLV, glv.cuh): p, k and the result replicated in the four lanes
static __device__ __noinline__ G1 g1_mul_glv4(G1 p, const uint32_t* k, QuadLane ql) {
  if (p.is_inf()) return p;
  const GlvSplit s = glv_decompose(k);
  Fq beta;
#pragma unroll
  for (int i = 0; i < 8; i++) beta.v[i] = GlvParams::beta(i);
  G1 tab[16];
  tab[0] = G1::infinity();
  tab[1] = p;
  if (s.neg1) tab[1].y = -tab[1].y;
  for (int i = 2; i < 16; i++) tab[i] = (i & 1) ? g1_add4(tab[i - 1], tab[1], ql) : g1_dbl4(tab[i >> 1], ql);
  const bool flip = s.neg1 != s.neg2;
  G1 acc = G1::infinity();
  for (int w = 32; w >= 0; w--) {
    if (!acc.is_inf()) { acc = g1_dbl4(acc, ql); acc = g1_dbl4(acc, ql); acc = g1_dbl4(acc, ql); acc = g1_dbl4(acc, ql); }
    const uint32_t d1 = (s.m1[w >> 3] >> ((w & 7) * 4)) & 15u;
    const uint32_t d2 = (s.m2[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d1) acc = g1_add4(acc, tab[d1], ql);
    if (d2) {
      G1 t = tab[d2];
      t.x = quad_mul(t.x, beta);
      if (flip) t.y = -t.y;
      acc = g1_add4(acc, t, ql);
    }
  }
  return acc;
}

}  // namespace kb
