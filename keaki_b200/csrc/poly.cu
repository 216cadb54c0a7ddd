// Polynomial layer: radix-2 NTT over Fr (src/vec.rs:37, src/kzg.rs:185), quotient polynomials for
// `open` (src/kzg.rs:104-124) and the FK23 all-openings algorithm `open_fk` (src/kzg.rs:157-203).
//
// open_fk data flow here (same mathematics as the reference, fewer group operations):
//   hat_s = DFT_2d(s), s = reversed SRS prefix || d x infinity, depends only on (SRS, d): computed
//           once and cached in the context (the reference recomputes it on every call, :182);
//   hat_a = DFT_2d(0^d || p) over Fr; hat_h[i] = (hat_a[i] / 2d) * hat_s[i]   (:185-191, iDFT scale folded in);
//   h = first d entries of DFT^-1_2d(hat_h) (unscaled); proofs = DFT_d(h)            (:194-200).
// G1 transforms use XYZZ points and scalar-multiplication butterflies, skipping unit twiddles.
#include "ctx.cuh"
#include "quad.cuh"
#include "consts_gen.cuh"

namespace kb {

static int log2_exact(uint64_t n) {
  if (n == 0 || (n & (n - 1))) return -1;
  int k = 0;
  while ((1ull << k) < n) k++;
  return k;
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) { return bits ? __brev(x) >> (32 - bits) : 0; }

// tw[j] = w^j for j < n/2 (w = primitive n-th root, or its inverse), Montgomery form;
// tw[n/2] = n^{-1}.  `canon` != null also receives from_mont(tw[j]).
__global__ void __launch_bounds__(256) fr_twiddle_kernel(const uint32_t* __restrict__ root28, const uint32_t* __restrict__ two_inv,
                                                         int logn, uint32_t* __restrict__ tw, uint32_t* __restrict__ canon) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t half = logn ? (1u << (logn - 1)) : 0u;
  if (j > half) return;
  if (j == half) {
    Fr ti = fp_load<FrParams>(two_inv), acc = Fr::one();
    for (int k = 0; k < logn; k++) acc = acc * ti;
    fp_store<FrParams>(tw + 8 * (size_t)half, acc);
    return;
  }
  Fr w = fp_load<FrParams>(root28);
  for (int k = logn; k < 28; k++) w = sqr(w);
  Fr acc = Fr::one();
  for (uint32_t e = j; e; e >>= 1) { if (e & 1u) acc = acc * w; w = sqr(w); }
  fp_store<FrParams>(tw + 8 * (size_t)j, acc);
  if (canon) fp_store<FrParams>(canon + 8 * (size_t)j, fp_from_mont<FrParams>(acc));
}

__global__ void __launch_bounds__(256) fr_bitrev_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int logn) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << logn)) return;
  fp_store<FrParams>(out + 8 * (size_t)bitrev(i, logn), fp_load<FrParams>(in + 8 * (size_t)i));
}

__global__ void __launch_bounds__(256) fr_butterfly_kernel(uint32_t* __restrict__ a, const uint32_t* __restrict__ tw, int logn, int s) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t half_n = 1u << (logn - 1);
  if (t >= half_n) return;
  uint32_t half = 1u << s;
  uint32_t k = t & (half - 1), i = ((t >> s) << (s + 1)) + k, j = i + half;
  Fr u = fp_load<FrParams>(a + 8 * (size_t)i);
  Fr v = fp_load<FrParams>(a + 8 * (size_t)j) * fp_load<FrParams>(tw + 8 * (size_t)(k << (logn - 1 - s)));
  fp_store<FrParams>(a + 8 * (size_t)i, u + v);
  fp_store<FrParams>(a + 8 * (size_t)j, u - v);
}

__global__ void __launch_bounds__(256) fr_scale_kernel(uint32_t* __restrict__ a, const uint32_t* __restrict__ k, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fp_store<FrParams>(a + 8 * (size_t)i, fp_load<FrParams>(a + 8 * (size_t)i) * fp_load<FrParams>(k));
}

struct Twiddles {
  DevBuf<uint32_t> tw, canon;
  Twiddles(kb_ctx* ctx, int logn, bool inverse, bool want_canon)
      : tw(ctx, 8 * ((size_t)(logn ? 1ull << (logn - 1) : 0) + 1)), canon(ctx, want_canon ? 8 * ((size_t)(logn ? 1ull << (logn - 1) : 0) + 1) : 0) {
    DevBuf<uint32_t> c(ctx, 16);
    KB_CUDA(cudaMemcpyAsync(c, inverse ? consts::FR_ROOT_2_28_INV : consts::FR_ROOT_2_28, 32, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(cudaMemcpyAsync(c.p + 8, consts::FR_TWO_INV, 32, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t half = logn ? (1u << (logn - 1)) : 0u;
    KB_LAUNCH(ctx, fr_twiddle_kernel, cdiv(half + 1, 256), 256, 0, c.p, c.p + 8, logn, tw.p, want_canon ? canon.p : nullptr);
  }
};

// in-place on d_data (natural order in and out); scale = apply 1/n after an inverse transform
static void fr_ntt_dev(kb_ctx* ctx, uint32_t* d_data, int logn, bool inverse, bool scale) {
  uint64_t n = 1ull << logn;
  if (logn == 0) return;
  Twiddles tw(ctx, logn, inverse, false);
  DevBuf<uint32_t> tmp(ctx, 8 * n);
  KB_LAUNCH(ctx, fr_bitrev_kernel, cdiv(n, 256), 256, 0, d_data, tmp.p, logn);
  for (int s = 0; s < logn; s++) KB_LAUNCH(ctx, fr_butterfly_kernel, cdiv(n / 2, 256), 256, 0, tmp.p, tw.tw.p, logn, s);
  KB_CUDA(cudaMemcpyAsync(d_data, tmp.p, 32 * n, cudaMemcpyDeviceToDevice, ctx->stream));
  if (inverse && scale) KB_LAUNCH(ctx, fr_scale_kernel, cdiv(n, 256), 256, 0, d_data, tw.tw.p + 8 * (n / 2), (uint32_t)n);
}

void fr_ntt(kb_ctx* ctx, uint32_t* d_data, uint64_t n, bool inverse) {
  int logn = log2_exact(n);
  if (logn < 0 || logn > 28) throw ApiError(KB_ERR_DOMAIN, "kb_fr_ntt: size must be a power of two <= 2^28");
  fr_ntt_dev(ctx, d_data, logn, inverse, true);
}

// ------------------------------------------------------------------------------------------
// quotient (p(x) - p(z)) / (x - z): chunked Horner, q_{i-1} = sum_{j >= i} p_j z^(j-i)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) quot_chunk_eval_kernel(const uint32_t* __restrict__ p, uint64_t d, const uint32_t* __restrict__ z,
                                                              uint32_t L, uint32_t T, uint32_t* __restrict__ S, uint32_t* __restrict__ ZL) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint64_t lo = (uint64_t)t * L, hi = lo + L < d ? lo + L : d;
  Fr zz = fp_load<FrParams>(z), acc = Fr::zero(), zp = Fr::one();
  for (uint64_t i = hi; i-- > lo;) { acc = acc * zz + fp_load<FrParams>(p + 8 * i); zp = zp * zz; }
  fp_store<FrParams>(S + 8 * (size_t)t, acc);       // sum_{j in chunk} p_j z^(j - lo)
  fp_store<FrParams>(ZL + 8 * (size_t)t, zp);       // z^(chunk length)
}
// H[t] = S[t] + z^len_t * H[t+1], H[T] = 0   (suffix Horner values at chunk starts)
__global__ void quot_chain_kernel(const uint32_t* __restrict__ S, const uint32_t* __restrict__ ZL, uint32_t T, uint32_t* __restrict__ H) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Fr acc = Fr::zero();
  fp_store<FrParams>(H + 8 * (size_t)T, acc);
  for (uint32_t t = T; t-- > 0;) {
    acc = fp_load<FrParams>(S + 8 * (size_t)t) + fp_load<FrParams>(ZL + 8 * (size_t)t) * acc;
    fp_store<FrParams>(H + 8 * (size_t)t, acc);
  }
}
__global__ void __launch_bounds__(128) quot_write_kernel(const uint32_t* __restrict__ p, uint64_t d, const uint32_t* __restrict__ z,
                                                         uint32_t L, uint32_t T, const uint32_t* __restrict__ H, uint32_t* __restrict__ q) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint64_t lo = (uint64_t)t * L, hi = lo + L < d ? lo + L : d;
  Fr zz = fp_load<FrParams>(z), acc = fp_load<FrParams>(H + 8 * (size_t)(t + 1));
  for (uint64_t i = hi; i-- > lo;) {
    acc = acc * zz + fp_load<FrParams>(p + 8 * i);
    if (i >= 1) fp_store<FrParams>(q + 8 * (i - 1), acc);
  }
}

void open_batch(kb_ctx* ctx, const uint32_t* d_coeffs, uint64_t d, const uint32_t* d_points, uint64_t m,
                uint32_t* d_proofs, uint8_t* d_inf) {
  if (d >= 1 && d - 1 > ctx->srs_n)
    throw ApiError(KB_ERR_POLY_TOO_LARGE, "PolynomialTooLarge(" + std::to_string(d - 1) + ", " + std::to_string(ctx->srs_n) + ")");
  if (d <= 1) {  // constant polynomial: quotient is zero, proof is the identity
    if (m) { KB_CUDA(cudaMemsetAsync(d_proofs, 0, 64 * m, ctx->stream)); KB_CUDA(cudaMemsetAsync(d_inf, 1, m, ctx->stream)); }
    return;
  }
  uint32_t L = 16;
  while ((uint64_t)L * L < d) L <<= 1;
  uint32_t T = cdiv(d, L);
  DevBuf<uint32_t> S(ctx, 8 * (size_t)T), ZL(ctx, 8 * (size_t)T), H(ctx, 8 * ((size_t)T + 1)), q(ctx, 8 * (d - 1));
  for (uint64_t j = 0; j < m; j++) {
    const uint32_t* z = d_points + 8 * j;
    KB_LAUNCH(ctx, quot_chunk_eval_kernel, cdiv(T, 128), 128, 0, d_coeffs, d, z, L, T, S.p, ZL.p);
    KB_LAUNCH(ctx, quot_chain_kernel, 1, 32, 0, S.p, ZL.p, T, H.p);
    KB_LAUNCH(ctx, quot_write_kernel, cdiv(T, 128), 128, 0, d_coeffs, d, z, L, T, H.p, q.p);
    msm_g1(ctx, q.p, 0, d - 1, d_proofs + 16 * j, d_inf + j);
  }
}

// ------------------------------------------------------------------------------------------
// G1 NTT (XYZZ points, 32 limbs each)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) g1_bitrev_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int logn) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << logn)) return;
  st_g1x(out + 32 * (size_t)bitrev(i, logn), ld_g1x(in + 32 * (size_t)i));
}

__global__ void __launch_bounds__(128) g1_butterfly_kernel(uint32_t* __restrict__ a, const uint32_t* __restrict__ tw_canon, int logn, int s) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t half_n = 1u << (logn - 1);
  if (t >= half_n) return;
  uint32_t half = 1u << s;
  uint32_t k = t & (half - 1), i = ((t >> s) << (s + 1)) + k, j = i + half;
  G1 u = ld_g1x(a + 32 * (size_t)i);
  G1 v = ld_g1x(a + 32 * (size_t)j);
  if (k != 0 && !v.is_inf()) {
    uint32_t kk[8];
    const uint32_t* src = tw_canon + 8 * (size_t)(k << (logn - 1 - s));
    for (int q = 0; q < 8; q++) kk[q] = src[q];
    v = g1_mul_glv(v, kk);
  }
  st_g1x(a + 32 * (size_t)i, ec_add(u, v));
  st_g1x(a + 32 * (size_t)j, ec_add(u, neg(v)));
}

// The same butterfly with four lanes per butterfly (quad.cuh).  Small domains are latency-bound: a stage of a few
// thousand butterflies waits for ONE scalar multiplication per thread (about 2,200 dependent products), so the quad's
// shorter dependent chain matters and its extra lanes cost nothing.
__global__ void __launch_bounds__(128) g1_butterfly4_kernel(uint32_t* __restrict__ a, const uint32_t* __restrict__ tw_canon, int logn, int s) {
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const QuadLane ql = quad_lane();
  const uint32_t half_n = 1u << (logn - 1);
  if (t >= half_n) return;   // whole quads leave together
  const uint32_t half = 1u << s;
  const uint32_t k = t & (half - 1), i = ((t >> s) << (s + 1)) + k, j = i + half;
  G1 u = ld_g1x(a + 32 * (size_t)i);
  G1 v = ld_g1x(a + 32 * (size_t)j);
  if (k != 0 && !v.is_inf()) {
    uint32_t kk[8];
    const uint32_t* src = tw_canon + 8 * (size_t)(k << (logn - 1 - s));
    for (int q = 0; q < 8; q++) kk[q] = src[q];
    v = g1_mul_glv4(v, kk, ql);
  }
  const G1 r0 = g1_add4(u, v, ql), r1 = g1_add4(u, neg(v), ql);
  if (ql.q == 0) { st_g1x(a + 32 * (size_t)i, r0); st_g1x(a + 32 * (size_t)j, r1); }
}

// R merged stages (s .. s+R-1) in ONE pass for the latency-bound sizes.  A stage of a small transform waits for one scalar
// multiplication per butterfly, and the stages are dependent: 13 multiplications in sequence for 2^13 points.  The
// transform is linear, so the 2^R outputs of a radix-2^R butterfly are signed sums of (twiddle product) x (input) terms
// that can all be computed at once:
//     y_f = a_0 + sum_{e = 1}^{2^R - 1} (-1)^popcount(e & f) w^E(e, f mod 2^h(e)) a_e,     h(e) = index of the top bit of e,
//     E(e, v) = sum_{t : bit t of e set} (k + (v mod 2^t) 2^s) 2^(L - 1 - s - t)
// (the path of input e through the R radix-2 stages picks up stage t's twiddle exactly when bit t of e is set, and by then
// the low t bits of its position are those of the output f).  Input e needs 2^h(e) distinct products: 1 + 4 + 16 = 21
// scalar multiplications per radix-8 butterfly instead of 12 - but one multiplication deep instead of three.
__device__ __forceinline__ int r8_job(int e, int v) { return e == 1 ? 0 : e < 4 ? 1 + (e - 2) * 2 + v : 5 + (e - 4) * 4 + v; }
template <int R> struct R8Jobs { static constexpr int N = R == 1 ? 1 : R == 2 ? 5 : 21; };

// QUAD: four lanes share one multiplication (shorter dependent chain, four times the lanes); worth it only while the
// lanes are free - a radix-8 pass over 2^13 points has 21.5 K multiplications, and with one THREAD each they still run at
// lone-warp speed (a warp or two per scheduler) while the quad form is throughput-bound (measured 2.05 ms against 0.6)
template <int R, bool QUAD>
__global__ void __launch_bounds__(128) g1_radix_mul_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ tw_canon, int logn, int s,
                                                           uint32_t* __restrict__ T) {
  constexpr int NJ = R8Jobs<R>::N;
  const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t qi = QUAD ? gt >> 2 : gt;
  const QuadLane ql = quad_lane();
  const uint32_t groups = 1u << (logn - R);
  if (qi >= groups * NJ) return;   // whole quads leave together
  const uint32_t group = qi / NJ, jidx = qi % NJ;
  int e, v;
  if (jidx == 0) { e = 1; v = 0; }
  else if (jidx < 5) { e = 2 + (jidx - 1) / 2; v = (jidx - 1) & 1; }
  else { e = 4 + (jidx - 5) / 4; v = (jidx - 5) & 3; }
  const uint32_t k = group & ((1u << s) - 1), base = ((group >> s) << (s + R)) + k;
  uint32_t E = 0;
#pragma unroll
  for (int t = 0; t < R; t++)
    if ((e >> t) & 1) E += (k + ((uint32_t)(v & ((1 << t) - 1)) << s)) << (logn - 1 - s - t);
  E &= (1u << logn) - 1;
  G1 p = ld_g1x(a + 32 * (size_t)(base + ((uint32_t)e << s)));
  const uint32_t idx = E & ((1u << (logn - 1)) - 1);
  if (idx != 0 && !p.is_inf()) {
    uint32_t kk[8];
    const uint32_t* src = tw_canon + 8 * (size_t)idx;
    for (int q = 0; q < 8; q++) kk[q] = src[q];
    if (QUAD) p = g1_mul_glv4(p, kk, ql);
    else p = g1_mul_glv(p, kk);
  }
  if (E >> (logn - 1)) p = neg(p);      // w^(n/2) = -1
  if (!QUAD || ql.q == 0) st_g1x(T + 32 * (size_t)qi, p);
}

template <int R>
__global__ void __launch_bounds__(128) g1_radix_sum_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ T, int logn, int s,
                                                           uint32_t* __restrict__ out) {
  constexpr int NJ = R8Jobs<R>::N;
  const uint32_t qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const QuadLane ql = quad_lane();
  if (qi >= (1u << logn)) return;
  const uint32_t group = qi >> R, f = qi & ((1u << R) - 1);
  const uint32_t k = group & ((1u << s) - 1), base = ((group >> s) << (s + R)) + k;
  G1 acc = ld_g1x(a + 32 * (size_t)base);
#pragma unroll 1
  for (int e = 1; e < (1 << R); e++) {
    const int h = e >= 4 ? 2 : e >= 2 ? 1 : 0;
    G1 t = ld_g1x(T + 32 * ((size_t)group * NJ + r8_job(e, f & ((1 << h) - 1))));
    if (__popc(e & f) & 1) t = neg(t);
    acc = g1_add4(acc, t, ql);
  }
  if (ql.q == 0) st_g1x(out + 32 * (size_t)(base + (f << s)), acc);
}

template <int R>
static void g1_radix_pass(kb_ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t* T, const uint32_t* tw_canon, int logn, int s) {
  const uint64_t jobs = (uint64_t)R8Jobs<R>::N << (logn - R);
  if (jobs <= 4096) KB_LAUNCH(ctx, (g1_radix_mul_kernel<R, true>), cdiv(4 * jobs, 128), 128, 0, in, tw_canon, logn, s, T);
  else KB_LAUNCH(ctx, (g1_radix_mul_kernel<R, false>), cdiv(jobs, 128), 128, 0, in, tw_canon, logn, s, T);
  KB_LAUNCH(ctx, (g1_radix_sum_kernel<R>), cdiv(4ull << logn, 128), 128, 0, in, T, logn, s, out);
}

// in-place natural-order G1 transform on d_pts (XYZZ), unscaled
static void g1_ntt_dev(kb_ctx* ctx, uint32_t* d_pts, int logn, bool inverse) {
  if (logn == 0) return;
  uint64_t n = 1ull << logn;
  Twiddles tw(ctx, logn, inverse, true);
  DevBuf<uint32_t> tmp(ctx, 32 * n);
  KB_LAUNCH(ctx, g1_bitrev_kernel, cdiv(n, 256), 256, 0, d_pts, tmp.p, logn);
  if (n <= (1ull << 14) && !ctx->ntt_radix2) {   // latency-bound sizes: merged stages, ping-pong between tmp and d_pts
    DevBuf<uint32_t> T(ctx, 32 * ((uint64_t)21 << (logn >= 3 ? logn - 3 : 0)));
    uint32_t* cur = tmp.p;
    uint32_t* nxt = d_pts;
    int s = 0;
    while (s < logn) {
      const int r = logn - s >= 3 ? 3 : logn - s;
      if (r == 3) g1_radix_pass<3>(ctx, cur, nxt, T.p, tw.canon.p, logn, s);
      else if (r == 2) g1_radix_pass<2>(ctx, cur, nxt, T.p, tw.canon.p, logn, s);
      else g1_radix_pass<1>(ctx, cur, nxt, T.p, tw.canon.p, logn, s);
      s += r;
      uint32_t* x = cur; cur = nxt; nxt = x;
    }
    if (cur != d_pts) KB_CUDA(cudaMemcpyAsync(d_pts, cur, 128 * n, cudaMemcpyDeviceToDevice, ctx->stream));
    return;
  }
  const bool quads = n <= (1ull << 14);   // latency-bound sizes
  for (int s = 0; s < logn; s++) {
    if (quads) KB_LAUNCH(ctx, g1_butterfly4_kernel, cdiv(2 * n, 128), 128, 0, tmp.p, tw.canon.p, logn, s);
    else KB_LAUNCH(ctx, g1_butterfly_kernel, cdiv(n / 2, 128), 128, 0, tmp.p, tw.canon.p, logn, s);
  }
  KB_CUDA(cudaMemcpyAsync(d_pts, tmp.p, 128 * n, cudaMemcpyDeviceToDevice, ctx->stream));
}

// s[i] = srs[d-1-i] for i < d, infinity for d <= i < 2d   (src/kzg.rs:167-174)
__global__ void __launch_bounds__(256) fk_build_s_kernel(const uint32_t* __restrict__ srs, uint32_t d, uint32_t* __restrict__ s) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * d) return;
  G1 p = i < d ? to_xyzz(ld_g1(srs + 16 * (size_t)(d - 1 - i))) : G1::infinity();
  st_g1x(s + 32 * (size_t)i, p);
}
// a = 0^d || p   (src/kzg.rs:178-179)
__global__ void __launch_bounds__(256) fk_build_a_kernel(const uint32_t* __restrict__ p, uint32_t d, uint32_t* __restrict__ a) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * d) return;
  Fr v = i < d ? Fr::zero() : fp_load<FrParams>(p + 8 * (size_t)(i - d));
  fp_store<FrParams>(a + 8 * (size_t)i, v);
}
// hat_h[i] = (hat_a[i] * inv2d) * hat_s[i]   (src/kzg.rs:188-191)
__global__ void __launch_bounds__(128) fk_pointwise_kernel(const uint32_t* __restrict__ hat_s, const uint32_t* __restrict__ hat_a,
                                                           const uint32_t* __restrict__ inv2d, uint32_t n2, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  Fr k = fp_from_mont<FrParams>(fp_load<FrParams>(hat_a + 8 * (size_t)i) * fp_load<FrParams>(inv2d));
  st_g1x(out + 32 * (size_t)i, g1_mul_glv(ld_g1x(hat_s + 32 * (size_t)i), k.v));
}
__global__ void __launch_bounds__(128) fk_pointwise4_kernel(const uint32_t* __restrict__ hat_s, const uint32_t* __restrict__ hat_a,
                                                            const uint32_t* __restrict__ inv2d, uint32_t n2, uint32_t* __restrict__ out) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const QuadLane ql = quad_lane();
  if (i >= n2) return;
  Fr k = fp_from_mont<FrParams>(fp_load<FrParams>(hat_a + 8 * (size_t)i) * fp_load<FrParams>(inv2d));
  const G1 r = g1_mul_glv4(ld_g1x(hat_s + 32 * (size_t)i), k.v, ql);
  if (ql.q == 0) st_g1x(out + 32 * (size_t)i, r);
}
__global__ void __launch_bounds__(128) g1_xyzz_to_affine_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out_xy,
                                                                uint8_t* __restrict__ out_inf) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1 p = ld_g1x(in + 32 * (size_t)i);
  st_g1(out_xy + 16 * (size_t)i, to_affine(p));
  out_inf[i] = p.is_inf() ? 1 : 0;
}

void fk_free(kb_ctx* ctx) {
  for (auto& c : ctx->fk_cache) if (c.d_hat_s) cudaFree(c.d_hat_s);
  ctx->fk_cache.clear();
}

void open_all_fk(kb_ctx* ctx, const uint32_t* d_coeffs, uint64_t d, uint32_t* d_proofs, uint8_t* d_inf) {
  int logd = log2_exact(d);
  if (logd < 0 || logd + 1 > 28) throw ApiError(KB_ERR_DOMAIN, "kb_open_all_fk: d must be a power of two with 2d <= 2^28");
  if (d > ctx->srs_n) throw ApiError(KB_ERR_POLY_TOO_LARGE, "open_fk: d = " + std::to_string(d) + " exceeds SRS length " + std::to_string(ctx->srs_n));
  const uint32_t n2 = (uint32_t)(2 * d);
  // hat_s: cached per d
  uint32_t* hat_s = nullptr;
  for (auto& c : ctx->fk_cache) if (c.d == d) hat_s = c.d_hat_s;
  if (!hat_s) {
    KB_CUDA(cudaMalloc((void**)&hat_s, 128 * (size_t)n2));
    KB_LAUNCH(ctx, fk_build_s_kernel, cdiv(n2, 256), 256, 0, ctx->d_srs, (uint32_t)d, hat_s);
    g1_ntt_dev(ctx, hat_s, logd + 1, false);
    kb_ctx::FkCache c; c.d = d; c.d_hat_s = hat_s;
    if (ctx->fk_cache.size() >= 4) {   // bounded: the oldest transform goes (the stream is idle between calls: entries are never in use here)
      KB_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->fk_cache.front().d_hat_s);
      ctx->fk_cache.erase(ctx->fk_cache.begin());
    }
    ctx->fk_cache.push_back(c);
  }
  DevBuf<uint32_t> a(ctx, 8 * (size_t)n2), h(ctx, 32 * (size_t)n2);
  KB_LAUNCH(ctx, fk_build_a_kernel, cdiv(n2, 256), 256, 0, d_coeffs, (uint32_t)d, a.p);
  fr_ntt_dev(ctx, a.p, logd + 1, false, false);
  Twiddles tw(ctx, logd + 1, true, false);  // only for (2d)^-1 at tw[d]
  if (n2 <= (1u << 14)) KB_LAUNCH(ctx, fk_pointwise4_kernel, cdiv(4ull * n2, 128), 128, 0, hat_s, a.p, tw.tw.p + 8 * (size_t)d, n2, h.p);
  else KB_LAUNCH(ctx, fk_pointwise_kernel, cdiv(n2, 128), 128, 0, hat_s, a.p, tw.tw.p + 8 * (size_t)d, n2, h.p);
  g1_ntt_dev(ctx, h.p, logd + 1, true);
  g1_ntt_dev(ctx, h.p, logd, false);  // first d entries (src/kzg.rs:197-200)
  KB_LAUNCH(ctx, g1_xyzz_to_affine_kernel, cdiv(d, 128), 128, 0, h.p, (uint32_t)d, d_proofs, d_inf);
}

}  // namespace kb
