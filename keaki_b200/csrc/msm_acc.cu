// MSM bucket accumulation — the dominant kernel of `commit` (src/kzg.rs:98).  Hot translation unit:
// the field multiplier and the XYZZ mixed addition are fully inlined here (KB_INLINE_ALL).
#define KB_INLINE_ALL
#include "ctx.cuh"
#include "quad.cuh"

#include <algorithm>

namespace kb {

// entries one thread sums at most (see "Over-full buckets" below): above the largest bucket uniform scalars produce
// (26 +- 5 at 2^20 - 109 in the 12,388 buckets the 15-bit top window of a scalar below r reaches - 32 +- 6 at 2^16) and short enough that a lone thread's chain (5 us per dependent addition) stays
// well below the kernel's own time
static constexpr uint32_t MSM_SEG = 256;   // above the fullest bucket of uniform scalars: the top window of c = 20 has 14 bits below r, 6,194 buckets of 169 entries on average at 2^20 points

// ------------------------------------------------------------------------------------------
// bucket accumulation: one thread per bucket, XYZZ mixed additions, next base prefetched
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) msm_accumulate_kernel(const uint32_t* __restrict__ tab, uint64_t tab_n, uint64_t first,
                                                                const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ entries,
                                                                const uint32_t* __restrict__ perm, uint32_t nb, uint32_t* __restrict__ buckets,
                                                                int into) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nb) return;
  const uint32_t b = perm[tid];   // buckets by decreasing population: the 32 lanes of a warp run equally long
  uint32_t lo = offsets[b], hi = offsets[b + 1];
  if (hi - lo > MSM_SEG) hi = lo + MSM_SEG;   // the rest of an over-full bucket is summed in segments by other threads (below)
  G1 acc = into ? ld_g1x(buckets + 32 * (uint64_t)b) : G1::infinity();   // a second pass continues the first one's sums
  if (lo < hi) {
    uint32_t e = entries[lo];
    G1Affine nxt = ld_g1(tab + 16 * ((uint64_t)((e >> 26) & 31u) * tab_n + first + (e & 0x3ffffffu)));
    for (uint32_t k = lo; k < hi; k++) {
      G1Affine cur = nxt;
      bool negate = (e >> 31) != 0;
      if (k + 1 < hi) {
        e = entries[k + 1];
        nxt = ld_g1(tab + 16 * ((uint64_t)((e >> 26) & 31u) * tab_n + first + (e & 0x3ffffffu)));
      }
      if (negate) cur.y = -cur.y;
      acc = ec_add_mixed(acc, cur);
    }
  }
  st_g1x(buckets + 32 * (uint64_t)b, acc);
}

// ------------------------------------------------------------------------------------------
// Over-full buckets.  All windows share one bucket set, so structured scalars (all equal, 0/1 vectors, small
// coefficients) put O(n) entries into a handful of buckets, and one thread per bucket would run O(n) dependent
// additions (5 us each for a lone thread).  The owner of a bucket therefore sums only its first MSM_SEG entries; the remainder is cut into segments
// of MSM_SEG entries, each summed by a thread of its own into a partial, and the partials of a bucket are folded onto
// it (a thread for a few segments, a block for many).  Uniform scalars stay below MSM_SEG: the kernels below then find
// empty lists and return (about 20 us per call).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msm_overflow_list_kernel(const uint32_t* __restrict__ offsets, uint32_t nb, uint32_t* __restrict__ counters /* nseg, nsplit */,
                                                                uint2* __restrict__ seg_list, uint4* __restrict__ split_list) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint32_t size = offsets[b + 1] - offsets[b];
  if (size <= MSM_SEG) return;
  const uint32_t k = (size - 1) / MSM_SEG;   // segments after the first
  const uint32_t base = atomicAdd(&counters[0], k);
  split_list[atomicAdd(&counters[1], 1u)] = make_uint4(b, base, k, 0u);
  for (uint32_t j = 0; j < k; j++) seg_list[base + j] = make_uint2(b, j + 1);
}
__global__ void __launch_bounds__(256, 2) msm_overflow_acc_kernel(const uint32_t* __restrict__ tab, uint64_t tab_n, uint64_t first,
                                                                  const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ entries,
                                                                  const uint32_t* __restrict__ counters, const uint2* __restrict__ seg_list,
                                                                  uint32_t* __restrict__ partial) {
  const uint32_t nseg = counters[0];
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nseg; t += gridDim.x * blockDim.x) {
    const uint2 sg = seg_list[t];
    const uint32_t lo = offsets[sg.x] + sg.y * MSM_SEG, end = offsets[sg.x + 1];
    const uint32_t hi = end - lo > MSM_SEG ? lo + MSM_SEG : end;
    G1 acc = G1::infinity();
    for (uint32_t k = lo; k < hi; k++) {
      const uint32_t e = entries[k];
      G1Affine cur = ld_g1(tab + 16 * ((uint64_t)((e >> 26) & 31u) * tab_n + first + (e & 0x3ffffffu)));
      if (e >> 31) cur.y = -cur.y;
      acc = ec_add_mixed(acc, cur);
    }
    st_g1x(partial + 32 * (uint64_t)t, acc);
  }
}
__device__ __noinline__ G1 g1_add_cold(G1 a, G1 b) { return ec_add(a, b); }
// buckets with few segments (the common case: a bucket slightly over the threshold): one thread adds them in sequence
static constexpr uint32_t MSM_FOLD_SERIAL = 8;
__global__ void __launch_bounds__(128) msm_overflow_fold_small_kernel(const uint32_t* __restrict__ counters, const uint4* __restrict__ split_list,
                                                                      const uint32_t* __restrict__ partial, uint32_t* __restrict__ buckets) {
  const uint32_t nsplit = counters[1];
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nsplit; s += gridDim.x * blockDim.x) {
    const uint4 sp = split_list[s];
    if (sp.z > MSM_FOLD_SERIAL) continue;
    G1 acc = ld_g1x(buckets + 32 * (uint64_t)sp.x);
    for (uint32_t i = 0; i < sp.z; i++) acc = g1_add_cold(acc, ld_g1x(partial + 32 * (uint64_t)(sp.y + i)));
    st_g1x(buckets + 32 * (uint64_t)sp.x, acc);
  }
}
// buckets with many segments: one block per bucket, strided partial sums + a shared-memory tree
__global__ void __launch_bounds__(256) msm_overflow_fold_kernel(const uint32_t* __restrict__ counters, const uint4* __restrict__ split_list,
                                                                const uint32_t* __restrict__ partial, uint32_t* __restrict__ buckets) {
  __shared__ uint32_t sm[128 * 32];
  const uint32_t nsplit = counters[1];
  for (uint32_t s = blockIdx.x; s < nsplit; s += gridDim.x) {
    const uint4 sp = split_list[s];
    if (sp.z <= MSM_FOLD_SERIAL) continue;   // block-uniform
    G1 acc = G1::infinity();
    for (uint32_t i = threadIdx.x; i < sp.z; i += blockDim.x) acc = g1_add_cold(acc, ld_g1x(partial + 32 * (uint64_t)(sp.y + i)));
    for (int half = 128; half >= 1; half >>= 1) {
      if (threadIdx.x >= half && threadIdx.x < 2 * half) {
        uint32_t* d = sm + 32 * (threadIdx.x - half);
#pragma unroll
        for (int q = 0; q < 8; q++) { d[q] = acc.x.v[q]; d[8 + q] = acc.y.v[q]; d[16 + q] = acc.zz.v[q]; d[24 + q] = acc.zzz.v[q]; }
      }
      __syncthreads();
      if (threadIdx.x < half) {
        const uint32_t* d = sm + 32 * threadIdx.x;
        G1 o;
#pragma unroll
        for (int q = 0; q < 8; q++) { o.x.v[q] = d[q]; o.y.v[q] = d[8 + q]; o.zz.v[q] = d[16 + q]; o.zzz.v[q] = d[24 + q]; }
        acc = g1_add_cold(acc, o);
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) st_g1x(buckets + 32 * (uint64_t)sp.x, g1_add_cold(ld_g1x(buckets + 32 * (uint64_t)sp.x), acc));
    __syncthreads();
  }
}

void launch_msm_accumulate(kb_ctx* ctx, const uint32_t* tab, uint64_t tab_n, uint64_t first, const uint32_t* offsets,
                           const uint32_t* entries, const uint32_t* perm, uint32_t nb, uint32_t* buckets, bool into, uint64_t max_entries) {
  KB_LAUNCH(ctx, msm_accumulate_kernel, cdiv(nb, 256), 256, 0, tab, tab_n, first, offsets, entries, perm, nb, buckets, into ? 1 : 0);
  const uint64_t max_seg = max_entries / MSM_SEG + 1;   // segments (and split buckets) there can be at most
  DevBuf<uint32_t> counters(ctx, 2), partial(ctx, 32 * max_seg);
  DevBuf<uint2> seg_list(ctx, max_seg);
  DevBuf<uint4> split_list(ctx, max_seg);
  KB_CUDA(cudaMemsetAsync(counters.p, 0, 8, ctx->stream));
  KB_LAUNCH(ctx, msm_overflow_list_kernel, cdiv(nb, 256), 256, 0, offsets, nb, counters.p, seg_list.p, split_list.p);
  const unsigned acc_blocks = (unsigned)std::min<uint64_t>(cdiv(max_seg, 256), 2ull * ctx->sm_count);
  KB_LAUNCH(ctx, msm_overflow_acc_kernel, acc_blocks, 256, 0, tab, tab_n, first, offsets, entries, counters.p, seg_list.p, partial.p);
  const unsigned small_blocks = (unsigned)std::min<uint64_t>(cdiv(max_seg, 128), 4ull * ctx->sm_count);
  KB_LAUNCH(ctx, msm_overflow_fold_small_kernel, small_blocks, 128, 0, counters.p, split_list.p, partial.p, buckets);
  const unsigned fold_blocks = (unsigned)std::min<uint64_t>(max_seg, 2ull * ctx->sm_count);
  KB_LAUNCH(ctx, msm_overflow_fold_kernel, fold_blocks, 256, 0, counters.p, split_list.p, partial.p, buckets);
}


// ------------------------------------------------------------------------------------------
// bucket reduction: sum_b (b + 1) * B_b.  The group operations are real calls whose bodies have the field
// multiplier inlined, so ptxas interleaves the independent products inside one addition (the reduction is a
// chain of dependent additions per thread: instruction-level parallelism inside each is what hides latency).
// ------------------------------------------------------------------------------------------
__device__ __noinline__ G1 g1_add(G1 a, G1 b) { return ec_add(a, b); }
__device__ __noinline__ G1 g1_dbl(G1 a) { return ec_dbl(a); }

// Bucket reduction, two-digit form.  Bucket b (weight b + 1) is written b = hi * C + lo with C = 2^k columns and
// R = nb / C rows, so that
//     sum_b (b + 1) B_b = sum_lo (lo + 1) L_lo + sum_hi (hi * C) H_hi,   L_lo = column sums, H_hi = row sums
// (the running-sum form needs the same two additions per bucket but as 2^(c-1)-long dependent chains; here every
// addition of the bulk is independent work for the multiplier pipe and only C + R points are left for the weights).
// Both families of sums are folds of the flat bucket array: rows by adding G consecutive elements, columns by adding
// G elements n_out apart; each launch runs one fold step of each family.
struct FoldJob { const uint32_t* in; uint32_t* out; uint32_t n_out, G, sj, si; };   // out[j] = sum_{i<G} in[j * sj + i * si]

__global__ void __launch_bounds__(128) g1x_fold2_kernel(FoldJob ja, FoldJob jb, uint32_t blocks_a) {
  const bool second = blockIdx.x >= blocks_a;
  const FoldJob J = second ? jb : ja;
  const uint32_t j = (blockIdx.x - (second ? blocks_a : 0u)) * blockDim.x + threadIdx.x;
  if (j >= J.n_out) return;
  const uint32_t* p = J.in + 32 * ((uint64_t)j * J.sj);
  G1 acc = ld_g1x(p);
  if (J.G > 1) {
    G1 nxt = ld_g1x(p + 32 * (uint64_t)J.si);
    for (uint32_t i = 1; i < J.G; i++) {
      G1 cur = nxt;
      if (i + 1 < J.G) nxt = ld_g1x(p + 32 * (uint64_t)(i + 1) * J.si);
      acc = g1_add(acc, cur);
    }
  }
  st_g1x(J.out + 32 * (uint64_t)j, acc);
}

// The same two folds with G adjacent LANES per output (G a power of two <= 32, 32 / G outputs per warp): lane i of a
// group loads element i and a log2(G)-level shuffle tree adds them.  Used once the arrays are small: the fold is then
// bound by the length of its dependent chain (log2 G additions against G - 1 in a thread of its own), not by throughput.
__device__ __forceinline__ G1 g1x_shfl_down(const G1& a, int d, int width) {
  G1 r;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    r.x.v[q] = __shfl_down_sync(0xffffffffu, a.x.v[q], d, width);
    r.y.v[q] = __shfl_down_sync(0xffffffffu, a.y.v[q], d, width);
    r.zz.v[q] = __shfl_down_sync(0xffffffffu, a.zz.v[q], d, width);
    r.zzz.v[q] = __shfl_down_sync(0xffffffffu, a.zzz.v[q], d, width);
  }
  return r;
}
__global__ void __launch_bounds__(128) g1x_fold2_lanes_kernel(FoldJob ja, FoldJob jb, uint32_t blocks_a) {
  const bool second = blockIdx.x >= blocks_a;
  const FoldJob J = second ? jb : ja;
  const uint32_t t = (blockIdx.x - (second ? blocks_a : 0u)) * blockDim.x + threadIdx.x;
  const uint32_t j = t / J.G, li = t % J.G;
  const bool live = j < J.n_out;                 // whole groups are live or not; shuffles need every lane
  G1 acc = live ? ld_g1x(J.in + 32 * ((uint64_t)j * J.sj + (uint64_t)li * J.si)) : G1::infinity();
  for (int d = (int)J.G >> 1; d >= 1; d >>= 1) {
    G1 o = g1x_shfl_down(acc, d, (int)J.G);
    if (li < (uint32_t)d) acc = g1_add(acc, o);
  }
  if (live && li == 0) st_g1x(J.out + 32 * (uint64_t)j, acc);
}

// weight * P by double-and-add, MSB first
__device__ __forceinline__ G1 g1_mul_small(const G1& p, uint32_t w) {
  G1 m = G1::infinity();
  if (w == 0 || p.is_inf()) return m;
  for (int bit = 31 - __clz(w); bit >= 0; bit--) {
    m = g1_dbl(m);
    if ((w >> bit) & 1u) m = g1_add(m, p);
  }
  return m;
}

// thread t < C: (t + 1) L_t;  C <= t < C + R: ((t - C) * C) H_(t-C);  each block leaves the sum of its 128 terms
__global__ void __launch_bounds__(128) msm_weigh_kernel(const uint32_t* __restrict__ L, uint32_t C, const uint32_t* __restrict__ H, uint32_t R,
                                                        uint32_t* __restrict__ out) {
  __shared__ uint32_t sm[64 * 32];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  G1 acc = G1::infinity();
  if (t < C) acc = g1_mul_small(ld_g1x(L + 32 * (uint64_t)t), t + 1u);
  else if (t < C + R) acc = g1_mul_small(ld_g1x(H + 32 * (uint64_t)(t - C)), (t - C) * C);
  for (int half = 64; half >= 1; half >>= 1) {
    if (threadIdx.x >= half && threadIdx.x < 2 * half) {
      uint32_t* s = sm + 32 * (threadIdx.x - half);
#pragma unroll
      for (int q = 0; q < 8; q++) { s[q] = acc.x.v[q]; s[8 + q] = acc.y.v[q]; s[16 + q] = acc.zz.v[q]; s[24 + q] = acc.zzz.v[q]; }
    }
    __syncthreads();
    if (threadIdx.x < half) {
      const uint32_t* s = sm + 32 * threadIdx.x;
      G1 o;
#pragma unroll
      for (int q = 0; q < 8; q++) { o.x.v[q] = s[q]; o.y.v[q] = s[8 + q]; o.zz.v[q] = s[16 + q]; o.zzz.v[q] = s[24 + q]; }
      acc = g1_add(acc, o);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_g1x(out + 32 * (uint64_t)blockIdx.x, acc);
}

__device__ __forceinline__ void g1x_to_smem(uint32_t* s, const G1& a) {
#pragma unroll
  for (int q = 0; q < 8; q++) { s[q] = a.x.v[q]; s[8 + q] = a.y.v[q]; s[16 + q] = a.zz.v[q]; s[24 + q] = a.zzz.v[q]; }
}
__device__ __forceinline__ G1 g1x_from_smem(const uint32_t* s) {
  G1 o;
#pragma unroll
  for (int q = 0; q < 8; q++) { o.x.v[q] = s[q]; o.y.v[q] = s[8 + q]; o.zz.v[q] = s[16 + q]; o.zzz.v[q] = s[24 + q]; }
  return o;
}
// sum over the NQ quads of a block (acc replicated per quad); the result is valid in quad 0
template <int NQ>
__device__ __forceinline__ G1 quad_block_sum(G1 acc, const QuadLane& ql, uint32_t* sm /* NQ/2 * 32 words */) {
  const int qi = threadIdx.x >> 2;
  for (int half = NQ / 2; half >= 1; half >>= 1) {
    if (qi >= half && qi < 2 * half && ql.q == 0) g1x_to_smem(sm + 32 * (qi - half), acc);
    __syncthreads();
    if (qi < half) acc = g1_add4(acc, g1x_from_smem(sm + 32 * qi), ql);
    __syncthreads();
  }
  return acc;
}

// quad j < C: (j + 1) L_j;  C <= j < C + R: ((j - C) * C) H_(j-C), double-and-add over the quad; each block of 32 quads
// leaves the sum of its terms
__global__ void __launch_bounds__(128) msm_weigh4_kernel(const uint32_t* __restrict__ L, uint32_t C, const uint32_t* __restrict__ H, uint32_t R,
                                                         uint32_t* __restrict__ out) {
  __shared__ uint32_t sm[16 * 32];
  const QuadLane ql = quad_lane();
  const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  G1 p = G1::infinity();
  uint32_t w = 0;
  if (j < C) { p = ld_g1x(L + 32 * (uint64_t)j); w = j + 1u; }
  else if (j < C + R) { p = ld_g1x(H + 32 * (uint64_t)(j - C)); w = (j - C) * C; }
  G1 acc = G1::infinity();
  if (w != 0 && !p.is_inf()) {
    acc = p;
    for (int bit = 30 - __clz(w); bit >= 0; bit--) {
      acc = g1_dbl4(acc, ql);
      if ((w >> bit) & 1u) acc = g1_add4(acc, p, ql);
    }
  }
  acc = quad_block_sum<32>(acc, ql, sm);
  if (threadIdx.x == 0) st_g1x(out + 32 * (uint64_t)blockIdx.x, acc);
}

// sum of n <= 64 points by one block of 64 quads, affine result
__global__ void __launch_bounds__(256) g1_sum4_final_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out_xy,
                                                            uint8_t* __restrict__ out_inf) {
  __shared__ uint32_t sm[32 * 32];
  const QuadLane ql = quad_lane();
  const uint32_t qi = threadIdx.x >> 2;
  G1 acc = qi < n ? ld_g1x(in + 32 * (uint64_t)qi) : G1::infinity();
  acc = quad_block_sum<64>(acc, ql, sm);
  if (threadIdx.x == 0) {
    st_g1(out_xy, to_affine(acc));
    if (out_inf) *out_inf = acc.is_inf() ? 1 : 0;
  }
}

// tree sum of XYZZ points: each block folds up to 256 * per inputs into one output
// (with out_xy set, the single block of the last step also writes the affine result)
__global__ void __launch_bounds__(256) g1_tree_sum_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t per,
                                                          uint32_t* __restrict__ out, uint32_t* __restrict__ out_xy, uint8_t* __restrict__ out_inf) {
  __shared__ uint32_t sm[128 * 32];
  uint64_t base = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * per;
  G1 acc = G1::infinity();
  for (uint32_t k = 0; k < per; k++) if (base + k < n) acc = g1_add(acc, ld_g1x(in + 32 * (base + k)));
  for (int half = 128; half >= 1; half >>= 1) {
    if (threadIdx.x >= half && threadIdx.x < 2 * half) {
      uint32_t* s = sm + 32 * (threadIdx.x - half);
#pragma unroll
      for (int q = 0; q < 8; q++) { s[q] = acc.x.v[q]; s[8 + q] = acc.y.v[q]; s[16 + q] = acc.zz.v[q]; s[24 + q] = acc.zzz.v[q]; }
    }
    __syncthreads();
    if (threadIdx.x < half) {
      const uint32_t* s = sm + 32 * threadIdx.x;
      G1 o;
#pragma unroll
      for (int q = 0; q < 8; q++) { o.x.v[q] = s[q]; o.y.v[q] = s[8 + q]; o.zz.v[q] = s[16 + q]; o.zzz.v[q] = s[24 + q]; }
      acc = g1_add(acc, o);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (out_xy) {
      st_g1(out_xy, to_affine(acc));
      if (out_inf) *out_inf = acc.is_inf() ? 1 : 0;
    } else st_g1x(out + 32 * (uint64_t)blockIdx.x, acc);
  }
}

__global__ void __launch_bounds__(256) g1_affine_to_xyzz_kernel(const uint32_t* __restrict__ pts, const uint8_t* __restrict__ inf,
                                                                uint64_t n, uint32_t* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine a = ld_g1(pts + 16 * i);
  if (inf && inf[i]) a = G1Affine::infinity();
  st_g1x(out + 32 * i, to_xyzz(a));
}

// Sums n XYZZ points (array is consumed) and writes the affine result.
void g1_xyzz_sum_to_affine(kb_ctx* ctx, uint32_t* d_xyzz, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf) {
  if (n == 0) {
    KB_CUDA(cudaMemsetAsync(d_out_xy, 0, 64, ctx->stream));
    if (d_out_inf) KB_CUDA(cudaMemsetAsync(d_out_inf, 1, 1, ctx->stream));
    return;
  }
  uint64_t cur_n = n;
  DevBuf<uint32_t> tmp(ctx, 32 * (size_t)cdiv(n, 256));
  uint32_t* src = d_xyzz;
  uint32_t* dst = tmp;
  for (;;) {
    // keep blocks full when there is a lot to fold, but never fewer than needed
    uint32_t per = cur_n >= (1u << 16) ? 4 : 1;
    unsigned blocks = cdiv(cur_n, 256ull * per);
    const bool last = blocks == 1;
    KB_LAUNCH(ctx, g1_tree_sum_kernel, blocks, 256, 0, src, cur_n, per, dst, last ? d_out_xy : nullptr, d_out_inf);
    if (last) break;
    cur_n = blocks;
    uint32_t* t = src; src = dst; dst = t;
  }
}

void g1_sum(kb_ctx* ctx, const uint32_t* d_pts, const uint8_t* d_inf, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf) {
  DevBuf<uint32_t> x(ctx, 32 * (size_t)(n ? n : 1));
  if (n) KB_LAUNCH(ctx, g1_affine_to_xyzz_kernel, cdiv(n, 256), 256, 0, d_pts, d_inf, n, x);
  g1_xyzz_sum_to_affine(ctx, x, n, d_out_xy, d_out_inf);
}


void launch_msm_reduce(kb_ctx* ctx, const uint32_t* buckets, uint32_t nb, uint32_t* d_out_xy, uint8_t* d_out_inf) {
  // nb = 2^(c-1) buckets as R rows of C = 2^k columns, k = ceil((c-1)/2)
  int lg = 0;
  while ((1u << lg) < nb) lg++;
  const uint32_t C = 1u << ((lg + 1) / 2), R = nb / C;
  DevBuf<uint32_t> bufL(ctx, 32 * (size_t)(nb / 4 + C)), bufH(ctx, 32 * (size_t)(nb / 4 + R));
  const uint32_t *inL = buckets, *inH = buckets;
  uint32_t nL = nb, nH = nb;
  uint32_t *outL = bufL, *outH = bufH;
  // large arrays: one thread per output, G = 8 (throughput-bound, keeps the multiplier pipe full);
  // below 2^15 elements: G lanes per output
  while (nL > C || nH > R) {
    const bool lanes = (nL > C ? nL : nH) < (1u << 15);
    const uint32_t gmax = lanes ? 32u : 8u;
    FoldJob ja = {nullptr, nullptr, 0, 1, 0, 0}, jb = ja;
    if (nL > C) {   // columns: fold the array onto its first n_out elements (n_out stays a multiple of C)
      uint32_t G = nL / C < gmax ? nL / C : gmax, n_out = nL / G;
      ja = FoldJob{inL, outL, n_out, G, 1u, n_out};
      inL = outL; outL += 32 * (size_t)n_out; nL = n_out;
    }
    if (nH > R) {   // rows: G consecutive elements (the current row length nH / R is a multiple of G)
      uint32_t G = nH / R < gmax ? nH / R : gmax, n_out = nH / G;
      jb = FoldJob{inH, outH, n_out, G, G, 1u};
      inH = outH; outH += 32 * (size_t)n_out; nH = n_out;
    }
    if (!lanes) {
      const unsigned ba = cdiv(ja.n_out, 128), bb = cdiv(jb.n_out, 128);
      KB_LAUNCH(ctx, g1x_fold2_kernel, ba + bb, 128, 0, ja, jb, ba);
    } else {
      const unsigned ba = cdiv((uint64_t)ja.n_out * ja.G, 128), bb = cdiv((uint64_t)jb.n_out * jb.G, 128);
      KB_LAUNCH(ctx, g1x_fold2_lanes_kernel, ba + bb, 128, 0, ja, jb, ba);
    }
  }
  const unsigned wb = cdiv(C + R, 128);
  DevBuf<uint32_t> part(ctx, 32 * (size_t)(wb > 64 ? wb : 64));
  if (const unsigned wb4 = cdiv(4ull * (C + R), 128); wb4 <= 64) {   // quad-cooperative tail
    KB_LAUNCH(ctx, msm_weigh4_kernel, wb4, 128, 0, inL, C, inH, R, part);
    KB_LAUNCH(ctx, g1_sum4_final_kernel, 1, 256, 0, part, wb4, d_out_xy, d_out_inf);
    return;
  }
  KB_LAUNCH(ctx, msm_weigh_kernel, wb, 128, 0, inL, C, inH, R, part);
  g1_xyzz_sum_to_affine(ctx, part, wb, d_out_xy, d_out_inf);
}

}  // namespace kb
