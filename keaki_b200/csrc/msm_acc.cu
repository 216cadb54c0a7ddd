// MSM bucket accumulation — the dominant kernel of `commit` (src/kzg.rs:98).  Hot translation unit:
// the field multiplier and the XYZZ mixed addition are fully inlined here (KB_INLINE_ALL).
#define KB_INLINE_ALL
#include "ctx.cuh"

namespace kb {

// ------------------------------------------------------------------------------------------
// bucket accumulation: one thread per bucket, XYZZ mixed additions, next base prefetched
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) msm_accumulate_kernel(const uint32_t* __restrict__ tab, uint64_t tab_n, uint64_t first,
                                                                const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ entries,
                                                                const uint32_t* __restrict__ perm, uint32_t nb, uint32_t* __restrict__ buckets) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nb) return;
  const uint32_t b = perm[tid];   // buckets by decreasing population: the 32 lanes of a warp run equally long
  uint32_t lo = offsets[b], hi = offsets[b + 1];
  G1 acc = G1::infinity();
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

void launch_msm_accumulate(kb_ctx* ctx, const uint32_t* tab, uint64_t tab_n, uint64_t first, const uint32_t* offsets,
                           const uint32_t* entries, const uint32_t* perm, uint32_t nb, uint32_t* buckets) {
  KB_LAUNCH(ctx, msm_accumulate_kernel, cdiv(nb, 256), 256, 0, tab, tab_n, first, offsets, entries, perm, nb, buckets);
}


// ------------------------------------------------------------------------------------------
// bucket reduction: sum_b (b + 1) * B_b.  The group operations are real calls whose bodies have the field
// multiplier inlined, so ptxas interleaves the independent products inside one addition (the reduction is a
// chain of dependent additions per thread: instruction-level parallelism inside each is what hides latency).
// ------------------------------------------------------------------------------------------
__device__ __noinline__ G1 g1_add(G1 a, G1 b) { return ec_add(a, b); }
__device__ __noinline__ G1 g1_dbl(G1 a) { return ec_dbl(a); }

static constexpr int MSM_SEG = 8;  // buckets per thread in the reduction

// sum_b (b + 1) B_b over segments of MSM_SEG buckets: running sums inside the segment, the segment's weight offset
// by a short double-and-add, then a tree sum of the segment results and one affine normalisation.
__global__ void __launch_bounds__(128) msm_reduce_seg_kernel(const uint32_t* __restrict__ buckets, uint32_t nb,
                                                             uint32_t* __restrict__ partial) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lo = t * MSM_SEG;
  if (lo >= nb) return;
  uint32_t hi = lo + MSM_SEG < nb ? lo + MSM_SEG : nb;
  G1 run = G1::infinity(), sum = G1::infinity();
  for (uint32_t j = hi; j-- > lo;) {
    run = g1_add(run, ld_g1x(buckets + 32 * (uint64_t)j));
    sum = g1_add(sum, run);
  }
  // sum = sum_j (j - lo + 1) B_j ; add lo * run
  if (lo != 0 && !run.is_inf()) {
    G1 m = G1::infinity();
    for (int bit = 31 - __clz(lo); bit >= 0; bit--) {
      m = g1_dbl(m);
      if ((lo >> bit) & 1u) m = g1_add(m, run);
    }
    sum = g1_add(sum, m);
  }
  st_g1x(partial + 32 * (uint64_t)t, sum);
}

// tree sum of XYZZ points: each block folds up to 256 * per inputs into one output
__global__ void __launch_bounds__(256) g1_tree_sum_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t per,
                                                          uint32_t* __restrict__ out) {
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
  if (threadIdx.x == 0) st_g1x(out + 32 * (uint64_t)blockIdx.x, acc);
}

__global__ void g1_finalize_kernel(const uint32_t* __restrict__ xyzz, uint32_t* __restrict__ out_xy, uint8_t* __restrict__ out_inf) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  G1 p = ld_g1x(xyzz);
  G1Affine a = to_affine(p);
  st_g1(out_xy, a);
  if (out_inf) *out_inf = p.is_inf() ? 1 : 0;
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
  while (cur_n > 1) {
    // keep blocks full when there is a lot to fold, but never fewer than needed
    uint32_t per = cur_n >= (1u << 16) ? 4 : 1;
    unsigned blocks = cdiv(cur_n, 256ull * per);
    KB_LAUNCH(ctx, g1_tree_sum_kernel, blocks, 256, 0, src, cur_n, per, dst);
    cur_n = blocks;
    uint32_t* t = src; src = dst; dst = t;
  }
  KB_LAUNCH(ctx, g1_finalize_kernel, 1, 32, 0, src, d_out_xy, d_out_inf);
}

void g1_sum(kb_ctx* ctx, const uint32_t* d_pts, const uint8_t* d_inf, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf) {
  DevBuf<uint32_t> x(ctx, 32 * (size_t)(n ? n : 1));
  if (n) KB_LAUNCH(ctx, g1_affine_to_xyzz_kernel, cdiv(n, 256), 256, 0, d_pts, d_inf, n, x);
  g1_xyzz_sum_to_affine(ctx, x, n, d_out_xy, d_out_inf);
}


void launch_msm_reduce(kb_ctx* ctx, const uint32_t* buckets, uint32_t nb, uint32_t* d_out_xy, uint8_t* d_out_inf) {
  const uint32_t nseg = cdiv(nb, MSM_SEG);
  DevBuf<uint32_t> partial(ctx, 32 * (size_t)nseg);
  KB_LAUNCH(ctx, msm_reduce_seg_kernel, cdiv(nseg, 128), 128, 0, buckets, nb, partial);
  g1_xyzz_sum_to_affine(ctx, partial, nseg, d_out_xy, d_out_inf);
}

}  // namespace kb
