// G1 multi-scalar multiplication over the resident SRS — the kernel behind `commit`
// (src/kzg.rs:89-101: VariableBaseMSM::msm_unchecked(&setup.g1_aff, p)) and `open` (src/kzg.rs:123).
//
// B200-first design:
//  * The bases are the fixed SRS, so HBM capacity is traded for work: for a window width c the
//    table tab[w][i] = 2^(c w) * g1[i] is built once per SRS (affine, 64 B per entry).  All
//    ceil(255/c) windows of all points then fall into ONE set of 2^(c-1) signed-digit buckets:
//    one bucket reduction instead of one per window and no final Horner doubling chain.
//  * Scalars are taken out of Montgomery form and recoded into signed c-bit digits on the fly
//    (never stored); entries (point, window, sign) are counting-sorted by bucket with global
//    atomics; one thread owns one bucket and accumulates its entries with XYZZ mixed additions,
//    prefetching the next base while the current addition runs.
//  * Bucket reduction, two-digit form (msm_acc.cu launch_msm_reduce): bucket b = hi * C + lo has weight
//    (lo + 1) + hi * C, so the weighted sum splits into column sums and row sums of the flat bucket
//    array (independent additions), a short double-and-add over the C + R folded points and a tree sum.
// The result leaves as canonical affine coordinates, so it is bit-identical to arkworks' for any
// summation order.
#include "ctx.cuh"
#include "msm_digits.cuh"
#include "consts_gen.cuh"
#include <cstdlib>

namespace kb {


// ------------------------------------------------------------------------------------------
// window choice and table build
// ------------------------------------------------------------------------------------------
static int msm_choose_c(uint64_t end) {
  // tuning override; c >= 8 keeps ceil(255 / c) <= 32 windows (the digit array and the 5-bit window field of an entry)
  if (const char* e = getenv("KB_MSM_C")) { int c = atoi(e); if (c >= 8 && c <= 24) return c; }
  // measured on B200 (tools/exp/msm_c_sweep.py): c = 17 wins from 2^12 to 2^16 points, c = 20 from 2^17 on;
  // both divide 255 into windows whose TOP window is full (255 = 15 x 17 = 12 x 20 + 15) - a short top window (c = 13: 7 bits,
  // c = 14 or 18: 3 bits) funnels every top digit into a few dozen buckets, whose owners then run n / 48 dependent additions
  if (end <= (1ull << 16)) return 17;
  return 20;
}
static uint64_t msm_table_cap(int c, uint64_t srs_n) {
  uint64_t cap = c == 17 ? (1ull << 16) : srs_n;
  if (getenv("KB_MSM_C")) cap = srs_n;
  return cap < srs_n ? cap : srs_n;
}

__global__ void __launch_bounds__(128) msm_build_table_kernel(const uint32_t* __restrict__ srs, uint64_t n, int c, int nwin,
                                                              uint32_t* __restrict__ tab) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = ld_g1(srs + 16 * i);
  st_g1(tab + 16 * i, p);
  G1 acc = to_xyzz(p);
  for (int w = 1; w < nwin; w++) {
    for (int k = 0; k < c; k++) acc = ec_dbl(acc);
    st_g1(tab + 16 * ((uint64_t)w * n + i), to_affine(acc));
  }
}

static const MsmTable& msm_get_table(kb_ctx* ctx, int c) {
  for (auto& t : ctx->msm_tabs) if (t.c == c) return t;
  MsmTable t;
  t.c = c; t.nwin = msm_num_windows(c); t.n = msm_table_cap(c, ctx->srs_n);
  KB_CUDA(cudaMalloc((void**)&t.d, (size_t)t.nwin * t.n * 64));
  KB_LAUNCH(ctx, msm_build_table_kernel, cdiv(t.n, 128), 128, 0, ctx->d_srs, t.n, c, t.nwin, t.d);
  ctx->msm_tabs.push_back(t);
  return ctx->msm_tabs.back();
}

void msm_free_tables(kb_ctx* ctx) {
  for (auto& t : ctx->msm_tabs) if (t.d) cudaFree(t.d);
  ctx->msm_tabs.clear();
}

// ------------------------------------------------------------------------------------------
// recode + count, scan, scatter
// ------------------------------------------------------------------------------------------
// entry packing: bit 31 = negate, bits 26..30 = window, bits 0..25 = point index within the call
__device__ __forceinline__ void msm_load_digits(const uint32_t* scalars, uint64_t i, int c, int nwin, int32_t* dg) {
  Fr s = fp_load<FrParams>(scalars + 8 * i);
  Fr k = fp_from_mont<FrParams>(s);
  msm_signed_digits(k.v, c, nwin, dg);
}

__global__ void __launch_bounds__(256) msm_count_kernel(const uint32_t* __restrict__ scalars, uint64_t n, int c, int nwin,
                                                        uint32_t* __restrict__ counts) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t dg[32];
  msm_load_digits(scalars, i, c, nwin, dg);
  for (int w = 0; w < nwin; w++) {
    int32_t d = dg[w];
    if (d != 0) atomicAdd(&counts[(d < 0 ? -d : d) - 1], 1u);
  }
}

__global__ void __launch_bounds__(256) msm_scatter_kernel(const uint32_t* __restrict__ scalars, uint64_t n, int c, int nwin,
                                                          uint32_t* __restrict__ cursor, uint32_t* __restrict__ entries) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t dg[32];
  msm_load_digits(scalars, i, c, nwin, dg);
  // eight windows at a time: all the cursor atomics of a group are in flight together before their results are used
  for (int w0 = 0; w0 < nwin; w0 += 8) {
    uint32_t pos[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      int32_t d = w0 + k < nwin ? dg[w0 + k] : 0;
      pos[k] = d != 0 ? atomicAdd(&cursor[(uint32_t)(d < 0 ? -d : d) - 1u], 1u) : 0xffffffffu;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      int32_t d = w0 + k < nwin ? dg[w0 + k] : 0;
      if (d != 0) entries[pos[k]] = (uint32_t)i | ((uint32_t)(w0 + k) << 26) | (d < 0 ? 0x80000000u : 0u);
    }
  }
}

// Exclusive scan of `n` u32 counts: block-local scan (1024 per block) + serial scan of block sums.
__global__ void __launch_bounds__(256) scan_local_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                         uint32_t* __restrict__ block_sums, uint32_t n) {
  __shared__ uint32_t warp_tot[8];
  uint32_t base = blockIdx.x * 1024u + threadIdx.x * 4u;
  uint32_t v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] = base + k < n ? in[base + k] : 0u;
  uint32_t tsum = v[0] + v[1] + v[2] + v[3];
  uint32_t incl = tsum;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int k = 0; k < wid; k++) woff += warp_tot[k];
  uint32_t excl = woff + incl - tsum;
#pragma unroll
  for (int k = 0; k < 4; k++) { if (base + k < n) out[base + k] = excl; excl += v[k]; }
  if (threadIdx.x == 255) block_sums[blockIdx.x] = woff + incl;
}
__global__ void scan_sums_kernel(uint32_t* block_sums, uint32_t nblocks) {
  // single thread: nblocks <= 2^(c-1)/1024 <= 512
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint32_t acc = 0;
    for (uint32_t i = 0; i < nblocks; i++) { uint32_t t = block_sums[i]; block_sums[i] = acc; acc += t; }
    block_sums[nblocks] = acc;
  }
}
__global__ void __launch_bounds__(256) scan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ block_sums,
                                                       uint32_t* __restrict__ cursor, uint32_t n, uint32_t nblocks) {
  uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i < n) { uint32_t v = out[i] + block_sums[i >> 10]; out[i] = v; cursor[i] = v; }
  if (i == n) out[n] = block_sums[nblocks];  // total
}

// ------------------------------------------------------------------------------------------
// Load balancing: buckets sorted by population, largest first.  One thread owns one bucket, so a warp runs as
// long as its fullest bucket; with ~26 +- 5 entries per bucket (and twice that in the buckets the short top
// window feeds) a warp of arbitrary buckets idles ~30 % of its lanes.  A counting sort of the bucket indices by
// size gives every warp 32 buckets of (almost) equal length and schedules the long ones first.
// ------------------------------------------------------------------------------------------
static constexpr uint32_t MSM_SIZE_BINS = 1024;   // sizes >= 1023 share the first bin (they are scheduled first anyway)

__global__ void __launch_bounds__(256) msm_size_hist_kernel(const uint32_t* __restrict__ offsets, uint32_t nb, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[MSM_SIZE_BINS];
  for (uint32_t i = threadIdx.x; i < MSM_SIZE_BINS; i += 256) sh[i] = 0;
  __syncthreads();
  uint32_t b = blockIdx.x * 256u + threadIdx.x;
  if (b < nb) {
    uint32_t sz = offsets[b + 1] - offsets[b];
    atomicAdd(&sh[MSM_SIZE_BINS - 1 - (sz < MSM_SIZE_BINS - 1 ? sz : MSM_SIZE_BINS - 1)], 1u);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < MSM_SIZE_BINS; i += 256) if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// exclusive scan of the 1024 bins in place (one block)
__global__ void __launch_bounds__(1024) msm_size_scan_kernel(uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[MSM_SIZE_BINS];
  uint32_t v = hist[threadIdx.x];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (uint32_t o = 1; o < MSM_SIZE_BINS; o <<= 1) {
    uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  hist[threadIdx.x] = sh[threadIdx.x] - v;
}
// block-local ranks in shared memory, one global atomic per (block, occupied bin)
__global__ void __launch_bounds__(256) msm_size_scatter_kernel(const uint32_t* __restrict__ offsets, uint32_t nb, uint32_t* __restrict__ cursor,
                                                               uint32_t* __restrict__ perm) {
  __shared__ uint32_t sh[MSM_SIZE_BINS];
  for (uint32_t i = threadIdx.x; i < MSM_SIZE_BINS; i += 256) sh[i] = 0;
  __syncthreads();
  uint32_t b = blockIdx.x * 256u + threadIdx.x;
  uint32_t bin = 0, rank = 0;
  if (b < nb) {
    uint32_t sz = offsets[b + 1] - offsets[b];
    bin = MSM_SIZE_BINS - 1 - (sz < MSM_SIZE_BINS - 1 ? sz : MSM_SIZE_BINS - 1);
    rank = atomicAdd(&sh[bin], 1u);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < MSM_SIZE_BINS; i += 256) if (sh[i]) sh[i] = atomicAdd(&cursor[i], sh[i]);
  __syncthreads();
  if (b < nb) perm[sh[bin] + rank] = b;
}

// ------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------
// One pass over a range of points = sort (recode + counting sort of its (point, window) entries by bucket, buckets
// ordered by population) + accumulation.  The sort is bound by L2 atomics and scattered stores, the accumulation by the
// multiplier pipe, so the sort of the NEXT pass runs on a side stream underneath the accumulation of the current one.
struct MsmSort {
  DevBuf<uint32_t> counts, offsets, cursor, bsums, entries, size_bins, perm;
  uint64_t n;
  MsmSort(kb_ctx* ctx, uint32_t nb, uint64_t n_, int nwin)
      : counts(ctx, nb), offsets(ctx, nb + 1), cursor(ctx, nb), bsums(ctx, cdiv(nb, 1024) + 1), entries(ctx, (size_t)n_ * nwin),
        size_bins(ctx, MSM_SIZE_BINS), perm(ctx, nb), n(n_) {}
};
#define KB_LAUNCH_ON(ctx, st, kernel, grid, block, smem, ...) do { \
  kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__); \
  (ctx)->launches++; KB_CUDA(cudaGetLastError()); } while (0)

static void msm_sort(kb_ctx* ctx, cudaStream_t st, MsmSort& w, const MsmTable& tab, int c, const uint32_t* d_scalars) {
  const uint32_t nb = 1u << (c - 1);
  const uint32_t nblk = cdiv(nb, 1024);
  const uint64_t n = w.n;
  KB_CUDA(cudaMemsetAsync(w.counts, 0, nb * sizeof(uint32_t), st));
  KB_LAUNCH_ON(ctx, st, msm_count_kernel, cdiv(n, 256), 256, 0, d_scalars, n, c, tab.nwin, w.counts);
  KB_LAUNCH_ON(ctx, st, scan_local_kernel, nblk, 256, 0, w.counts, w.offsets, w.bsums, nb);
  KB_LAUNCH_ON(ctx, st, scan_sums_kernel, 1, 32, 0, w.bsums, nblk);
  KB_LAUNCH_ON(ctx, st, scan_add_kernel, cdiv(nb + 1, 256), 256, 0, w.offsets, w.bsums, w.cursor, nb, nblk);
  KB_LAUNCH_ON(ctx, st, msm_scatter_kernel, cdiv(n, 256), 256, 0, d_scalars, n, c, tab.nwin, w.cursor, w.entries);
  KB_CUDA(cudaMemsetAsync(w.size_bins, 0, MSM_SIZE_BINS * sizeof(uint32_t), st));
  KB_LAUNCH_ON(ctx, st, msm_size_hist_kernel, cdiv(nb, 256), 256, 0, w.offsets, nb, w.size_bins);
  KB_LAUNCH_ON(ctx, st, msm_size_scan_kernel, 1, MSM_SIZE_BINS, 0, w.size_bins);
  KB_LAUNCH_ON(ctx, st, msm_size_scatter_kernel, cdiv(nb, 256), 256, 0, w.offsets, nb, w.size_bins, w.perm);
}

static void msm_empty(kb_ctx* ctx, uint32_t* d_out_xy, uint8_t* d_out_inf) {
  KB_CUDA(cudaMemsetAsync(d_out_xy, 0, 64, ctx->stream));
  if (d_out_inf) KB_CUDA(cudaMemsetAsync(d_out_inf, 1, 1, ctx->stream));
}

// Scalars at `scalars`: device memory, or host memory when `from_host` (the reference-facing call: `commit` hands over
// a Vec<Fr>).  Large host inputs run as TWO passes over one bucket set (all windows of a point share it, so any split of
// the point range is valid): the first quarter of the points is copied, sorted and accumulated while the side stream
// copies and sorts the rest; the second accumulation adds onto the same buckets; one bucket reduction at the end.
// (Measured: for resident inputs the split does not pay - the second sort only partly hides under the first
// accumulation and two accumulations over half-filled buckets cost what it saves - so they take one pass.)
static void msm_run(kb_ctx* ctx, const uint32_t* scalars, bool from_host, uint64_t first, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf) {
  if (n == 0) return msm_empty(ctx, d_out_xy, d_out_inf);
  if (n > (1ull << 26)) throw ApiError(KB_ERR_ARG, "kb_msm_g1: more than 2^26 points in one call");
  const int c = msm_choose_c(first + n);
  const MsmTable& tab = msm_get_table(ctx, c);
  const uint32_t nb = 1u << (c - 1);
  DevBuf<uint32_t> staged(ctx, from_host ? (size_t)n * 8 : 0);
  const uint32_t* d_sc = from_host ? staged.p : scalars;
  DevBuf<uint32_t> buckets(ctx, 32 * (size_t)nb);
  if (!from_host || n < (1ull << 18)) {   // resident or small inputs: (one copy,) one pass
    if (from_host) KB_CUDA(cudaMemcpyAsync(staged, scalars, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream));
    MsmSort w(ctx, nb, n, tab.nwin);
    msm_sort(ctx, ctx->stream, w, tab, c, d_sc);
    timer_start(ctx, KB_T_MSM_ACC);
    launch_msm_accumulate(ctx, tab.d, tab.n, first, w.offsets, w.entries, w.perm, nb, buckets, false, n * tab.nwin);
    timer_stop(ctx, KB_T_MSM_ACC);
    launch_msm_reduce(ctx, buckets, nb, d_out_xy, d_out_inf);
    return;
  }
  uint64_t na = n / 4;
  if (const char* e = getenv("KB_MSM_FIRST_DIV")) { int d = atoi(e); if (d >= 2 && d <= 64) na = n / d; }   // tuning override
  const uint64_t nbp = n - na;
  MsmSort wa(ctx, nb, na, tab.nwin), wb(ctx, nb, nbp, tab.nwin);
  cudaStream_t side = ctx->copy_stream;
  try {
    KB_CUDA(cudaEventRecord(ctx->ev_copy[0], ctx->stream));      // scratch allocations and the table are ordered on the main stream
    KB_CUDA(cudaStreamWaitEvent(side, ctx->ev_copy[0], 0));
    if (from_host) {
      KB_CUDA(cudaMemcpyAsync(staged, scalars, (size_t)na * 32, cudaMemcpyHostToDevice, side));
      KB_CUDA(cudaEventRecord(ctx->ev_copy[1], side));
      KB_CUDA(cudaMemcpyAsync(staged.p + 8 * na, scalars + 8 * na, (size_t)nbp * 32, cudaMemcpyHostToDevice, side));
      KB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[1], 0));
    }
    msm_sort(ctx, ctx->stream, wa, tab, c, d_sc);                 // main stream: sort of the first pass
    // the second sort is L2-bound like the first (they would only share the L2 if run together) but overlaps well with
    // the multiplier-bound accumulation: gate it behind the first sort; the side stream has the higher priority, so its
    // blocks are dispatched as accumulation blocks retire
    KB_CUDA(cudaEventRecord(ctx->ev_copy[3], ctx->stream));
    KB_CUDA(cudaStreamWaitEvent(side, ctx->ev_copy[3], 0));
    msm_sort(ctx, side, wb, tab, c, d_sc + 8 * na);
    KB_CUDA(cudaEventRecord(ctx->ev_copy[2], side));
    timer_start(ctx, KB_T_MSM_ACC);
    launch_msm_accumulate(ctx, tab.d, tab.n, first, wa.offsets, wa.entries, wa.perm, nb, buckets, false, na * tab.nwin);
    KB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[2], 0));
    launch_msm_accumulate(ctx, tab.d, tab.n, first + na, wb.offsets, wb.entries, wb.perm, nb, buckets, true, nbp * tab.nwin);
    timer_stop(ctx, KB_T_MSM_ACC);
    launch_msm_reduce(ctx, buckets, nb, d_out_xy, d_out_inf);
  } catch (...) {
    cudaStreamSynchronize(side);   // the scratch buffers must outlive the work in flight on the side stream
    throw;
  }
}

void msm_g1(kb_ctx* ctx, const uint32_t* d_scalars, uint64_t first, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf) {
  msm_run(ctx, d_scalars, false, first, n, d_out_xy, d_out_inf);
}
void msm_g1_host(kb_ctx* ctx, const uint32_t* h_scalars, uint64_t first, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf) {
  msm_run(ctx, h_scalars, true, first, n, d_out_xy, d_out_inf);
}

// ------------------------------------------------------------------------------------------
// synthetic SRS: g1[i] = tau^i * G1 (harness; `KZGSetup::setup`, src/kzg.rs:55-70)
// ------------------------------------------------------------------------------------------
// Each thread owns a run of consecutive powers: tau^(t*RUN) by square-and-multiply, then RUN
// scalar multiplications of the generator.
__global__ void __launch_bounds__(128) srs_generate_kernel(const uint32_t* __restrict__ tau_m, uint64_t first_power, uint64_t n, uint32_t* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr tau = fp_load<FrParams>(tau_m);
  // tau^i
  Fr acc = Fr::one(), base = tau;
  for (uint64_t e = first_power + i; e; e >>= 1) { if (e & 1) acc = acc * base; base = sqr(base); }
  Fr k = fp_from_mont<FrParams>(acc);
  G1Affine g;
  g.x = Fq::one();
  g.y = Fq::one() + Fq::one();
  st_g1(out + 16 * i, to_affine(g1_mul_glv(to_xyzz(g), k.v)));
}

__global__ void srs_tau_g2_kernel(const uint32_t* __restrict__ tau_m, const uint32_t* __restrict__ g2_gen, uint32_t* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Fr k = fp_from_mont<FrParams>(fp_load<FrParams>(tau_m));
  G2Affine g = ld_g2(g2_gen);
  st_g2(out, to_affine(ec_mul(to_xyzz(g), k.v)));
}

__global__ void __launch_bounds__(128) g1_mul_gen_kernel(const uint32_t* __restrict__ scalars, uint64_t n, uint32_t* __restrict__ out_xy,
                                                         uint8_t* __restrict__ out_inf) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = fp_from_mont<FrParams>(fp_load<FrParams>(scalars + 8 * i));
  G1Affine g;
  g.x = Fq::one();
  g.y = Fq::one() + Fq::one();
  G1 r = g1_mul_glv(to_xyzz(g), k.v);
  st_g1(out_xy + 16 * i, to_affine(r));
  out_inf[i] = r.is_inf() ? 1 : 0;
}
void g1_mul_gen_batch(kb_ctx* ctx, const uint32_t* d_scalars, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf) {
  if (n) KB_LAUNCH(ctx, g1_mul_gen_kernel, cdiv(n, 128), 128, 0, d_scalars, n, d_out_xy, d_out_inf);
}

void srs_generate(kb_ctx* ctx, const uint32_t* d_tau, uint64_t first_power, uint64_t n, uint32_t* d_tau_g2_out) {
  if (ctx->d_srs) { KB_CUDA(cudaFree(ctx->d_srs)); ctx->d_srs = nullptr; }
  msm_free_tables(ctx);
  KB_CUDA(cudaMalloc((void**)&ctx->d_srs, (size_t)(n ? n : 1) * 64));
  ctx->srs_n = n;
  if (n) KB_LAUNCH(ctx, srs_generate_kernel, cdiv(n, 128), 128, 0, d_tau, first_power, n, ctx->d_srs);
  DevBuf<uint32_t> gen(ctx, 32);
  KB_CUDA(cudaMemcpyAsync(gen, consts::G2_GEN, 128, cudaMemcpyHostToDevice, ctx->stream));
  KB_LAUNCH(ctx, srs_tau_g2_kernel, 1, 32, 0, d_tau, gen, d_tau_g2_out);
}

}  // namespace kb
