// Warp-cooperative pairing kernel (pairing_warp.cuh): one warp per pairing, for batches too small to fill the GPU with
// one-thread-per-pairing work, and for the latency-critical per-commitment setup of kb_encrypt_batch.
// `E::pairing` at src/kem.rs:30,58 / src/kzg.rs:148; the window-base chain is the table setup behind src/kem.rs:31
// (s = e(com, G2)^r as a fixed-base power).
//
// The schedules (pairing_warp_gen.cuh) are expanded once per context into a dense descriptor array in device memory
// (32 B per lane and step, L2-resident: 374 KB for the pairing); the slot file of a warp is [quarter][slot] uint4 in
// shared memory, so lanes that name different slots spread over the banks.
#include "ctx.cuh"
#include "blake3.cuh"
#include "pairing_warp.cuh"
#include "pairing_warp_gen.cuh"
#include <cstdlib>
#include <vector>

namespace kb {

extern __shared__ uint4 wp_smem[];

struct WpDevMem {
  uint4* base;
  uint32_t ns;
  __device__ __forceinline__ Fq2 ld(uint32_t a) const {
    const uint4* p = base + a;
    const uint4 q0 = p[0], q1 = p[ns], q2 = p[2 * ns], q3 = p[3 * ns];
    Fq2 r;
    r.c0.v[0] = q0.x; r.c0.v[1] = q0.y; r.c0.v[2] = q0.z; r.c0.v[3] = q0.w;
    r.c0.v[4] = q1.x; r.c0.v[5] = q1.y; r.c0.v[6] = q1.z; r.c0.v[7] = q1.w;
    r.c1.v[0] = q2.x; r.c1.v[1] = q2.y; r.c1.v[2] = q2.z; r.c1.v[3] = q2.w;
    r.c1.v[4] = q3.x; r.c1.v[5] = q3.y; r.c1.v[6] = q3.z; r.c1.v[7] = q3.w;
    return r;
  }
  __device__ __forceinline__ Fq ld_half(uint32_t a, uint32_t h) const {
    const uint4* p = base + a + 2 * h * ns;
    const uint4 q0 = p[0], q1 = p[ns];
    Fq r;
    r.v[0] = q0.x; r.v[1] = q0.y; r.v[2] = q0.z; r.v[3] = q0.w;
    r.v[4] = q1.x; r.v[5] = q1.y; r.v[6] = q1.z; r.v[7] = q1.w;
    return r;
  }
  __device__ __forceinline__ void st_half(uint32_t a, uint32_t h, const Fq& x) const {
    uint4* p = base + a + 2 * h * ns;
    p[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    p[ns] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  }
  __device__ __forceinline__ void st(uint32_t a, const Fq2& x) const {
    uint4* p = base + a;
    p[0] = make_uint4(x.c0.v[0], x.c0.v[1], x.c0.v[2], x.c0.v[3]);
    p[ns] = make_uint4(x.c0.v[4], x.c0.v[5], x.c0.v[6], x.c0.v[7]);
    p[2 * ns] = make_uint4(x.c1.v[0], x.c1.v[1], x.c1.v[2], x.c1.v[3]);
    p[3 * ns] = make_uint4(x.c1.v[4], x.c1.v[5], x.c1.v[6], x.c1.v[7]);
  }
};

struct WpArgs {
  const uint4* words;        // dense descriptors: 2 uint4 per lane, 64 per step
  const uint32_t* consts;    // wpprog::CONSTS on the device
  const uint16_t* outs;      // output slots (device)
  const uint16_t* ins;       // input slots of GT_PROD (device)
  int nsteps, nslots, nconsts, nouts;
  uint16_t inputs[8];
  uint16_t const_idx[24], const_slot[24];
};

__device__ __forceinline__ void wp_run(const WpDevMem& m, const WpArgs& a, uint32_t lane) {
  const uint4* dp = a.words + 2 * lane;
  uint4 d0 = dp[0], d1 = dp[1];
#pragma unroll 1
  for (int s = 0; s < a.nsteps; s++) {
    const uint32_t d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    if (s + 1 < a.nsteps) { d0 = dp[64 * (size_t)(s + 1)]; d1 = dp[64 * (size_t)(s + 1) + 1]; }   // next step's descriptor in flight
    const Fq2 r = wp::lane_compute(m, d);
    __syncwarp();
    wp::lane_store(m, d, r);
    __syncwarp();
  }
}

__device__ __forceinline__ void wp_prologue(const WpDevMem& m, const WpArgs& a, uint32_t lane) {
  if (lane == 0) m.st(0, Fq2::zero());
  for (int k = lane; k < a.nconsts; k += 32) m.st(a.const_slot[k], st::ld_const2(a.consts + 16 * a.const_idx[k]));
}

// mode 0: the 96 canonical GT words; mode 2: the 96 Montgomery limbs of GT (tower order)
template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS) pairing_warp_kernel(WpArgs a, const uint32_t* __restrict__ g1, const uint8_t* __restrict__ g1_inf,
                                                                  const uint32_t* __restrict__ g2, const uint8_t* __restrict__ g2_inf, uint64_t n,
                                                                  unsigned long long* __restrict__ counter, int mode, uint32_t* __restrict__ gt_out) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  WpDevMem m{wp_smem + (size_t)warp * 4 * a.nslots, (uint32_t)a.nslots};
  wp_prologue(m, a, lane);
  for (;;) {
    unsigned long long i = 0;
    if (lane == 0) i = atomicAdd(counter, 1ull);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n) break;
    Fq2 P, qx, qy;
    P.c0 = fp_load<FqParams>(g1 + 16 * i); P.c1 = fp_load<FqParams>(g1 + 16 * i + 8);
    qx.c0 = fp_load<FqParams>(g2 + 32 * i); qx.c1 = fp_load<FqParams>(g2 + 32 * i + 8);
    qy.c0 = fp_load<FqParams>(g2 + 32 * i + 16); qy.c1 = fp_load<FqParams>(g2 + 32 * i + 24);
    const bool trivial = (g1_inf && g1_inf[i]) || (g2_inf && g2_inf[i]) || P.is_zero() || (qx.is_zero() && qy.is_zero());
    if (lane < 5) {
      Fq2 v = lane == 0 ? P : lane == 1 ? qx : qy;
      if (lane == 3) { v.c0 = P.c0; v.c1 = Fq::zero(); }
      if (lane == 4) { v.c0 = P.c1; v.c1 = Fq::zero(); }
      m.st(a.inputs[lane], v);
    }
    __syncwarp();
    wp_run(m, a, lane);
    if (lane < 6) {
      Fq2 x = m.ld(a.outs[lane]);
      if (trivial) x = lane == 0 ? Fq2::one() : Fq2::zero();   // arkworks skips pairs with an infinity: GT = 1
      if (mode == 0) {   // canonical words (Montgomery product with the integer 1)
        Fq unit = Fq::zero(); unit.v[0] = 1;
        x.c0 = st::f1mul(x.c0, unit); x.c1 = st::f1mul(x.c1, unit);
      }
      fp_store<FqParams>(gt_out + 96 * i + 16 * lane, x.c0);
      fp_store<FqParams>(gt_out + 96 * i + 16 * lane + 8, x.c1);
    }
    __syncwarp();
  }
}

// bases[w] = A^(2^(8 w)), w = 0..31, for a cyclotomic A (96 Montgomery limbs, tower order): one warp, 248 squarings of 9
// products each
__global__ void __launch_bounds__(32) gt_bases_warp_kernel(WpArgs a, const uint32_t* __restrict__ in, uint32_t* __restrict__ bases) {
  const uint32_t lane = threadIdx.x;
  WpDevMem m{wp_smem, (uint32_t)a.nslots};
  wp_prologue(m, a, lane);
  if (lane < 6) m.st(a.inputs[lane], st::ld_const2(in + 16 * lane));
  __syncwarp();
  wp_run(m, a, lane);
  for (int k = lane; k < a.nouts; k += 32) {
    const Fq2 x = m.ld(a.outs[k]);
    fp_store<FqParams>(bases + 16 * k, x.c0);
    fp_store<FqParams>(bases + 16 * k + 8, x.c1);
  }
}

// key = BLAKE3-XOF(GT bytes), out = key XOR msg_ct: the epilogue of `decrypt` (src/enc.rs:44-55) for GT words already in memory
__global__ void __launch_bounds__(64) gt_xof_xor_kernel(const uint32_t* __restrict__ gt_words, uint64_t n, const uint8_t* __restrict__ msg_ct,
                                                        const uint64_t* __restrict__ off, uint8_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[96];
#pragma unroll
  for (int k = 0; k < 96; k++) w[k] = gt_words[96 * i + k];
  const uint64_t lo = off[i], hi = off[i + 1];
  b3_gt_xof_xor(w, msg_ct + lo, out + lo, hi - lo);
}

// Small-batch `encrypt` (src/kem.rs:13-50 + the XOR of src/enc.rs:19-41): ONE WARP per message.  The thread-per-message
// kernels of we.cu multiply 32-48 table entries in sequence (1.9 ms however few messages there are, plus 0.65 ms for the 32
// dependent G2 additions); here the entries are gathered into the slot file and multiplied as a tree by the GT_PROD
// schedule (47 Fq12 products, 89 steps), and the 32 G2 table entries - one per lane - are summed by a shuffle tree.
// Same table lookups, same group elements: identical ciphertexts.
__device__ __forceinline__ G2 g2_shfl_down(const G2& p, int delta) {
  G2 r;
  const Fq2* src[4] = {&p.x, &p.y, &p.zz, &p.zzz};
  Fq2* dst[4] = {&r.x, &r.y, &r.zz, &r.zzz};
#pragma unroll
  for (int c = 0; c < 4; c++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      dst[c]->c0.v[i] = __shfl_down_sync(0xffffffffu, src[c]->c0.v[i], delta);
      dst[c]->c1.v[i] = __shfl_down_sync(0xffffffffu, src[c]->c1.v[i], delta);
    }
  }
  return r;
}

__global__ void __launch_bounds__(32) encrypt_small_kernel(WpArgs a, const uint32_t* __restrict__ com_tab_a, const uint32_t* __restrict__ com_tab_a1,
                                                           int com_wide, const uint32_t* __restrict__ gt_tab, const uint32_t* __restrict__ tau2_tab,
                                                           const uint32_t* __restrict__ g2_tab, const uint32_t* __restrict__ points,
                                                           const uint32_t* __restrict__ values, const uint32_t* __restrict__ rs,
                                                           const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ off, uint64_t n,
                                                           uint32_t* __restrict__ ct, uint8_t* __restrict__ ct_inf, uint8_t* __restrict__ msg_ct) {
  const uint32_t lane = threadIdx.x;
  WpDevMem m{wp_smem, (uint32_t)a.nslots};
  wp_prologue(m, a, lane);
  for (uint64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const Fr r = fp_load<FrParams>(rs + 8 * i), v = fp_load<FrParams>(values + 8 * i), al = fp_load<FrParams>(points + 8 * i);
    const Fr kr = fp_from_mont<FrParams>(r);
    const bool v_one = v == Fr::one(), v_bit = v_one || v.is_zero();
    const uint32_t* com_tab = v_one ? com_tab_a1 : com_tab_a;
    const Fr ks = v_bit ? Fr::zero() : fp_from_mont<FrParams>(-(v * r));
    // ---- GT side: operands 0..31 = the windows of A (or A'), 32..47 = those of gT; digit 0 or an unused operand = 1
    for (int p = lane; p < 48 * 6; p += 32) {
      const int el = p / 6, comp = p % 6;
      const uint32_t* src = nullptr;
      if (el < 32) {
        if (com_wide) { if (el < WE_WIN16) { const uint32_t d = half_of(kr.v, el); if (d) src = com_tab + 96 * ((size_t)el * WE_ENT16 + d - 1); } }
        else { const uint32_t d = byte_of(kr.v, el); if (d) src = com_tab + 96 * ((size_t)el * WE_ENT + d - 1); }
      } else {
        const uint32_t d = half_of(ks.v, el - 32);
        if (d) src = gt_tab + 96 * ((size_t)(el - 32) * WE_ENT16 + d - 1);
      }
      Fq2 x = comp == 0 ? Fq2::one() : Fq2::zero();
      if (src) x = st::ld_const2(src + 16 * comp);
      m.st(a.ins[p], x);
    }
    __syncwarp();
    wp_run(m, a, lane);
    if (lane < 6) {   // canonical words, staged in the slot file for the hashing lane
      Fq2 x = m.ld(a.outs[lane]);
      Fq unit = Fq::zero(); unit.v[0] = 1;
      x.c0 = st::f1mul(x.c0, unit); x.c1 = st::f1mul(x.c1, unit);
      m.st(a.outs[lane], x);
    }
    __syncwarp();
    if (lane == 0) {
      uint32_t w[96];
#pragma unroll
      for (int s = 0; s < 6; s++) {
        const Fq2 x = m.ld(a.outs[s]);
#pragma unroll
        for (int k = 0; k < 8; k++) { w[16 * s + k] = x.c0.v[k]; w[16 * s + 8 + k] = x.c1.v[k]; }
      }
      const uint64_t lo = off[i], hi = off[i + 1];
      b3_gt_xof_xor(w, msgs + lo, msg_ct + lo, hi - lo);
    }
    __syncwarp();
    // ---- G2 side: ct = r tau_2 - (r alpha) G2; lane l holds window l >> 1 of tau_2 (even lanes) or of -G2 (odd lanes)
    {
      const Fr ka = fp_from_mont<FrParams>(r * al);
      const uint32_t w = lane >> 1;
      const uint32_t d = half_of((lane & 1) ? ka.v : kr.v, w);
      G2 acc = G2::infinity();
      if (d) {
        G2Affine t = ld_g2(((lane & 1) ? g2_tab : tau2_tab) + 32 * ((size_t)w * WE_ENT16 + d - 1));
        if (lane & 1) t.y = -t.y;
        acc = to_xyzz(t);
      }
#pragma unroll 1
      for (int delta = 16; delta >= 1; delta >>= 1) acc = ec_add(acc, g2_shfl_down(acc, delta));
      if (lane == 0) {
        st_g2(ct + 32 * i, to_affine(acc));
        ct_inf[i] = acc.is_inf() ? 1 : 0;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static WpArgs wp_args(kb_ctx* ctx, int which) {
  const wpprog::Program& p = which == 0 ? wpprog::PAIRING : which == 1 ? wpprog::GT_BASES : wpprog::GT_PROD;
  WpArgs a{};
  a.words = reinterpret_cast<const uint4*>(ctx->d_wp_words[which]);
  a.consts = ctx->d_wp_consts;
  a.outs = ctx->d_wp_outs[which];
  a.nsteps = p.nsteps; a.nslots = p.nslots; a.nconsts = p.nconsts; a.nouts = p.nouts;
  for (int i = 0; i < p.ninputs && i < 8; i++) a.inputs[i] = p.inputs[i];   // GT_PROD's 288 input slots are read from d_wp_ins
  a.ins = ctx->d_wp_ins;
  for (int i = 0; i < p.nconsts; i++) { a.const_idx[i] = p.const_idx[i]; a.const_slot[i] = p.const_slot[i]; }
  return a;
}

void wp_init(kb_ctx* ctx) {
  static_assert(wpprog::NUM_CONSTS <= 24, "constant table of WpArgs");
  KB_CUDA(cudaMalloc((void**)&ctx->d_wp_consts, sizeof(wpprog::CONSTS)));
  KB_CUDA(cudaMemcpyAsync(ctx->d_wp_consts, wpprog::CONSTS, sizeof(wpprog::CONSTS), cudaMemcpyHostToDevice, ctx->stream));
  for (int which = 0; which < 3; which++) {
    const wpprog::Program& p = which == 0 ? wpprog::PAIRING : which == 1 ? wpprog::GT_BASES : wpprog::GT_PROD;
    std::vector<uint32_t> words((size_t)p.nsteps * 32 * 8);
    wpprog::expand(p, words.data());
    KB_CUDA(cudaMalloc((void**)&ctx->d_wp_words[which], words.size() * 4));
    KB_CUDA(cudaMemcpy(ctx->d_wp_words[which], words.data(), words.size() * 4, cudaMemcpyHostToDevice));
    KB_CUDA(cudaMalloc((void**)&ctx->d_wp_outs[which], p.nouts * 2));
    KB_CUDA(cudaMemcpy(ctx->d_wp_outs[which], p.outs, p.nouts * 2, cudaMemcpyHostToDevice));
  }
  KB_CUDA(cudaMalloc((void**)&ctx->d_wp_ins, wpprog::GT_PROD.ninputs * 2));
  KB_CUDA(cudaMemcpy(ctx->d_wp_ins, wpprog::GT_PROD.inputs, wpprog::GT_PROD.ninputs * 2, cudaMemcpyHostToDevice));
  // Batches up to this many pairings take the warp-cooperative kernel: measured crossover with the one-thread-per-pairing
  // kernel, whose floor is a lone warp's 9.1 ms (4096 pairings: 6.0 ms here, 8192: 11.4)
  ctx->wp_max_n = 6144;
  if (const char* e = getenv("KB_PAIRING_WARP_MAX")) ctx->wp_max_n = strtoull(e, nullptr, 10);
  ctx->wp_enc_max_n = 2048;   // one warp per message below this (measured crossover with the thread-per-message kernels: DESIGN.md 4.2)
  if (const char* e = getenv("KB_ENCRYPT_WARP_MAX")) ctx->wp_enc_max_n = strtoull(e, nullptr, 10);
  if (const char* e = getenv("KB_NTT_RADIX2")) ctx->ntt_radix2 = atoi(e) != 0;
}
void wp_free(kb_ctx* ctx) {
  cudaFree(ctx->d_wp_ins); ctx->d_wp_ins = nullptr;
  for (int which = 0; which < 3; which++) {
    cudaFree(ctx->d_wp_words[which]); ctx->d_wp_words[which] = nullptr;
    cudaFree(ctx->d_wp_outs[which]); ctx->d_wp_outs[which] = nullptr;
  }
  cudaFree(ctx->d_wp_consts); ctx->d_wp_consts = nullptr;
}

template <int WARPS>
static void wp_go(kb_ctx* ctx, const WpArgs& a, const uint32_t* d_g1, const uint8_t* d_g1_inf, const uint32_t* d_g2, const uint8_t* d_g2_inf, uint64_t n,
                  int mode, uint32_t* d_gt, unsigned long long* counter) {
  const int smem = WARPS * a.nslots * 64;
  static bool prepared[64] = {};   // function attributes are per device
  if (ctx->device >= 64 || !prepared[ctx->device]) {
    KB_CUDA(cudaFuncSetAttribute(pairing_warp_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (ctx->device < 64) prepared[ctx->device] = true;
  }
  int per_sm = 0;
  KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pairing_warp_kernel<WARPS>, 32 * WARPS, smem));
  uint64_t blocks = (uint64_t)ctx->sm_count * (per_sm > 0 ? per_sm : 1);
  if (blocks > cdiv(n, WARPS)) blocks = cdiv(n, WARPS);
  KB_LAUNCH(ctx, (pairing_warp_kernel<WARPS>), (unsigned)blocks, 32 * WARPS, smem, a, d_g1, d_g1_inf, d_g2, d_g2_inf, n, counter, mode, d_gt);
}

void wp_pairing_launch(kb_ctx* ctx, const uint32_t* d_g1, const uint8_t* d_g1_inf, const uint32_t* d_g2, const uint8_t* d_g2_inf, uint64_t n,
                       int mode, uint32_t* d_gt, const uint8_t* d_msg_ct, const uint64_t* d_off, uint8_t* d_out) {
  const WpArgs a = wp_args(ctx, 0);
  DevBuf<unsigned long long> counter(ctx, 1);
  KB_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), ctx->stream));
  DevBuf<uint32_t> words(ctx, mode == 1 ? 96 * n : 0);
  uint32_t* gt = mode == 1 ? words.p : d_gt;
  timer_start(ctx, KB_T_PAIRING);
  // few pairings: one warp per block, so that they spread over the SMs (and over their four schedulers) first
  if (n <= (uint64_t)ctx->sm_count * 4) wp_go<1>(ctx, a, d_g1, d_g1_inf, d_g2, d_g2_inf, n, mode == 1 ? 0 : mode, gt, counter.p);
  else wp_go<4>(ctx, a, d_g1, d_g1_inf, d_g2, d_g2_inf, n, mode == 1 ? 0 : mode, gt, counter.p);
  if (mode == 1) KB_LAUNCH(ctx, gt_xof_xor_kernel, cdiv(n, 64), 64, 0, gt, n, d_msg_ct, d_off, d_out);
  timer_stop(ctx, KB_T_PAIRING);
}

void wp_gt_bases_launch(kb_ctx* ctx, const uint32_t* d_a, uint32_t* d_bases) {
  const WpArgs a = wp_args(ctx, 1);
  const int smem = a.nslots * 64;
  static bool prepared[64] = {};
  if (ctx->device >= 64 || !prepared[ctx->device]) {
    KB_CUDA(cudaFuncSetAttribute(gt_bases_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (ctx->device < 64) prepared[ctx->device] = true;
  }
  KB_LAUNCH(ctx, gt_bases_warp_kernel, 1, 32, smem, a, d_a, d_bases);
}

void wp_encrypt_small_launch(kb_ctx* ctx, const uint32_t* com_tab_a, const uint32_t* com_tab_a1, int com_wide, const uint32_t* gt_tab16,
                             const uint32_t* tau2_tab16, const uint32_t* g2_tab16, const uint32_t* d_points, const uint32_t* d_values,
                             const uint32_t* d_r, const uint8_t* d_msgs, const uint64_t* d_off, uint64_t n, uint32_t* d_ct, uint8_t* d_ct_inf,
                             uint8_t* d_msg_ct) {
  const WpArgs a = wp_args(ctx, 2);
  const int smem = a.nslots * 64;
  static bool prepared[64] = {};
  if (ctx->device >= 64 || !prepared[ctx->device]) {
    KB_CUDA(cudaFuncSetAttribute(encrypt_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (ctx->device < 64) prepared[ctx->device] = true;
  }
  int per_sm = 0;
  KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encrypt_small_kernel, 32, smem));
  uint64_t blocks = (uint64_t)ctx->sm_count * (per_sm > 0 ? per_sm : 1);
  if (blocks > n) blocks = n;
  KB_LAUNCH(ctx, encrypt_small_kernel, (unsigned)blocks, 32, smem, a, com_tab_a, com_tab_a1, com_wide, gt_tab16, tau2_tab16, g2_tab16, d_points, d_values,
            d_r, d_msgs, d_off, n, d_ct, d_ct_inf, d_msg_ct);
}

}  // namespace kb
