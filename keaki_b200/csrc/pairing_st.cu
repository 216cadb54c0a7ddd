// Batched pairing / decapsulation kernel on the compiled single-thread pairing (pairing_st.cuh).
// `decapsulate` (src/kem.rs:55-72) + the XOR of `decrypt` (src/enc.rs:44-55) over the batch of `vec_decrypt`
// (src/vec.rs:72-81), and raw `E::pairing` (src/kem.rs:30,58; src/kzg.rs:148).
//
// One thread per pairing.  F and S (12 Fq2 slots) in shared memory as [slot][quarter][thread] uint4, so every
// access of a warp is 512 contiguous bytes; the G2 accumulator, P, Q and the saved Fq12 values of the final
// exponentiation in a per-thread global scratch with the same layout.  The kernel is persistent: a warp takes 32
// pairings at a time from a global counter, so the scratch is sized by the resident threads and the tail of the
// batch spreads over all SMs (dealing the tasks to the blocks statically was measured: 28.7 ms against 25.5 at 2^16).  GT never leaves the chip on the decrypt path: canonical bytes -> BLAKE3 XOF -> XOR
// happen in the epilogue.
#include "ctx.cuh"
#include "blake3.cuh"
#include "pairing_st.cuh"
#include "consts_gen.cuh"
#include <cstdlib>
#include <string>
#include <vector>

namespace kb {

extern __shared__ uint4 st_smem[];

template <int BLOCK, int NS>   // NS = slot addresses below NS are in shared memory (12, 9 or 6); 6..11 otherwise in scratch slot `address`
struct StDevMem {
  uint4* gl;          // scratch, already offset by the global thread index
  uint32_t gstride;   // threads in the launch

  __device__ __forceinline__ static Fq2 unpack(const uint4& q0, const uint4& q1, const uint4& q2, const uint4& q3) {
    Fq2 r;
    r.c0.v[0] = q0.x; r.c0.v[1] = q0.y; r.c0.v[2] = q0.z; r.c0.v[3] = q0.w;
    r.c0.v[4] = q1.x; r.c0.v[5] = q1.y; r.c0.v[6] = q1.z; r.c0.v[7] = q1.w;
    r.c1.v[0] = q2.x; r.c1.v[1] = q2.y; r.c1.v[2] = q2.z; r.c1.v[3] = q2.w;
    r.c1.v[4] = q3.x; r.c1.v[5] = q3.y; r.c1.v[6] = q3.z; r.c1.v[7] = q3.w;
    return r;
  }
  __device__ __forceinline__ Fq2 ld(int a) const {
    if (a < NS) {
      const uint4* p = st_smem + (size_t)(4 * a) * BLOCK + threadIdx.x;
      return unpack(p[0], p[BLOCK], p[2 * BLOCK], p[3 * BLOCK]);
    }
    const uint4* p = gl + (size_t)(4 * (a < 16 ? a : a - 16)) * gstride;
    return unpack(p[0], p[gstride], p[2 * (size_t)gstride], p[3 * (size_t)gstride]);
  }
  __device__ __forceinline__ void st(int a, const Fq2& x) const {
    const uint4 q0 = make_uint4(x.c0.v[0], x.c0.v[1], x.c0.v[2], x.c0.v[3]), q1 = make_uint4(x.c0.v[4], x.c0.v[5], x.c0.v[6], x.c0.v[7]);
    const uint4 q2 = make_uint4(x.c1.v[0], x.c1.v[1], x.c1.v[2], x.c1.v[3]), q3 = make_uint4(x.c1.v[4], x.c1.v[5], x.c1.v[6], x.c1.v[7]);
    if (a < NS) {
      uint4* p = st_smem + (size_t)(4 * a) * BLOCK + threadIdx.x;
      p[0] = q0; p[BLOCK] = q1; p[2 * BLOCK] = q2; p[3 * BLOCK] = q3;
    } else {
      uint4* p = gl + (size_t)(4 * (a < 16 ? a : a - 16)) * gstride;
      p[0] = q0; p[gstride] = q1; p[2 * (size_t)gstride] = q2; p[3 * (size_t)gstride] = q3;
    }
  }
};

// consts: FROB_GAMMA (18 x 16 limbs) || TW_X || TW_Y
// mode 0: write the 96 canonical GT words; mode 1: key = BLAKE3-XOF(GT bytes), out = key XOR msg_ct; mode 2: the 96
// Montgomery limbs of GT.
template <int BLOCK, int MINB, int NS>
__global__ void __launch_bounds__(BLOCK, MINB) pairing_st_kernel(const uint32_t* __restrict__ consts, const uint32_t* __restrict__ g1,
                                                                 const uint8_t* __restrict__ g1_inf, const uint32_t* __restrict__ g2,
                                                                 const uint8_t* __restrict__ g2_inf, uint64_t n, uint4* __restrict__ scratch,
                                                                 unsigned long long* __restrict__ counter, int mode, uint32_t* __restrict__ gt_out,
                                                                 const uint8_t* __restrict__ msg_ct, const uint64_t* __restrict__ off,
                                                                 uint8_t* __restrict__ out) {
  StDevMem<BLOCK, NS> m;
  m.gstride = gridDim.x * BLOCK;
  m.gl = scratch + (size_t)blockIdx.x * BLOCK + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= n) break;
    const uint64_t i = base + lane;
    const bool live = i < n;
    const uint64_t j = live ? i : n - 1;   // padding lanes recompute the last pairing
    Fq2 P, qx, qy;
    P.c0 = fp_load<FqParams>(g1 + 16 * j); P.c1 = fp_load<FqParams>(g1 + 16 * j + 8);
    qx.c0 = fp_load<FqParams>(g2 + 32 * j); qx.c1 = fp_load<FqParams>(g2 + 32 * j + 8);
    qy.c0 = fp_load<FqParams>(g2 + 32 * j + 16); qy.c1 = fp_load<FqParams>(g2 + 32 * j + 24);
    const bool trivial = (g1_inf && g1_inf[j]) || (g2_inf && g2_inf[j]) || P.is_zero() || (qx.is_zero() && qy.is_zero());
    m.st(st::G_P, P); m.st(st::G_QX, qx); m.st(st::G_QY, qy);
    st::miller(m, consts + 18 * 16);
    st::final_exp(m, consts);
    if (mode == 2) {   // the GT element itself, Montgomery limbs in tower order (per-commitment setup of the encryption tables)
      if (live) {
#pragma unroll 1
        for (int s = 0; s < 6; s++) {
          Fq2 x = m.ld(st::F + s);
          if (trivial) x = s == 0 ? Fq2::one() : Fq2::zero();
          fp_store<FqParams>(gt_out + 96 * i + 16 * s, x.c0);
          fp_store<FqParams>(gt_out + 96 * i + 16 * s + 8, x.c1);
        }
      }
      continue;
    }
    uint32_t w[96];
    st::gt_words(m, w);
    if (trivial) {   // arkworks skips pairs with an infinity: GT = 1
#pragma unroll
      for (int k = 0; k < 96; k++) w[k] = k == 0 ? 1u : 0u;
    }
    if (!live) continue;
    if (mode == 0) {
      uint4* o = reinterpret_cast<uint4*>(gt_out + 96 * i);
#pragma unroll
      for (int k = 0; k < 24; k++) o[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    } else {
      const uint64_t lo = off[i], hi = off[i + 1];
      b3_gt_xof_xor(w, msg_ct + lo, out + lo, hi - lo);
    }
  }
}

// ------------------------------------------------------------------------------------------
// GT side of `encrypt` on the same machinery: secret = prod_w table[w][digit_w] as a chain of Fq12 products with the
// accumulator in shared memory and lazily reduced Fq2 products (18 x 320 IMAD.WIDE per Fq12 product, no stack traffic),
// instead of the generic tower code of we.cu (54 full Montgomery products = 6,912 IMAD.WIDE, Fq12 values on the thread
// stack).  Same table entries in the same order: the same GT element, bit for bit.
// ------------------------------------------------------------------------------------------
__device__ const uint32_t GT_ONE_LIMBS[96] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};

template <int BLOCK>
struct EncMem {          // addresses 0..11: F and S on chip; 16..21: the six Fq2 of the table entry `ent` (read only)
  const uint32_t* ent;
  __device__ __forceinline__ Fq2 ld(int a) const {
    if (a < 12) {
      const uint4* p = st_smem + (size_t)(4 * a) * BLOCK + threadIdx.x;
      return StDevMem<BLOCK, 12>::unpack(p[0], p[BLOCK], p[2 * BLOCK], p[3 * BLOCK]);
    }
    const uint4* p = reinterpret_cast<const uint4*>(ent + 16 * (a - 16));
    return StDevMem<BLOCK, 12>::unpack(p[0], p[1], p[2], p[3]);
  }
  __device__ __forceinline__ void st(int a, const Fq2& x) const {
    uint4* p = st_smem + (size_t)(4 * a) * BLOCK + threadIdx.x;
    p[0] = make_uint4(x.c0.v[0], x.c0.v[1], x.c0.v[2], x.c0.v[3]); p[BLOCK] = make_uint4(x.c0.v[4], x.c0.v[5], x.c0.v[6], x.c0.v[7]);
    p[2 * BLOCK] = make_uint4(x.c1.v[0], x.c1.v[1], x.c1.v[2], x.c1.v[3]); p[3 * BLOCK] = make_uint4(x.c1.v[4], x.c1.v[5], x.c1.v[6], x.c1.v[7]);
  }
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 2) encrypt_gt_st_kernel(const uint32_t* __restrict__ com_tab_a, const uint32_t* __restrict__ com_tab_a1,
                                                                 const uint32_t* __restrict__ gt_tab, const uint32_t* __restrict__ values,
                                                                 const uint32_t* __restrict__ rs, const uint8_t* __restrict__ msgs,
                                                                 const uint64_t* __restrict__ off, uint64_t n, int com_wide, uint8_t* __restrict__ msg_ct) {
  const uint64_t i0 = blockIdx.x * (uint64_t)BLOCK + threadIdx.x;
  const bool live = i0 < n;
  const uint64_t i = live ? i0 : n - 1;              // padding lanes recompute the last message (warps stay whole)
  const Fr r = fp_load<FrParams>(rs + 8 * i), v = fp_load<FrParams>(values + 8 * i);
  const Fr kr = fp_from_mont<FrParams>(r);
  const bool v_one = v == Fr::one(), v_bit = v_one || v.is_zero();
  const uint32_t* com_tab = v_one ? com_tab_a1 : com_tab_a;
  const Fr ks = v_bit ? Fr::zero() : fp_from_mont<FrParams>(-(v * r));
  EncMem<BLOCK> m;
  const int nwin = com_wide ? WE_WIN16 : WE_WIN;
#pragma unroll 1
  for (int w = 0; w < nwin; w++) {
    const uint32_t d = com_wide ? half_of(kr.v, w) : byte_of(kr.v, w);
    m.ent = d ? com_tab + 96 * ((size_t)w * (com_wide ? WE_ENT16 : WE_ENT) + d - 1) : GT_ONE_LIMBS;
    if (w == 0) { for (int k = 0; k < 6; k++) m.st(st::F + k, m.ld(16 + k)); }
    else st::f12mul(m, 16, 0);
  }
  if (__any_sync(0xffffffffu, !v_bit)) {             // general values: times gT^(-v r); bit-valued lanes multiply by one
#pragma unroll 1
    for (int w = 0; w < WE_WIN16; w++) {
      const uint32_t d = half_of(ks.v, w);
      m.ent = d ? gt_tab + 96 * ((size_t)w * WE_ENT16 + d - 1) : GT_ONE_LIMBS;
      st::f12mul(m, 16, 0);
    }
  }
  uint32_t wd[96];
  st::gt_words(m, wd);
  if (!live) return;
  const uint64_t lo = off[i], hi = off[i + 1];
  b3_gt_xof_xor(wd, msgs + lo, msg_ct + lo, hi - lo);
}

void st_encrypt_gt_launch(kb_ctx* ctx, const uint32_t* com_tab_a, const uint32_t* com_tab_a1, const uint32_t* gt_tab16, const uint32_t* d_values,
                          const uint32_t* d_r, const uint8_t* d_msgs, const uint64_t* d_off, uint64_t n, int com_wide, uint8_t* d_msg_ct) {
  constexpr int BLOCK = 128;
  const int smem = BLOCK * 12 * 64;
  static bool prepared[64] = {};
  if (ctx->device >= 64 || !prepared[ctx->device]) {
    KB_CUDA(cudaFuncSetAttribute(encrypt_gt_st_kernel<BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (ctx->device < 64) prepared[ctx->device] = true;
  }
  KB_LAUNCH(ctx, (encrypt_gt_st_kernel<BLOCK>), cdiv(n, BLOCK), BLOCK, smem, com_tab_a, com_tab_a1, gt_tab16, d_values, d_r, d_msgs, d_off, n, com_wide, d_msg_ct);
}

// The SEGMENTED form for batches of more than one round.  A round of resident warps (1,184 on a B200) that all run a
// whole pairing takes 12.9 ms, and a batch is a whole number of such rounds: 2^16 pairings = 2,048 warp tasks = 1.73
// rounds cost two.  The pairing is a resumable step sequence (pairing_st.cuh), so it is cut into K segments of equal cost
// and the K x T (segment, task) units are dealt to rounds of 1,184 in segment-major order: segment s of a task runs in a
// later launch than its segment s - 1, every launch is (nearly) full, and the batch costs ceil(K T / 1184) / K rounds
// instead of ceil(T / 1184).  Between launches F is parked in the scratch, which is indexed by PAIRING here (not by
// resident thread), so the G2 accumulator and the saved powers simply stay where they are.
struct SegBounds { int k; int b[34]; };   // segment s = global steps [b[s], b[s+1]); steps 0..64 Miller loop, 65..257 final exponentiation

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) pairing_seg_kernel(const uint32_t* __restrict__ consts, const uint32_t* __restrict__ g1,
                                                                  const uint8_t* __restrict__ g1_inf, const uint32_t* __restrict__ g2,
                                                                  const uint8_t* __restrict__ g2_inf, uint64_t n, uint32_t tasks, uint4* __restrict__ scratch,
                                                                  uint32_t unit_first, uint32_t unit_count, SegBounds sb, int mode,
                                                                  uint32_t* __restrict__ gt_out, const uint8_t* __restrict__ msg_ct,
                                                                  const uint64_t* __restrict__ off, uint8_t* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31u, w = blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
  if (w >= unit_count) return;                       // whole warps leave
  const uint32_t u = unit_first + w, seg = u / tasks, task = u % tasks;
  const uint64_t i = 32ull * task + lane;
  const bool live = i < n;
  const uint64_t j = live ? i : n - 1;               // padding lanes recompute the last pairing (in their own scratch column)
  StDevMem<BLOCK, 12> m;
  m.gstride = tasks * 32u;
  m.gl = scratch + i;
  const int lo = sb.b[seg], hi = sb.b[seg + 1];
  Fq2 P;
  P.c0 = fp_load<FqParams>(g1 + 16 * j); P.c1 = fp_load<FqParams>(g1 + 16 * j + 8);
  if (lo == 0) {
    Fq2 qx, qy;
    qx.c0 = fp_load<FqParams>(g2 + 32 * j); qx.c1 = fp_load<FqParams>(g2 + 32 * j + 8);
    qy.c0 = fp_load<FqParams>(g2 + 32 * j + 16); qy.c1 = fp_load<FqParams>(g2 + 32 * j + 24);
    m.st(st::G_P, P); m.st(st::G_QX, qx); m.st(st::G_QY, qy);
  } else {
#pragma unroll 1
    for (int s = 0; s < 6; s++) m.st(st::F + s, m.ld(st::G_SAVE_F + s));
  }
  if (lo < st::MILLER_STEPS) st::miller_steps(m, consts + 18 * 16, lo, hi < st::MILLER_STEPS ? hi : st::MILLER_STEPS);
  if (hi > st::MILLER_STEPS) st::final_exp_steps(m, consts, (lo > st::MILLER_STEPS ? lo : st::MILLER_STEPS) - st::MILLER_STEPS, hi - st::MILLER_STEPS);
  if (hi < st::MILLER_STEPS + st::FE_STEPS) {
#pragma unroll 1
    for (int s = 0; s < 6; s++) m.st(st::G_SAVE_F + s, m.ld(st::F + s));
    return;
  }
  const Fq2 qx = m.ld(st::G_QX), qy = m.ld(st::G_QY);
  const bool trivial = (g1_inf && g1_inf[j]) || (g2_inf && g2_inf[j]) || P.is_zero() || (qx.is_zero() && qy.is_zero());
  if (mode == 2) {
    if (live) {
#pragma unroll 1
      for (int s = 0; s < 6; s++) {
        Fq2 x = m.ld(st::F + s);
        if (trivial) x = s == 0 ? Fq2::one() : Fq2::zero();
        fp_store<FqParams>(gt_out + 96 * i + 16 * s, x.c0);
        fp_store<FqParams>(gt_out + 96 * i + 16 * s + 8, x.c1);
      }
    }
    return;
  }
  uint32_t wd[96];
  st::gt_words(m, wd);
  if (trivial) {   // arkworks skips pairs with an infinity: GT = 1
#pragma unroll
    for (int k = 0; k < 96; k++) wd[k] = k == 0 ? 1u : 0u;
  }
  if (!live) return;
  if (mode == 0) {
    uint4* o = reinterpret_cast<uint4*>(gt_out + 96 * i);
#pragma unroll
    for (int k = 0; k < 24; k++) o[k] = make_uint4(wd[4 * k], wd[4 * k + 1], wd[4 * k + 2], wd[4 * k + 3]);
  } else {
    const uint64_t lo_b = off[i], hi_b = off[i + 1];
    b3_gt_xof_xor(wd, msg_ct + lo_b, out + lo_b, hi_b - lo_b);
  }
}

// K segments of (nearly) equal cost.  Costs in Fq2-product equivalents: Fq12 square 12, sparse product 13, tangent 11.4, chord
// 14, cyclotomic square 7.6, Fq12 product 18.
static SegBounds seg_bounds(int k) {
  const int total_steps = st::MILLER_STEPS + st::FE_STEPS;
  std::vector<double> c(total_steps);
  for (int s = 0; s < 64; s++) c[s] = (s == 0 ? 24.4 : 36.4) + (ate_digit(63 - s) != 0 ? 27.0 : 0.0);
  c[64] = 56.0;
  double* f = c.data() + st::MILLER_STEPS;
  for (int s = 0; s < st::FE_STEPS; s++) {
    if (s == 0) f[s] = 90.0;
    else if (s == 64) f[s] = 33.0;
    else if (s == 128) f[s] = 8.0;
    else if (s == 192) f[s] = 195.0;
    else {
      const int kk = (s - 1) % 64;   // position inside the exponentiation
      f[s] = kk == 0 ? 62.0 : 7.6 + (st::z_wnaf(62 - kk) != 0 ? 18.0 : 0.0);
    }
  }
  double sum = 0;
  for (double x : c) sum += x;
  SegBounds sb{};
  if (k > 32) k = 32;
  sb.k = k;
  sb.b[0] = 0;
  double acc = 0;
  int seg = 1;
  for (int s = 0; s < total_steps && seg < k; s++) {
    acc += c[s];
    if (acc >= sum * seg / k) sb.b[seg++] = s + 1;
  }
  while (seg <= k) sb.b[seg++] = total_steps;
  return sb;
}

template <int BLOCK, int MINB>
static void seg_go(kb_ctx* ctx, int k, const uint32_t* d_g1, const uint8_t* d_g1_inf, const uint32_t* d_g2, const uint8_t* d_g2_inf, uint64_t n, int mode,
                   uint32_t* d_gt, const uint8_t* d_msg_ct, const uint64_t* d_off, uint8_t* d_out) {
  const int smem = BLOCK * 12 * 64;
  static bool prepared[64] = {};   // function attributes are per device
  if (ctx->device >= 64 || !prepared[ctx->device]) {
    KB_CUDA(cudaFuncSetAttribute(pairing_seg_kernel<BLOCK, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (ctx->device < 64) prepared[ctx->device] = true;
  }
  const SegBounds sb = seg_bounds(k);
  const uint32_t warps_per_round = (uint32_t)ctx->sm_count * MINB * (BLOCK / 32);
  // the scratch is per pairing: bound it by working through super-chunks of at most 2^17 pairings (0.96 GB), equal in size
  const uint64_t chunks = cdiv(n, 1ull << 17);
  const uint64_t per_chunk = cdiv(cdiv(n, chunks), 32) * 32ull;
  DevBuf<uint4> scratch(ctx, (size_t)(st::SCRATCH_SLOTS + 6) * 4 * per_chunk);
  timer_start(ctx, KB_T_PAIRING);
  for (uint64_t lo = 0; lo < n; lo += per_chunk) {
    const uint64_t cnt = n - lo < per_chunk ? n - lo : per_chunk;
    const uint32_t tasks = cdiv(cnt, 32);
    const uint64_t units = (uint64_t)tasks * sb.k;
    for (uint64_t first = 0; first < units; first += warps_per_round) {
      const uint32_t count = (uint32_t)(units - first < warps_per_round ? units - first : warps_per_round);
      KB_LAUNCH(ctx, (pairing_seg_kernel<BLOCK, MINB>), cdiv(count, BLOCK / 32), BLOCK, smem, ctx->d_st_consts, d_g1 + 16 * lo,
                d_g1_inf ? d_g1_inf + lo : nullptr, d_g2 + 32 * lo, d_g2_inf ? d_g2_inf + lo : nullptr, cnt, tasks, scratch.p, (uint32_t)first, count, sb,
                mode, d_gt ? d_gt + 96 * lo : nullptr, d_msg_ct, d_off ? d_off + lo : nullptr, d_out);
    }
  }
  timer_stop(ctx, KB_T_PAIRING);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int BLOCK, int MINB, int NS>
static void st_go(kb_ctx* ctx, const uint32_t* d_g1, const uint8_t* d_g1_inf, const uint32_t* d_g2, const uint8_t* d_g2_inf, uint64_t n, int mode,
                  uint32_t* d_gt, const uint8_t* d_msg_ct, const uint64_t* d_off, uint8_t* d_out) {
  const int smem = BLOCK * NS * 64;
  static bool prepared[64] = {};   // function attributes are per device
  if (ctx->device >= 64 || !prepared[ctx->device]) {
    KB_CUDA(cudaFuncSetAttribute(pairing_st_kernel<BLOCK, MINB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (ctx->device < 64) prepared[ctx->device] = true;
  }
  // One launch per ROUND (as many pairings as there are resident threads), not one persistent launch for the whole batch:
  // warps that start a round together stay roughly in step and a round of 37,888 pairings takes 12.9 ms, while the warps
  // of a persistent launch drift apart and settle at 15-16 ms per round-equivalent (2^18 pairings: 90 ms in rounds against
  // 112 ms persistent; measured with tools/exp/dec_chunks.py).  Inside a launch the tasks are still dealt dynamically.
  const unsigned full = (unsigned)ctx->sm_count * MINB;
  uint64_t round = (uint64_t)full * BLOCK;
  if (const char* e = getenv("KB_PAIRING_ROUND")) { const uint64_t v = strtoull(e, nullptr, 10); round = v ? v : n; }   // measurement: pairings per launch, 0 = one persistent launch
  const unsigned rounds = cdiv(n, round);
  const unsigned need = cdiv(n, 32) * 32 / BLOCK + 1;   // never more threads than (padded) pairings
  const size_t threads = (size_t)(full < need ? full : need) * BLOCK;
  DevBuf<uint4> scratch(ctx, (size_t)st::SCRATCH_SLOTS * 4 * threads);
  DevBuf<unsigned long long> counter(ctx, rounds);
  KB_CUDA(cudaMemsetAsync(counter.p, 0, rounds * sizeof(unsigned long long), ctx->stream));
  timer_start(ctx, KB_T_PAIRING);
  for (unsigned r = 0; r < rounds; r++) {
    const uint64_t lo = (uint64_t)r * round, cnt = n - lo < round ? n - lo : round;
    unsigned blocks = cdiv(cnt, 32) * 32 / BLOCK + 1;
    if (blocks > full) blocks = full;
    KB_LAUNCH(ctx, (pairing_st_kernel<BLOCK, MINB, NS>), blocks, BLOCK, smem, ctx->d_st_consts, d_g1 + 16 * lo, d_g1_inf ? d_g1_inf + lo : nullptr,
              d_g2 + 32 * lo, d_g2_inf ? d_g2_inf + lo : nullptr, cnt, scratch.p, counter.p + r, mode, d_gt ? d_gt + 96 * lo : nullptr, d_msg_ct,
              d_off ? d_off + lo : nullptr, d_out);
  }
  timer_stop(ctx, KB_T_PAIRING);
}

void st_init(kb_ctx* ctx) {
  KB_CUDA(cudaMalloc((void**)&ctx->d_st_consts, (18 + 2) * 64));
  KB_CUDA(cudaMemcpyAsync(ctx->d_st_consts, consts::FROB_GAMMA, 18 * 64, cudaMemcpyHostToDevice, ctx->stream));
  KB_CUDA(cudaMemcpyAsync(ctx->d_st_consts + 18 * 16, consts::TW_X, 64, cudaMemcpyHostToDevice, ctx->stream));
  KB_CUDA(cudaMemcpyAsync(ctx->d_st_consts + 19 * 16, consts::TW_Y, 64, cudaMemcpyHostToDevice, ctx->stream));
  ctx->st_shape = 0;
  if (const char* e = getenv("KB_PAIRING_ST_SHAPE")) ctx->st_shape = atoi(e);   // tuning override
  ctx->enc_gt_st = true;
  if (const char* e = getenv("KB_ENCRYPT_GT")) ctx->enc_gt_st = std::string(e) != "tower";   // tower = the generic Fq12 code of we.cu
  ctx->st_segments = 12;
  if (const char* e = getenv("KB_PAIRING_SEGMENTS")) ctx->st_segments = atoi(e);   // 0 / 1: whole pairings per launch
}
void st_free(kb_ctx* ctx) { cudaFree(ctx->d_st_consts); ctx->d_st_consts = nullptr; }

void st_pairing_launch(kb_ctx* ctx, const uint32_t* d_g1, const uint8_t* d_g1_inf, const uint32_t* d_g2, const uint8_t* d_g2_inf, uint64_t n,
                       int mode, uint32_t* d_gt, const uint8_t* d_msg_ct, const uint64_t* d_off, uint8_t* d_out) {
  // small batches: one WARP per pairing (pairing_warp.cu) - a lone thread here needs 9 ms however few pairings there are
  if (n <= ctx->wp_max_n) { wp_pairing_launch(ctx, d_g1, d_g1_inf, d_g2, d_g2_inf, n, mode, d_gt, d_msg_ct, d_off, d_out); return; }
  // more than one round of resident warps: the segmented form packs the rounds (default shape only)
  if (ctx->st_shape == 0 && ctx->st_segments > 1 && n > (uint64_t)ctx->sm_count * 2 * 128) {
    seg_go<128, 2>(ctx, ctx->st_segments, d_g1, d_g1_inf, d_g2, d_g2_inf, n, mode, d_gt, d_msg_ct, d_off, d_out);
    return;
  }
#define KB_ST_GO(B, MB, NS) st_go<B, MB, NS>(ctx, d_g1, d_g1_inf, d_g2, d_g2_inf, n, mode, d_gt, d_msg_ct, d_off, d_out)
  // Launch shape: 8 warps per SM at 255 registers (two blocks of 128) is the measured optimum on B200 (2^16 pairings:
  // 25.5 ms; 7 warps in one block 27.1, 6 warps 29.5, 12 warps at 168 registers 38.0, 16 warps at 128 registers 37.1 -
  // fewer registers cost more in spills and lost instruction-level parallelism than the extra warps bring).
  const int shape = ctx->st_shape;
  switch (shape) {   // block, blocks per SM, on-chip slots
    case 5: KB_ST_GO(160, 1, 12); break;   // 5 warps, 255 registers
    case 6: KB_ST_GO(192, 1, 12); break;   // 6 warps
    case 7: KB_ST_GO(224, 1, 12); break;   // 7 warps
    case 12: KB_ST_GO(128, 3, 9); break;   // 12 warps, 168 registers, half of S in the scratch (measured slower: spills)
    default: KB_ST_GO(128, 2, 12); break;  // 8 warps, 255 registers, F and S on chip
  }
#undef KB_ST_GO
}

}  // namespace kb
