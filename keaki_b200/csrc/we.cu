// Batched witness encryption: `encapsulate` / `decapsulate` (src/kem.rs:13-72), the XOR layer of
// `encrypt` / `decrypt` (src/enc.rs:19-55), `E::pairing` and `verify` (src/kzg.rs:127-151).
// One thread per message / pairing (the batch dimension of src/vec.rs:63,75).
//
// Decapsulation is a full variable-base pairing per message.  Encapsulation is restructured
// without changing any output bit (SURVEY.md §8d allows this):
//   s_i  = e(r_i (C - v_i G1), G2) = A^{r_i} * gT^{-v_i r_i},  A = e(C, G2) (one pairing per
//          commitment, cached), gT = e(G1, G2);  both are fixed-base GT exponentiations served from
//          window tables (8-bit windows for A, 16-bit for gT), i.e. <= 48 Fq12 products per message;
//   ct_i = r_i (tau_2 - a_i G2) = r_i tau_2 - (r_i a_i) G2: two fixed-base G2 multiplications from
//          16-bit-window affine tables, i.e. <= 32 mixed additions per message.
// GT bytes -> BLAKE3 XOF -> XOR with the message happen in the same kernel; GT never leaves chip.
#include "ctx.cuh"
#include "blake3.cuh"
#include "consts_gen.cuh"

namespace kb {

__constant__ PairingConsts c_pc;

void we_upload_consts() {
  PairingConsts pc;
  static_assert(sizeof(PairingConsts) == 20 * 64, "PairingConsts layout");
  memcpy(&pc.frob, consts::FROB_GAMMA, sizeof(pc.frob));
  memcpy(&pc.tw_x, consts::TW_X, 64);
  memcpy(&pc.tw_y, consts::TW_Y, 64);
  KB_CUDA(cudaMemcpyToSymbol(c_pc, &pc, sizeof(pc)));
}

__device__ __forceinline__ G1Affine load_g1_flag(const uint32_t* xy, const uint8_t* inf, uint64_t i) {
  if (inf && inf[i]) return G1Affine::infinity();
  return ld_g1(xy + 16 * i);
}
__device__ __forceinline__ G2Affine load_g2_flag(const uint32_t* xy, const uint8_t* inf, uint64_t i) {
  if (inf && inf[i]) return G2Affine::infinity();
  return ld_g2(xy + 32 * i);
}

// ------------------------------------------------------------------------------------------
// fixed-base tables
// ------------------------------------------------------------------------------------------
// bases[w] = 2^(8w) * B (XYZZ), single thread (one-time per base point)
__global__ void g2_window_bases_kernel(const uint32_t* __restrict__ base_xy, uint32_t* __restrict__ bases /* 32 * 64 limbs */) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  G2 acc = to_xyzz(ld_g2(base_xy));
  for (int w = 0; w < WE_WIN; w++) {
    st_fq2(bases + 64 * w, acc.x); st_fq2(bases + 64 * w + 16, acc.y);
    st_fq2(bases + 64 * w + 32, acc.zz); st_fq2(bases + 64 * w + 48, acc.zzz);
    if (w + 1 < WE_WIN) for (int k = 0; k < 8; k++) acc = ec_dbl(acc);
  }
}
// tab[w][d-1] = d * bases[w], affine
__global__ void __launch_bounds__(128) g2_table_fill_kernel(const uint32_t* __restrict__ bases, uint32_t* __restrict__ tab) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= WE_WIN * WE_ENT) return;
  uint32_t w = t / WE_ENT, d = t % WE_ENT + 1;
  G2 b;
  b.x = ld_fq2(bases + 64 * w); b.y = ld_fq2(bases + 64 * w + 16); b.zz = ld_fq2(bases + 64 * w + 32); b.zzz = ld_fq2(bases + 64 * w + 48);
  G2 acc = G2::infinity();
  for (int bit = 7; bit >= 0; bit--) {
    acc = ec_dbl(acc);
    if ((d >> bit) & 1u) acc = ec_add(acc, b);
  }
  st_g2(tab + 32 * (size_t)t, to_affine(acc));
}

// tab16[w][d-1] = d * 2^(16w) * B = tab8[2w+1][d >> 8] + tab8[2w][d & 255], affine
__global__ void __launch_bounds__(128) g2_table16_kernel(const uint32_t* __restrict__ tab8, uint32_t* __restrict__ tab16) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint32_t)WE_WIN16 * WE_ENT16) return;
  uint32_t w = t / WE_ENT16, d = t % WE_ENT16 + 1, hi = d >> 8, lo = d & 255u;
  G2 acc = hi ? to_xyzz(ld_g2(tab8 + 32 * ((size_t)(2 * w + 1) * WE_ENT + hi - 1))) : G2::infinity();
  if (lo) acc = ec_add_mixed(acc, ld_g2(tab8 + 32 * ((size_t)(2 * w) * WE_ENT + lo - 1)));
  st_g2(tab16 + 32 * (size_t)t, to_affine(acc));
}
// tab16[w][d-1] = base^(d 2^(16w)) = tab8[2w+1][d >> 8] * tab8[2w][d & 255]
__global__ void __launch_bounds__(128) gt_table16_kernel(const uint32_t* __restrict__ tab8, uint32_t* __restrict__ tab16) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint32_t)WE_WIN16 * WE_ENT16) return;
  uint32_t w = t / WE_ENT16, d = t % WE_ENT16 + 1, hi = d >> 8, lo = d & 255u;
  Fq12 acc;
  if (hi) {
    acc = ld_fq12(tab8 + 96 * ((size_t)(2 * w + 1) * WE_ENT + hi - 1));
    if (lo) acc = acc * ld_fq12(tab8 + 96 * ((size_t)(2 * w) * WE_ENT + lo - 1));
  } else {
    acc = ld_fq12(tab8 + 96 * ((size_t)(2 * w) * WE_ENT + lo - 1));
  }
  st_fq12(tab16 + 96 * (size_t)t, acc);
}

// bases[w] = A^(2^(8w)), single thread; A must be in GT (cyclotomic)
__global__ void gt_window_bases_kernel(const uint32_t* __restrict__ a, uint32_t* __restrict__ bases /* 32 * 96 */) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Fq12 acc = ld_fq12(a);
  for (int w = 0; w < WE_WIN; w++) {
    st_fq12(bases + 96 * w, acc);
    if (w + 1 < WE_WIN) for (int k = 0; k < 8; k++) acc = cyclotomic_sqr(acc);
  }
}
// tab[w][d-1] = bases[w]^d
__global__ void __launch_bounds__(128) gt_table_fill_kernel(const uint32_t* __restrict__ bases, uint32_t* __restrict__ tab) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= WE_WIN * WE_ENT) return;
  uint32_t w = t / WE_ENT, d = t % WE_ENT + 1;
  Fq12 b = ld_fq12(bases + 96 * w);
  Fq12 acc = b;
  int top = 31 - __clz(d);
  for (int bit = top - 1; bit >= 0; bit--) {
    acc = cyclotomic_sqr(acc);
    if ((d >> bit) & 1u) acc = acc * b;
  }
  st_fq12(tab + 96 * (size_t)t, acc);
}

// both per-commitment tables in one launch: A (bases) and A' = A / gT, whose window bases are bases[w] * conj(gT^(2^(8w)))
// (entry (w, digit 1) of the SRS-constant 8-bit gT table; gT is unitary, so the conjugate is the inverse) - 32 independent
// Fq12 products instead of a second chain of 248 dependent squarings, folded into the fill
__global__ void __launch_bounds__(128) gt_table_fill2_kernel(const uint32_t* __restrict__ bases, const uint32_t* __restrict__ gt_tab,
                                                             uint32_t* __restrict__ tab_a, uint32_t* __restrict__ tab_a1) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * WE_WIN * WE_ENT) return;
  const bool second = t >= WE_WIN * WE_ENT;
  if (second) t -= WE_WIN * WE_ENT;
  const uint32_t w = t / WE_ENT, d = t % WE_ENT + 1;
  Fq12 b = ld_fq12(bases + 96 * w);
  if (second) b = b * conj(ld_fq12(gt_tab + 96 * ((size_t)w * WE_ENT)));
  Fq12 acc = b;
  const int top = 31 - __clz(d);
  for (int bit = top - 1; bit >= 0; bit--) {
    acc = cyclotomic_sqr(acc);
    if ((d >> bit) & 1u) acc = acc * b;
  }
  st_fq12((second ? tab_a1 : tab_a) + 96 * (size_t)t, acc);
}

static void build_g2_table(kb_ctx* ctx, const uint32_t* d_base_xy, uint32_t* d_tab) {
  DevBuf<uint32_t> bases(ctx, WE_WIN * 64);
  KB_LAUNCH(ctx, g2_window_bases_kernel, 1, 32, 0, d_base_xy, bases);
  KB_LAUNCH(ctx, g2_table_fill_kernel, cdiv(WE_WIN * WE_ENT, 128), 128, 0, bases, d_tab);
}
static void build_g2_table16(kb_ctx* ctx, const uint32_t* d_tab8, uint32_t* d_tab16) {
  KB_LAUNCH(ctx, g2_table16_kernel, cdiv((uint64_t)WE_WIN16 * WE_ENT16, 128), 128, 0, d_tab8, d_tab16);
}
static void build_gt_table(kb_ctx* ctx, const uint32_t* d_a, uint32_t* d_tab) {
  DevBuf<uint32_t> bases(ctx, WE_WIN * 96);
  if (ctx->wp_max_n) wp_gt_bases_launch(ctx, d_a, bases);   // the 248 dependent squarings, 9 products wide on one warp
  else KB_LAUNCH(ctx, gt_window_bases_kernel, 1, 32, 0, d_a, bases);
  KB_LAUNCH(ctx, gt_table_fill_kernel, cdiv(WE_WIN * WE_ENT, 128), 128, 0, bases, d_tab);
}
void we_init_tables(kb_ctx* ctx) {
  KB_CUDA(cudaMalloc((void**)&ctx->d_g2_tab, G2_TAB_LIMBS * 4));
  KB_CUDA(cudaMalloc((void**)&ctx->d_tau2_tab, G2_TAB_LIMBS * 4));
  KB_CUDA(cudaMalloc((void**)&ctx->d_gt_tab, GT_TAB_LIMBS * 4));
  KB_CUDA(cudaMalloc((void**)&ctx->d_com_tab, GT_TAB_LIMBS * 4));
  KB_CUDA(cudaMalloc((void**)&ctx->d_com1_tab, GT_TAB_LIMBS * 4));
  DevBuf<uint32_t> g2(ctx, 32), g1(ctx, 16), gt(ctx, 96);
  KB_CUDA(cudaMemcpyAsync(g2, consts::G2_GEN, 128, cudaMemcpyHostToDevice, ctx->stream));
  KB_CUDA(cudaMemcpyAsync(g1, consts::G1_GEN, 64, cudaMemcpyHostToDevice, ctx->stream));
  KB_CUDA(cudaMalloc((void**)&ctx->d_g2_tab16, G2_TAB16_LIMBS * 4));
  KB_CUDA(cudaMalloc((void**)&ctx->d_tau2_tab16, G2_TAB16_LIMBS * 4));
  KB_CUDA(cudaMalloc((void**)&ctx->d_gt_tab16, GT_TAB16_LIMBS * 4));
  build_g2_table(ctx, g2, ctx->d_g2_tab);
  build_g2_table16(ctx, ctx->d_g2_tab, ctx->d_g2_tab16);
  st_pairing_launch(ctx, g1, nullptr, g2, nullptr, 1, 2, gt, nullptr, nullptr, nullptr);
  build_gt_table(ctx, gt, ctx->d_gt_tab);
  KB_LAUNCH(ctx, gt_table16_kernel, cdiv((uint64_t)WE_WIN16 * WE_ENT16, 128), 128, 0, ctx->d_gt_tab, ctx->d_gt_tab16);
  KB_CUDA(cudaStreamSynchronize(ctx->stream));
}

void we_set_tau2(kb_ctx* ctx, const uint32_t* d_tau2) {
  build_g2_table(ctx, d_tau2, ctx->d_tau2_tab);
  build_g2_table16(ctx, ctx->d_tau2_tab, ctx->d_tau2_tab16);
}

void we_free(kb_ctx* ctx) {
  cudaFree(ctx->d_g2_tab); cudaFree(ctx->d_tau2_tab); cudaFree(ctx->d_gt_tab); cudaFree(ctx->d_com_tab);
  cudaFree(ctx->d_g2_tab16); cudaFree(ctx->d_tau2_tab16); cudaFree(ctx->d_gt_tab16);
  if (ctx->d_com_tab16) cudaFree(ctx->d_com_tab16);
  if (ctx->d_com1_tab16) cudaFree(ctx->d_com1_tab16);
  cudaFree(ctx->d_com1_tab);
  ctx->d_com1_tab = ctx->d_com1_tab16 = nullptr;
  ctx->d_com_tab16 = nullptr; ctx->com_tab16_valid = false;
  ctx->d_g2_tab = ctx->d_tau2_tab = ctx->d_gt_tab = ctx->d_com_tab = nullptr;
  ctx->d_g2_tab16 = ctx->d_tau2_tab16 = ctx->d_gt_tab16 = nullptr;
}

// ------------------------------------------------------------------------------------------
// encrypt
// ------------------------------------------------------------------------------------------

// Two kernels per batch: the GT side (two fixed-base exponentiations, hash, XOR) keeps an Fq12 accumulator and a table
// entry live (255 registers); the G2 side (two fixed-base multiplications, one affine normalisation) needs a third of
// that, so it runs at three times the occupancy in a kernel of its own.
__global__ void __launch_bounds__(128) encrypt_kernel(const uint32_t* __restrict__ com_tab_a, const uint32_t* __restrict__ com_tab_a1,
                                                      const uint32_t* __restrict__ gt_tab, const uint32_t* __restrict__ values, const uint32_t* __restrict__ rs,
                                                      const uint8_t* __restrict__ msgs, const uint64_t* __restrict__ off, uint64_t n, int com_wide,
                                                      uint8_t* __restrict__ msg_ct) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr r = fp_load<FrParams>(rs + 8 * i);
  Fr v = fp_load<FrParams>(values + 8 * i);
  Fr kr = fp_from_mont<FrParams>(r);             // r
  // value = 0: secret = A^r; value = 1: secret = (A / gT)^r from the second per-commitment table (the values of a
  // laconic-OT sender are bits, tests/laconic_ot.rs:89-109); any other value takes the general form below
  const bool v_one = v == Fr::one(), v_bit = v_one || v.is_zero();
  const uint32_t* com_tab = v_one ? com_tab_a1 : com_tab_a;
  Fr ks = v_bit ? Fr::zero() : fp_from_mont<FrParams>(-(v * r));      // -v r

  // secret = A^r * gT^(-v r): 8- or 16-bit windows for A (per-commitment table), 16-bit windows for gT
  Fq12 s = Fq12::one();
  bool started = false;
  if (com_wide) {
    for (int w = 0; w < WE_WIN16; w++) {
      uint32_t d = half_of(kr.v, w);
      if (d) { Fq12 t = ld_fq12(com_tab + 96 * ((size_t)w * WE_ENT16 + d - 1)); s = started ? s * t : t; started = true; }
    }
  } else {
    for (int w = 0; w < WE_WIN; w++) {
      uint32_t d = byte_of(kr.v, w);
      if (d) { Fq12 t = ld_fq12(com_tab + 96 * ((size_t)w * WE_ENT + d - 1)); s = started ? s * t : t; started = true; }
    }
  }
  for (int w = 0; w < WE_WIN16; w++) {
    uint32_t d = half_of(ks.v, w);
    if (d) { Fq12 t = ld_fq12(gt_tab + 96 * ((size_t)w * WE_ENT16 + d - 1)); s = started ? s * t : t; started = true; }
  }
  uint32_t words[96];
  gt_to_words(s, words);
  uint64_t lo = off[i], hi = off[i + 1];
  b3_gt_xof_xor(words, msgs + lo, msg_ct + lo, hi - lo);
}

// ct = r tau_2 - (r alpha) G2
__global__ void __launch_bounds__(128, 4) encrypt_ct_kernel(const uint32_t* __restrict__ tau2_tab, const uint32_t* __restrict__ g2_tab,
                                                            const uint32_t* __restrict__ points, const uint32_t* __restrict__ rs, uint64_t n,
                                                            uint32_t* __restrict__ ct, uint8_t* __restrict__ ct_inf) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr r = fp_load<FrParams>(rs + 8 * i);
  Fr a = fp_load<FrParams>(points + 8 * i);
  Fr kr = fp_from_mont<FrParams>(r);             // r
  Fr ka = fp_from_mont<FrParams>(r * a);         // r alpha
  G2 acc = G2::infinity();
  for (int w = 0; w < WE_WIN16; w++) {
    uint32_t d = half_of(kr.v, w);
    if (d) acc = ec_add_mixed(acc, ld_g2(tau2_tab + 32 * ((size_t)w * WE_ENT16 + d - 1)));
    d = half_of(ka.v, w);
    if (d) { G2Affine t = ld_g2(g2_tab + 32 * ((size_t)w * WE_ENT16 + d - 1)); t.y = -t.y; acc = ec_add_mixed(acc, t); }
  }
  st_g2(ct + 32 * i, to_affine(acc));
  ct_inf[i] = acc.is_inf() ? 1 : 0;
}

void encrypt_batch(kb_ctx* ctx, const uint32_t* h_com_xy, uint8_t com_inf, const uint32_t* d_points, const uint32_t* d_values,
                   const uint32_t* d_r, const uint8_t* d_msgs, const uint64_t* d_off, uint64_t n,
                   uint32_t* d_ct, uint8_t* d_ct_inf, uint8_t* d_msg_ct) {
  // (re)build the A = e(com, G2) table when the commitment changes
  uint32_t key[17];
  memcpy(key, h_com_xy, 64);
  key[16] = com_inf ? 1u : 0u;
  if (com_inf) memset(key, 0, 64);
  if (!ctx->com_tab_valid || memcmp(key, ctx->com_cached, sizeof(key)) != 0) {
    DevBuf<uint32_t> com(ctx, 16), g2(ctx, 32), a(ctx, 96);
    KB_CUDA(cudaMemcpyAsync(com, key, 64, cudaMemcpyHostToDevice, ctx->stream));
    KB_CUDA(cudaMemcpyAsync(g2, consts::G2_GEN, 128, cudaMemcpyHostToDevice, ctx->stream));
    // A = e(com, G2): one pairing (warp-cooperative kernel: 1.4 ms; a lone thread of the batch kernel needs 9), the 248
    // dependent squarings of its window bases on one warp, then both power tables in one launch
    timer_start(ctx, KB_T_SETUP);
    st_pairing_launch(ctx, com, nullptr, g2, nullptr, 1, 2, a, nullptr, nullptr, nullptr);
    DevBuf<uint32_t> bases(ctx, WE_WIN * 96);
    if (ctx->wp_max_n) wp_gt_bases_launch(ctx, a, bases.p);
    else KB_LAUNCH(ctx, gt_window_bases_kernel, 1, 32, 0, a.p, bases.p);
    KB_LAUNCH(ctx, gt_table_fill2_kernel, cdiv(2 * WE_WIN * WE_ENT, 128), 128, 0, bases.p, ctx->d_gt_tab, ctx->d_com_tab, ctx->d_com1_tab);
    timer_stop(ctx, KB_T_SETUP);
    memcpy(ctx->com_cached, key, sizeof(key));
    ctx->com_tab_valid = true;
    ctx->com_tab16_valid = false;
    ctx->com_msgs = 0;
  }
  if (!n) return;
  // 16-bit tables halve the products per message and cost 2^20 products each to build: they are bought once the messages
  // already encrypted under this commitment would have paid for them (2 * 2^20 / 32 = 2^16: the ski-rental rule, at most
  // twice the cost of knowing the future)
  if (!ctx->com_tab16_valid && ctx->com_msgs >= (1ull << 16)) {
    if (!ctx->d_com_tab16) KB_CUDA(cudaMalloc((void**)&ctx->d_com_tab16, GT_TAB16_LIMBS * 4));
    KB_LAUNCH(ctx, gt_table16_kernel, cdiv((uint64_t)WE_WIN16 * WE_ENT16, 128), 128, 0, ctx->d_com_tab, ctx->d_com_tab16);
    if (!ctx->d_com1_tab16) KB_CUDA(cudaMalloc((void**)&ctx->d_com1_tab16, GT_TAB16_LIMBS * 4));
    KB_LAUNCH(ctx, gt_table16_kernel, cdiv((uint64_t)WE_WIN16 * WE_ENT16, 128), 128, 0, ctx->d_com1_tab, ctx->d_com1_tab16);
    ctx->com_tab16_valid = true;
  }
  ctx->com_msgs += n;
  const bool wide = ctx->com_tab16_valid;
  if (n <= ctx->wp_enc_max_n) {   // small batches: one warp per message (pairing_warp.cu), a thread needs 2.6 ms however few there are
    timer_start(ctx, KB_T_ENCRYPT);
    wp_encrypt_small_launch(ctx, wide ? ctx->d_com_tab16 : ctx->d_com_tab, wide ? ctx->d_com1_tab16 : ctx->d_com1_tab, wide ? 1 : 0, ctx->d_gt_tab16,
                            ctx->d_tau2_tab16, ctx->d_g2_tab16, d_points, d_values, d_r, d_msgs, d_off, n, d_ct, d_ct_inf, d_msg_ct);
    timer_stop(ctx, KB_T_ENCRYPT);
    return;
  }
  timer_start(ctx, KB_T_ENCRYPT);
  if (ctx->enc_gt_st)
    st_encrypt_gt_launch(ctx, wide ? ctx->d_com_tab16 : ctx->d_com_tab, wide ? ctx->d_com1_tab16 : ctx->d_com1_tab, ctx->d_gt_tab16, d_values, d_r, d_msgs,
                         d_off, n, wide ? 1 : 0, d_msg_ct);
  else
    KB_LAUNCH(ctx, encrypt_kernel, cdiv(n, 128), 128, 0, wide ? ctx->d_com_tab16 : ctx->d_com_tab, wide ? ctx->d_com1_tab16 : ctx->d_com1_tab,
              ctx->d_gt_tab16, d_values, d_r, d_msgs, d_off, n, wide ? 1 : 0, d_msg_ct);
  KB_LAUNCH(ctx, encrypt_ct_kernel, cdiv(n, 128), 128, 0, ctx->d_tau2_tab16, ctx->d_g2_tab16, d_points, d_r, n, d_ct, d_ct_inf);
  timer_stop(ctx, KB_T_ENCRYPT);
}

// ------------------------------------------------------------------------------------------
// verify (src/kzg.rs:127-151): e(C - v G1, G2) == e(pi, tau_2 - a G2), as the reference computes it - two pairings per
// item and a comparison.  The operands are prepared per item (generic G1 multiplication of the generator, fixed-base G2
// multiplication from the 16-bit-window table), the 2n pairings run on the pairing VM, the GT images are compared.
// A pair with a point at infinity gives GT = 1 exactly as in arkworks (the VM kernel's `trivial` path).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) verify_prep_kernel(const uint32_t* __restrict__ com, const uint8_t* __restrict__ com_inf,
                                                          const uint32_t* __restrict__ points, const uint32_t* __restrict__ values,
                                                          const uint32_t* __restrict__ proofs, const uint8_t* __restrict__ pinf,
                                                          const uint32_t* __restrict__ tau2_xy, const uint32_t* __restrict__ g2_tab16,
                                                          const uint32_t* __restrict__ g2_gen, const uint32_t* __restrict__ g1_gen, uint64_t n,
                                                          uint32_t* __restrict__ g1 /* 2n */, uint8_t* __restrict__ g1_inf,
                                                          uint32_t* __restrict__ g2 /* 2n */, uint8_t* __restrict__ g2_inf) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr kv = fp_from_mont<FrParams>(fp_load<FrParams>(values + 8 * i));
  Fr ka = fp_from_mont<FrParams>(fp_load<FrParams>(points + 8 * i));
  // commitment - [value]_1
  G1 vg = g1_mul_glv(to_xyzz(ld_g1(g1_gen)), kv.v);
  G1 lhs = ec_add(to_xyzz(load_g1_flag(com, com_inf, i)), neg(vg));
  st_g1(g1 + 16 * i, to_affine(lhs));
  g1_inf[i] = lhs.is_inf() ? 1 : 0;
  st_g2(g2 + 32 * i, ld_g2(g2_gen));
  g2_inf[i] = 0;
  // proof, [tau]_2 - [point]_2
  G1Affine pi = load_g1_flag(proofs, pinf, i);
  st_g1(g1 + 16 * (n + i), pi);
  g1_inf[n + i] = pi.is_inf() ? 1 : 0;
  G2 acc = to_xyzz(ld_g2(tau2_xy));
  for (int w = 0; w < WE_WIN16; w++) {
    uint32_t d = half_of(ka.v, w);
    if (d) { G2Affine t = ld_g2(g2_tab16 + 32 * ((size_t)w * WE_ENT16 + d - 1)); t.y = -t.y; acc = ec_add_mixed(acc, t); }
  }
  st_g2(g2 + 32 * (n + i), to_affine(acc));
  g2_inf[n + i] = acc.is_inf() ? 1 : 0;
}

__global__ void __launch_bounds__(256) gt_pairs_equal_kernel(const uint32_t* __restrict__ gt /* 2n x 96 words */, uint64_t n, uint8_t* __restrict__ ok) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* a = reinterpret_cast<const uint4*>(gt + 96 * i);
  const uint4* b = reinterpret_cast<const uint4*>(gt + 96 * (n + i));
  uint32_t diff = 0;
#pragma unroll 4
  for (int k = 0; k < 24; k++) { uint4 x = a[k], y = b[k]; diff |= (x.x ^ y.x) | (x.y ^ y.y) | (x.z ^ y.z) | (x.w ^ y.w); }
  ok[i] = diff == 0 ? 1 : 0;
}

void verify_batch(kb_ctx* ctx, const uint32_t* d_com, const uint8_t* d_com_inf, const uint32_t* d_points, const uint32_t* d_values,
                  const uint32_t* d_proofs, const uint8_t* d_pinf, uint64_t n, uint8_t* d_ok) {
  if (!n) return;
  DevBuf<uint32_t> g2gen(ctx, 32), g1gen(ctx, 16);
  KB_CUDA(cudaMemcpyAsync(g2gen, consts::G2_GEN, 128, cudaMemcpyHostToDevice, ctx->stream));
  KB_CUDA(cudaMemcpyAsync(g1gen, consts::G1_GEN, 64, cudaMemcpyHostToDevice, ctx->stream));
  DevBuf<uint32_t> g1(ctx, 16 * 2 * n), g2(ctx, 32 * 2 * n), gt(ctx, 96 * 2 * n);
  DevBuf<uint8_t> g1i(ctx, 2 * n), g2i(ctx, 2 * n);
  // tau_2 itself is entry (w = 0, d = 1) of its fixed-base table
  KB_LAUNCH(ctx, verify_prep_kernel, cdiv(n, 128), 128, 0, d_com, d_com_inf, d_points, d_values, d_proofs, d_pinf,
            ctx->d_tau2_tab, ctx->d_g2_tab16, g2gen.p, g1gen.p, n, g1.p, g1i.p, g2.p, g2i.p);
  pairing_batch(ctx, g1, g1i, g2, g2i, 2 * n, reinterpret_cast<uint8_t*>(gt.p));
  KB_LAUNCH(ctx, gt_pairs_equal_kernel, cdiv(n, 256), 256, 0, gt.p, n, d_ok);
}

}  // namespace kb
