// Library-internal context and launch helpers (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <stdexcept>
#include "../../include/keaki_b200.h"
#include "pairing.cuh"
#include "glv.cuh"

namespace kb {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ApiError : std::runtime_error {
  int32_t code;
  ApiError(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define KB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
  throw kb::CudaError(std::string(#x) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); } while (0)

// Fixed-base tables for the MSM over the SRS: tab[w * n + i] = 2^(c w) * g1[i], affine.
struct MsmTable { int c = 0, nwin = 0; uint64_t n = 0; uint32_t* d = nullptr; };

enum { KB_T_TOTAL = 0, KB_T_MSM_ACC = 1, KB_T_PAIRING = 2, KB_T_ENCRYPT = 3, KB_T_SETUP = 4, KB_T_COUNT = 5 };

// One host thread per peer device of a multi-device context, alive for the life of the context: a sharded call posts one
// job per device and waits for all of them (no thread creation on the hot path).
class DeviceWorker {
 public:
  DeviceWorker() : th_([this] { loop(); }) {}
  ~DeviceWorker() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; }
    cv_.notify_all();
    th_.join();
  }
  void post(std::function<void()> job) {
    { std::lock_guard<std::mutex> l(m_); job_ = std::move(job); busy_ = true; }
    cv_.notify_all();
  }
  void wait() { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [this] { return !busy_; }); }

 private:
  void loop() {
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return stop_ || (busy_ && job_); });
        if (stop_) return;
        job = std::move(job_);
        job_ = nullptr;
      }
      job();
      { std::lock_guard<std::mutex> l(m_); busy_ = false; }
      cv_.notify_all();
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::function<void()> job_;
  bool busy_ = false, stop_ = false;
  std::thread th_;
};

}  // namespace kb

struct kb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaMemPool_t pool = nullptr;         // private stream-ordered pool for the scratch of every call (the device's default pool is left alone)
  cudaStream_t copy_stream = nullptr;   // host-to-device copies that overlap the main stream's kernels (msm_g1_host)
  cudaEvent_t ev_copy[4] = {};
  std::string err;
  uint64_t launches = 0;
  int sm_count = 148;
  // multi-device context (kb_ctx_create_multi): this is the first device, `peers` are the contexts of the others
  std::vector<kb_ctx*> peers;
  std::vector<kb::DeviceWorker*> workers;   // one per peer
  bool peer_access = true;

  // SRS (affine, Montgomery), resident for the life of the context
  uint32_t* d_srs = nullptr;
  uint64_t srs_n = 0;
  std::vector<kb::MsmTable> msm_tabs;

  // fixed-base tables for witness encryption (8-bit windows, 32 windows x 255 entries)
  uint32_t* d_g2_tab = nullptr;      // d * 2^(8w) * G2, affine (32 limbs each)
  uint32_t* d_tau2_tab = nullptr;    // d * 2^(8w) * tau_2
  uint32_t* d_gt_tab = nullptr;      // e(G1, G2)^(d 2^(8w)), Fq12 Montgomery (96 limbs each)
  uint32_t* d_com_tab = nullptr;     // e(com, G2)^(d 2^(8w)) for the cached commitment
  uint32_t* d_g2_tab16 = nullptr;    // 16-bit-window versions (16 windows x 65535 entries) of the SRS-constant tables
  uint32_t* d_tau2_tab16 = nullptr;
  uint32_t* d_gt_tab16 = nullptr;
  uint32_t com_cached[17] = {0};     // xy + inf flag of the commitment the table was built for
  bool com_tab_valid = false;
  uint32_t* d_com_tab16 = nullptr;   // 16-bit-window table of the cached commitment, built once it has served 2^15 messages
  bool com_tab16_valid = false;
  uint32_t* d_com1_tab = nullptr;    // the same two tables for A' = e(com - G1, G2) = A / gT: messages with value = 1
  uint32_t* d_com1_tab16 = nullptr;  //   (laconic OT: every value is a bit) need A'^r only
  uint64_t com_msgs = 0;             // messages encrypted under the cached commitment so far

  // pairing VM (pairing_vm.cu): program + constants resident on the device
  uint64_t* d_vm_prog = nullptr;
  uint32_t* d_vm_consts = nullptr;
  int vm_slots = 0, vm_gslots = 0, vm_block = 128, vm_minb = 2;
  uint64_t vm_out = 0;               // the six output slots, one byte each
  // compiled single-thread pairing (pairing_st.cu): Frobenius / twist constants; which kernel the batched paths use
  uint32_t* d_st_consts = nullptr;
  int st_shape = 0;
  bool enc_gt_st = true;             // GT side of encrypt on the shared-memory Fq12 machinery of pairing_st (KB_ENCRYPT_GT=tower: we.cu's generic code)
  int st_segments = 12;              // batches of more than one round: segments per pairing (pairing_st.cu, KB_PAIRING_SEGMENTS)
  int pairing_impl = 1;              // 0 = pairing VM (two lanes + interpreter), 1 = compiled single-thread kernel; KB_PAIRING_IMPL=vm|st
  // warp-cooperative pairing (pairing_warp.cu): dense step descriptors of the two schedules (pairing, GT window bases),
  // their output slots, the Fq2 constants; batches of at most wp_max_n pairings use it (KB_PAIRING_WARP_MAX, 0 = never)
  uint32_t* d_wp_words[3] = {nullptr, nullptr, nullptr};   // schedules: pairing, GT window bases, GT product of 48
  uint16_t* d_wp_outs[3] = {nullptr, nullptr, nullptr};
  uint16_t* d_wp_ins = nullptr;      // input slots of the GT product schedule
  uint32_t* d_wp_consts = nullptr;
  uint64_t wp_max_n = 0;
  uint64_t wp_enc_max_n = 0;         // kb_encrypt_batch: batches of at most this many messages take one warp per message (KB_ENCRYPT_WARP_MAX)
  bool ntt_radix2 = false;           // KB_NTT_RADIX2=1: one radix-2 stage per launch even for small G1 transforms (measurement)

  // FK open-all cache: hat_s = DFT_2d(reversed SRS prefix) per d
  struct FkCache { uint64_t d = 0; uint32_t* d_hat_s = nullptr; };
  std::vector<FkCache> fk_cache;        // at most KB_FK_CACHE_MAX entries, oldest evicted (256 MiB per entry at d = 2^20)

  cudaEvent_t ev[2 * kb::KB_T_COUNT] = {};
  float last_ms[kb::KB_T_COUNT] = {-1.f, -1.f, -1.f, -1.f, -1.f};
};

namespace kb {

// Stream-ordered scratch buffer (pool allocator: no cudaMalloc on the hot path after warm-up).
template <class T>
struct DevBuf {
  T* p = nullptr;
  cudaStream_t s;
  DevBuf(kb_ctx* ctx, size_t count) : s(ctx->stream) {
    if (count) KB_CUDA(cudaMallocFromPoolAsync((void**)&p, count * sizeof(T), ctx->pool, s));
  }
  ~DevBuf() { if (p) cudaFreeAsync(p, s); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  operator T*() const { return p; }
};

inline bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Input staged on the device: either the caller's device pointer, or an async copy of host memory.
template <class T>
struct DevIn {
  const T* p = nullptr;
  T* owned = nullptr;
  cudaStream_t s;
  DevIn(kb_ctx* ctx, const T* src, size_t count) : s(ctx->stream) {
    if (!src || !count) return;
    if (is_device_ptr(src)) { p = src; return; }
    KB_CUDA(cudaMallocFromPoolAsync((void**)&owned, count * sizeof(T), ctx->pool, s));
    KB_CUDA(cudaMemcpyAsync(owned, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    p = owned;
  }
  ~DevIn() { if (owned) cudaFreeAsync(owned, s); }
  DevIn(const DevIn&) = delete;
  DevIn& operator=(const DevIn&) = delete;
  operator const T*() const { return p; }
};

// Output: device scratch copied back to host at the end, or the caller's device pointer directly.
template <class T>
struct DevOut {
  T* p = nullptr;
  T* owned = nullptr;
  T* host = nullptr;
  size_t count;
  cudaStream_t s;
  DevOut(kb_ctx* ctx, T* dst, size_t count_) : count(count_), s(ctx->stream) {
    if (!dst || !count) return;
    if (is_device_ptr(dst)) { p = dst; return; }
    KB_CUDA(cudaMallocFromPoolAsync((void**)&owned, count * sizeof(T), ctx->pool, s));
    p = owned; host = dst;
  }
  void finish() { if (owned && host) KB_CUDA(cudaMemcpyAsync(host, owned, count * sizeof(T), cudaMemcpyDeviceToHost, s)); }
  ~DevOut() { if (owned) cudaFreeAsync(owned, s); }
  DevOut(const DevOut&) = delete;
  DevOut& operator=(const DevOut&) = delete;
  operator T*() const { return p; }
};

#define KB_LAUNCH(ctx, kernel, grid, block, smem, ...) do { \
  kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
  (ctx)->launches++; KB_CUDA(cudaGetLastError()); } while (0)

inline void timer_start(kb_ctx* c, int which) { KB_CUDA(cudaEventRecord(c->ev[2 * which], c->stream)); }
inline void timer_stop(kb_ctx* c, int which) { KB_CUDA(cudaEventRecord(c->ev[2 * which + 1], c->stream)); c->last_ms[which] = -2.f; }
inline void timers_collect(kb_ctx* c) {
  for (int i = 0; i < KB_T_COUNT; i++)
    if (c->last_ms[i] == -2.f) { float ms = -1.f; if (cudaEventElapsedTime(&ms, c->ev[2 * i], c->ev[2 * i + 1]) != cudaSuccess) { cudaGetLastError(); ms = -1.f; } c->last_ms[i] = ms; }
}

// ---- fixed-base window tables of the witness encryption (we.cu builds them, we.cu and pairing_warp.cu read them)
static constexpr int WE_WIN = 32;          // 8-bit windows over 256 bits
static constexpr int WE_ENT = 255;         // non-zero digits
static constexpr size_t G2_TAB_LIMBS = (size_t)WE_WIN * WE_ENT * 32;
static constexpr size_t GT_TAB_LIMBS = (size_t)WE_WIN * WE_ENT * 96;
// Bases that are fixed for the life of an SRS (G2, tau_2, gT = e(G1, G2)) get 16-bit windows: 16 x 65535 entries
// (128 MiB per G2 table, 384 MiB for gT) halve the group operations per message; HBM capacity is what B200 has to
// spare.  A = e(com, G2) changes per commitment: it starts with 8-bit windows (32 + 8160 Fq12 products to build) and
// is upgraded to 16-bit windows (one more Fq12 product per entry, ~1 M) once the commitment has served 2^16
// messages - the point where the 16 products saved per message would have paid for the build (laconic OT encrypts
// 2 n messages under one commitment, tests/laconic_ot.rs:89-109).
static constexpr int WE_WIN16 = 16;
static constexpr int WE_ENT16 = 65535;
static constexpr size_t G2_TAB16_LIMBS = (size_t)WE_WIN16 * WE_ENT16 * 32;
static constexpr size_t GT_TAB16_LIMBS = (size_t)WE_WIN16 * WE_ENT16 * 96;

__device__ __forceinline__ uint32_t byte_of(const uint32_t* k, int w) { return (k[w >> 2] >> ((w & 3) * 8)) & 255u; }
__device__ __forceinline__ uint32_t half_of(const uint32_t* k, int w) { return (k[w >> 1] >> ((w & 1) * 16)) & 65535u; }

inline unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// device-side loaders shared by the kernels
__device__ __forceinline__ Fq2 ld_fq2(const uint32_t* p) { Fq2 r; r.c0 = fp_load<FqParams>(p); r.c1 = fp_load<FqParams>(p + 8); return r; }
__device__ __forceinline__ void st_fq2(uint32_t* p, const Fq2& a) { fp_store<FqParams>(p, a.c0); fp_store<FqParams>(p + 8, a.c1); }
__device__ __forceinline__ G1Affine ld_g1(const uint32_t* p) { G1Affine r; r.x = fp_load<FqParams>(p); r.y = fp_load<FqParams>(p + 8); return r; }
__device__ __forceinline__ void st_g1(uint32_t* p, const G1Affine& a) { fp_store<FqParams>(p, a.x); fp_store<FqParams>(p + 8, a.y); }
__device__ __forceinline__ G2Affine ld_g2(const uint32_t* p) { G2Affine r; r.x = ld_fq2(p); r.y = ld_fq2(p + 16); return r; }
__device__ __forceinline__ void st_g2(uint32_t* p, const G2Affine& a) { st_fq2(p, a.x); st_fq2(p + 16, a.y); }
__device__ __forceinline__ G1 ld_g1x(const uint32_t* p) { G1 r; r.x = fp_load<FqParams>(p); r.y = fp_load<FqParams>(p + 8); r.zz = fp_load<FqParams>(p + 16); r.zzz = fp_load<FqParams>(p + 24); return r; }
__device__ __forceinline__ void st_g1x(uint32_t* p, const G1& a) { fp_store<FqParams>(p, a.x); fp_store<FqParams>(p + 8, a.y); fp_store<FqParams>(p + 16, a.zz); fp_store<FqParams>(p + 24, a.zzz); }
__device__ __forceinline__ Fq12 ld_fq12(const uint32_t* p) {
  Fq12 r;
  r.c0.c0 = ld_fq2(p); r.c0.c1 = ld_fq2(p + 16); r.c0.c2 = ld_fq2(p + 32);
  r.c1.c0 = ld_fq2(p + 48); r.c1.c1 = ld_fq2(p + 64); r.c1.c2 = ld_fq2(p + 80);
  return r;
}
__device__ __forceinline__ void st_fq12(uint32_t* p, const Fq12& a) {
  st_fq2(p, a.c0.c0); st_fq2(p + 16, a.c0.c1); st_fq2(p + 32, a.c0.c2);
  st_fq2(p + 48, a.c1.c0); st_fq2(p + 64, a.c1.c1); st_fq2(p + 80, a.c1.c2);
}

// ---- entry points implemented per translation unit (msm.cu, we.cu, poly.cu) ----
void msm_g1(kb_ctx* ctx, const uint32_t* d_scalars, uint64_t first, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf);
void msm_g1_host(kb_ctx* ctx, const uint32_t* h_scalars, uint64_t first, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf);
void g1_sum(kb_ctx* ctx, const uint32_t* d_pts, const uint8_t* d_inf, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf);
void g1_xyzz_sum_to_affine(kb_ctx* ctx, uint32_t* d_xyzz /* n*32, destroyed */, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf);
void msm_free_tables(kb_ctx* ctx);
void g1_mul_gen_batch(kb_ctx* ctx, const uint32_t* d_scalars, uint64_t n, uint32_t* d_out_xy, uint8_t* d_out_inf);
void launch_msm_accumulate(kb_ctx* ctx, const uint32_t* tab, uint64_t tab_n, uint64_t first, const uint32_t* offsets,
                           const uint32_t* entries, const uint32_t* perm, uint32_t nb, uint32_t* buckets, bool into, uint64_t max_entries);
void launch_msm_reduce(kb_ctx* ctx, const uint32_t* buckets, uint32_t nb, uint32_t* d_out_xy, uint8_t* d_out_inf);
void srs_generate(kb_ctx* ctx, const uint32_t* d_tau, uint64_t first_power, uint64_t n, uint32_t* d_tau_g2_out);
void we_init_tables(kb_ctx* ctx);                       // G2 generator + gT tables (ctx creation)
void we_set_tau2(kb_ctx* ctx, const uint32_t* d_tau2);  // tau_2 table (SRS upload)
void we_free(kb_ctx* ctx);
void we_upload_consts();
void vm_init(kb_ctx* ctx);                             // pairing VM program upload (ctx creation)
void vm_free(kb_ctx* ctx);
void st_init(kb_ctx* ctx);                             // compiled pairing: constants upload (ctx creation)
void st_free(kb_ctx* ctx);
void st_pairing_launch(kb_ctx* ctx, const uint32_t* d_g1, const uint8_t* d_g1_inf, const uint32_t* d_g2, const uint8_t* d_g2_inf, uint64_t n,
                       int mode, uint32_t* d_gt, const uint8_t* d_msg_ct, const uint64_t* d_off, uint8_t* d_out);
void wp_init(kb_ctx* ctx);                             // warp-cooperative pairing: schedules upload (ctx creation)
void wp_free(kb_ctx* ctx);
void wp_pairing_launch(kb_ctx* ctx, const uint32_t* d_g1, const uint8_t* d_g1_inf, const uint32_t* d_g2, const uint8_t* d_g2_inf, uint64_t n,
                       int mode, uint32_t* d_gt, const uint8_t* d_msg_ct, const uint64_t* d_off, uint8_t* d_out);
void wp_gt_bases_launch(kb_ctx* ctx, const uint32_t* d_a, uint32_t* d_bases);   // bases[w] = a^(2^(8w)), w < 32
void wp_encrypt_small_launch(kb_ctx* ctx, const uint32_t* com_tab_a, const uint32_t* com_tab_a1, int com_wide, const uint32_t* gt_tab16,
                             const uint32_t* tau2_tab16, const uint32_t* g2_tab16, const uint32_t* d_points, const uint32_t* d_values,
                             const uint32_t* d_r, const uint8_t* d_msgs, const uint64_t* d_off, uint64_t n, uint32_t* d_ct, uint8_t* d_ct_inf,
                             uint8_t* d_msg_ct);
void st_encrypt_gt_launch(kb_ctx* ctx, const uint32_t* com_tab_a, const uint32_t* com_tab_a1, const uint32_t* gt_tab16, const uint32_t* d_values,
                          const uint32_t* d_r, const uint8_t* d_msgs, const uint64_t* d_off, uint64_t n, int com_wide, uint8_t* d_msg_ct);
void pairing_batch(kb_ctx* ctx, const uint32_t* d_g1, const uint8_t* d_g1_inf, const uint32_t* d_g2, const uint8_t* d_g2_inf,
                   uint64_t n, uint8_t* d_gt_bytes);
void decrypt_batch(kb_ctx* ctx, const uint32_t* d_proofs, const uint8_t* d_pinf, const uint32_t* d_ct, const uint8_t* d_cinf,
                   const uint8_t* d_msg_ct, const uint64_t* d_off, uint64_t n, uint8_t* d_out);
void encrypt_batch(kb_ctx* ctx, const uint32_t* h_com_xy, uint8_t com_inf, const uint32_t* d_points, const uint32_t* d_values,
                   const uint32_t* d_r, const uint8_t* d_msgs, const uint64_t* d_off, uint64_t n,
                   uint32_t* d_ct, uint8_t* d_ct_inf, uint8_t* d_msg_ct);
void verify_batch(kb_ctx* ctx, const uint32_t* d_com, const uint8_t* d_com_inf, const uint32_t* d_points, const uint32_t* d_values,
                  const uint32_t* d_proofs, const uint8_t* d_pinf, uint64_t n, uint8_t* d_ok);
void fr_ntt(kb_ctx* ctx, uint32_t* d_data, uint64_t n, bool inverse);
void open_batch(kb_ctx* ctx, const uint32_t* d_coeffs, uint64_t d, const uint32_t* d_points, uint64_t m,
                uint32_t* d_proofs, uint8_t* d_inf);
void open_all_fk(kb_ctx* ctx, const uint32_t* d_coeffs, uint64_t d, uint32_t* d_proofs, uint8_t* d_inf);
void fk_free(kb_ctx* ctx);
void wire_upload_consts();
void srs_validate(kb_ctx* ctx, unsigned long long* d_first_bad);
void g1_serialize(kb_ctx* ctx, const uint32_t* d_xy, const uint8_t* d_inf, uint64_t n, int compress, uint8_t* d_out);
void g2_serialize(kb_ctx* ctx, const uint32_t* d_xy, const uint8_t* d_inf, uint64_t n, int compress, uint8_t* d_out);
void g1_deserialize(kb_ctx* ctx, const uint8_t* d_in, uint64_t n, int compress, int validate, uint32_t* d_xy, uint8_t* d_inf, uint8_t* d_ok);
void g2_deserialize(kb_ctx* ctx, const uint8_t* d_in, uint64_t n, int compress, int validate, uint32_t* d_xy, uint8_t* d_inf, uint8_t* d_ok);
void debug_fp_op(kb_ctx* ctx, int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, uint64_t n);

}  // namespace kb
