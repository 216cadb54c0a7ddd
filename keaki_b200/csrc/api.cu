// C ABI of libkeaki_b200.so (declared in include/keaki_b200.h).  Thin: argument checks, staging of
// host buffers, stream synchronisation and error translation.  There is no host arithmetic and no
// CPU fallback anywhere behind these entry points.
#include "ctx.cuh"
#include <cstdlib>
#include <thread>

using namespace kb;

namespace {

int32_t fail(kb_ctx* ctx, int32_t code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define KB_API_BEGIN(ctx)                          \
  if (!(ctx)) return KB_ERR_ARG;                   \
  try {                                            \
    KB_CUDA(cudaSetDevice((ctx)->device));         \
    (ctx)->last_ms[KB_T_SETUP] = -1.f;             \
    timer_start((ctx), KB_T_TOTAL);

#define KB_API_END(ctx)                                                   \
    timer_stop((ctx), KB_T_TOTAL);                                        \
    KB_CUDA(cudaStreamSynchronize((ctx)->stream));                        \
    timers_collect(ctx);                                                  \
    return KB_OK;                                                         \
  } catch (const ApiError& e) { cudaStreamSynchronize((ctx)->stream); return fail((ctx), e.code, e.what());       \
  } catch (const CudaError& e) { cudaGetLastError(); return fail((ctx), KB_ERR_CUDA, e.what());                  \
  } catch (const std::exception& e) { return fail((ctx), KB_ERR_CUDA, e.what()); }

void need(bool ok, const char* what) { if (!ok) throw ApiError(KB_ERR_ARG, what); }
void need_srs(kb_ctx* ctx) { if (!ctx->d_srs) throw ApiError(KB_ERR_NO_SRS, "no SRS: call kb_srs_upload or kb_srs_generate first"); }

}  // namespace

extern "C" {

const char* kb_version(void) { return "keaki_b200 0.1.0 (sm_100a)"; }

int32_t kb_ctx_create(int32_t device, kb_ctx** out) {
  if (!out) return KB_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return KB_ERR_CUDA;  // no usable CUDA device: there is deliberately no CPU path
  }
  kb_ctx* ctx = new kb_ctx();
  try {
    ctx->device = device;
    KB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    KB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) throw CudaError(std::string("device ") + prop.name + " is not sm_100-class; this library is built for sm_100a only");
    ctx->sm_count = prop.multiProcessorCount;
    KB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {
      int prio_lo = 0, prio_hi = 0;   // the side stream (copies, second-pass sort) gets the higher priority
      KB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      int prio = prio_hi;
      if (const char* e = getenv("KB_SIDE_PRIO")) prio = atoi(e) > 0 ? prio_lo : atoi(e) < 0 ? prio_hi : 0;   // tuning override
      KB_CUDA(cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, prio));
    }
    for (auto& e : ctx->ev_copy) KB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    {   // a pool of the context's own (kept warm between calls; the device's default pool and a co-resident framework's are not touched)
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = device;
      KB_CUDA(cudaMemPoolCreate(&ctx->pool, &props));
      uint64_t thresh = UINT64_MAX;
      KB_CUDA(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    }
    for (auto& e : ctx->ev) KB_CUDA(cudaEventCreate(&e));
    we_upload_consts();
    wire_upload_consts();
    wp_init(ctx);
    st_init(ctx);   // before the WE tables: gT = e(G1, G2) is computed by the compiled pairing kernel
    we_init_tables(ctx);
    vm_init(ctx);
    if (const char* e = getenv("KB_PAIRING_IMPL")) ctx->pairing_impl = (e[0] == 'v') ? 0 : 1;   // tuning override
    KB_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (const std::exception& e) {
    fprintf(stderr, "kb_ctx_create: %s\n", e.what());
    cudaGetLastError();
    kb_ctx_destroy(ctx);   // frees whatever was created so far (every member starts null)
    cudaGetLastError();
    return KB_ERR_CUDA;
  }
  *out = ctx;
  return KB_OK;
}

void kb_ctx_destroy(kb_ctx* ctx) {
  if (!ctx) return;
  for (DeviceWorker* w : ctx->workers) delete w;
  ctx->workers.clear();
  for (kb_ctx* p : ctx->peers) kb_ctx_destroy(p);
  ctx->peers.clear();
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  msm_free_tables(ctx);
  fk_free(ctx);
  we_free(ctx);
  vm_free(ctx);
  st_free(ctx);
  wp_free(ctx);
  if (ctx->d_srs) cudaFree(ctx->d_srs);
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_copy) if (e) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  delete ctx;
}

const char* kb_last_error(const kb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t kb_launch_count(const kb_ctx* ctx) {
  if (!ctx) return 0;
  uint64_t n = ctx->launches;
  for (const kb_ctx* p : ctx->peers) n += p->launches;
  return n;
}
int32_t kb_ctx_device_count(const kb_ctx* ctx) { return ctx ? 1 + (int32_t)ctx->peers.size() : 0; }
float kb_last_kernel_ms(const kb_ctx* ctx, int32_t which) { return (ctx && which >= 0 && which < KB_T_COUNT) ? ctx->last_ms[which] : -1.f; }
uint64_t kb_srs_len(const kb_ctx* ctx) { return ctx ? ctx->srs_n : 0; }

static int32_t single_srs_upload(kb_ctx* ctx, const uint32_t* g1_aff_xy, uint64_t n, const uint32_t* tau_g2_xy) {
  KB_API_BEGIN(ctx)
  need(tau_g2_xy != nullptr && (n == 0 || g1_aff_xy != nullptr), "kb_srs_upload: null pointer");
  if (ctx->d_srs) { KB_CUDA(cudaFree(ctx->d_srs)); ctx->d_srs = nullptr; }
  msm_free_tables(ctx);
  fk_free(ctx);
  KB_CUDA(cudaMalloc((void**)&ctx->d_srs, (size_t)(n ? n : 1) * 64));
  ctx->srs_n = n;
  if (n) KB_CUDA(cudaMemcpyAsync(ctx->d_srs, g1_aff_xy, n * 64, cudaMemcpyDefault, ctx->stream));
  DevIn<uint32_t> tau2(ctx, tau_g2_xy, 32);
  we_set_tau2(ctx, tau2);
  KB_API_END(ctx)
}

static int32_t single_srs_generate(kb_ctx* ctx, const uint32_t* tau, uint64_t first_power, uint64_t n, uint32_t* out_g1_xy, uint32_t* out_tau_g2_xy) {
  KB_API_BEGIN(ctx)
  need(tau != nullptr, "kb_srs_generate: null tau");
  fk_free(ctx);
  DevIn<uint32_t> dtau(ctx, tau, 8);
  DevBuf<uint32_t> tau2(ctx, 32);
  srs_generate(ctx, dtau, first_power, n, tau2);
  we_set_tau2(ctx, tau2);
  if (out_g1_xy && n) KB_CUDA(cudaMemcpyAsync(out_g1_xy, ctx->d_srs, n * 64, cudaMemcpyDefault, ctx->stream));
  if (out_tau_g2_xy) KB_CUDA(cudaMemcpyAsync(out_tau_g2_xy, tau2.p, 128, cudaMemcpyDefault, ctx->stream));
  KB_API_END(ctx)
}

int32_t kb_srs_validate(kb_ctx* ctx, uint64_t* first_bad) {
  unsigned long long bad = ~0ull;
  KB_API_BEGIN(ctx)
  need_srs(ctx);
  DevBuf<unsigned long long> d_bad(ctx, 1);
  srs_validate(ctx, d_bad);
  KB_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(bad), cudaMemcpyDeviceToHost, ctx->stream));
  KB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (first_bad) *first_bad = bad;
  if (bad != ~0ull)
    throw ApiError(KB_ERR_INVALID_POINT, bad < ctx->srs_n ? "SRS: G1 power " + std::to_string(bad) + " is not a point of the curve"
                                                          : std::string("SRS: [tau]_2 is not a point of the r-torsion of the twist"));
  KB_API_END(ctx)
}

static int32_t single_msm_g1(kb_ctx* ctx, const uint32_t* scalars, uint64_t first, uint64_t n, uint32_t out_xy[16], uint8_t* out_inf) {
  KB_API_BEGIN(ctx)
  need(out_xy != nullptr && (n == 0 || scalars != nullptr), "kb_msm_g1: null pointer");
  need_srs(ctx);
  if (first + n > ctx->srs_n)
    throw ApiError(KB_ERR_POLY_TOO_LARGE, "PolynomialTooLarge(" + std::to_string(first + n) + ", " + std::to_string(ctx->srs_n) + ")");
  DevOut<uint32_t> o(ctx, out_xy, 16);
  DevBuf<uint8_t> inf_scratch(ctx, 1);
  DevOut<uint8_t> oi(ctx, out_inf, 1);
  if (n && !is_device_ptr(scalars)) msm_g1_host(ctx, scalars, first, n, o, out_inf ? oi.p : inf_scratch.p);   // copy overlapped with compute
  else msm_g1(ctx, scalars, first, n, o, out_inf ? oi.p : inf_scratch.p);
  o.finish(); oi.finish();
  KB_API_END(ctx)
}

int32_t kb_g1_mul_gen_batch(kb_ctx* ctx, const uint32_t* scalars, uint64_t n, uint32_t* out_xy, uint8_t* out_inf) {
  KB_API_BEGIN(ctx)
  need(n == 0 || (scalars && out_xy && out_inf), "kb_g1_mul_gen_batch: null pointer");
  DevIn<uint32_t> s(ctx, scalars, n * 8);
  DevOut<uint32_t> o(ctx, out_xy, n * 16);
  DevOut<uint8_t> oi(ctx, out_inf, n);
  g1_mul_gen_batch(ctx, s, n, o, oi);
  o.finish(); oi.finish();
  KB_API_END(ctx)
}

int32_t kb_g1_sum(kb_ctx* ctx, const uint32_t* pts_xy, const uint8_t* inf, uint64_t n, uint32_t out_xy[16], uint8_t* out_inf) {
  KB_API_BEGIN(ctx)
  need(out_xy != nullptr && (n == 0 || pts_xy != nullptr), "kb_g1_sum: null pointer");
  DevIn<uint32_t> p(ctx, pts_xy, n * 16);
  DevIn<uint8_t> pi(ctx, inf, n);
  DevOut<uint32_t> o(ctx, out_xy, 16);
  DevBuf<uint8_t> inf_scratch(ctx, 1);
  DevOut<uint8_t> oi(ctx, out_inf, 1);
  g1_sum(ctx, p, pi, n, o, out_inf ? oi.p : inf_scratch.p);
  o.finish(); oi.finish();
  KB_API_END(ctx)
}

int32_t kb_open_batch(kb_ctx* ctx, const uint32_t* coeffs, uint64_t d, const uint32_t* points, uint64_t m,
                      uint32_t* proofs_xy, uint8_t* proofs_inf) {
  KB_API_BEGIN(ctx)
  need((d == 0 || coeffs) && (m == 0 || (points && proofs_xy && proofs_inf)), "kb_open_batch: null pointer");
  need_srs(ctx);
  DevIn<uint32_t> c(ctx, coeffs, d * 8), z(ctx, points, m * 8);
  DevOut<uint32_t> o(ctx, proofs_xy, m * 16);
  DevOut<uint8_t> oi(ctx, proofs_inf, m);
  open_batch(ctx, c, d, z, m, o, oi);
  o.finish(); oi.finish();
  KB_API_END(ctx)
}

int32_t kb_open_all_fk(kb_ctx* ctx, const uint32_t* coeffs, uint64_t d, uint32_t* proofs_xy, uint8_t* proofs_inf) {
  KB_API_BEGIN(ctx)
  need(coeffs && proofs_xy && proofs_inf && d > 0, "kb_open_all_fk: null pointer or d = 0");
  need_srs(ctx);
  DevIn<uint32_t> c(ctx, coeffs, d * 8);
  DevOut<uint32_t> o(ctx, proofs_xy, d * 16);
  DevOut<uint8_t> oi(ctx, proofs_inf, d);
  open_all_fk(ctx, c, d, o, oi);
  o.finish(); oi.finish();
  KB_API_END(ctx)
}

int32_t kb_fr_ntt(kb_ctx* ctx, uint32_t* data, uint64_t n, int32_t inverse) {
  KB_API_BEGIN(ctx)
  need(data != nullptr && n > 0, "kb_fr_ntt: null pointer or n = 0");
  if (is_device_ptr(data)) {
    fr_ntt(ctx, data, n, inverse != 0);
  } else {
    DevBuf<uint32_t> d(ctx, n * 8);
    KB_CUDA(cudaMemcpyAsync(d.p, data, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    fr_ntt(ctx, d.p, n, inverse != 0);
    KB_CUDA(cudaMemcpyAsync(data, d.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  }
  KB_API_END(ctx)
}

static int32_t single_encrypt_batch(kb_ctx* ctx, const uint32_t com_xy[16], uint8_t com_inf, const uint32_t* points, const uint32_t* values,
                                    const uint32_t* r, const uint8_t* msgs, const uint64_t* msg_off, uint64_t n,
                                    uint32_t* ct_g2_xy, uint8_t* ct_inf, uint8_t* msg_ct) {
  KB_API_BEGIN(ctx)
  need(com_xy != nullptr, "kb_encrypt_batch: null commitment");
  need(n == 0 || (points && values && r && msg_off && ct_g2_xy && ct_inf), "kb_encrypt_batch: null pointer");
  need_srs(ctx);
  uint32_t com_host[16];
  KB_CUDA(cudaMemcpy(com_host, com_xy, 64, cudaMemcpyDefault));
  uint64_t total = 0;
  if (n) KB_CUDA(cudaMemcpy(&total, msg_off + n, 8, cudaMemcpyDefault));
  need(total == 0 || (msgs && msg_ct), "kb_encrypt_batch: null message buffer");
  DevIn<uint32_t> dp(ctx, points, n * 8), dv(ctx, values, n * 8), dr(ctx, r, n * 8);
  DevIn<uint8_t> dm(ctx, msgs, total);
  DevIn<uint64_t> doff(ctx, msg_off, n + 1);
  DevOut<uint32_t> oc(ctx, ct_g2_xy, n * 32);
  DevOut<uint8_t> oi(ctx, ct_inf, n), om(ctx, msg_ct, total);
  encrypt_batch(ctx, com_host, com_inf, dp, dv, dr, dm, doff, n, oc, oi, om);
  oc.finish(); oi.finish(); om.finish();
  KB_API_END(ctx)
}

static int32_t single_decrypt_batch(kb_ctx* ctx, const uint32_t* proofs_xy, const uint8_t* proofs_inf, const uint32_t* ct_g2_xy,
                                    const uint8_t* ct_inf, const uint8_t* msg_ct, const uint64_t* msg_off, uint64_t n, uint8_t* msgs_out) {
  KB_API_BEGIN(ctx)
  need(n == 0 || (proofs_xy && ct_g2_xy && msg_off), "kb_decrypt_batch: null pointer");
  uint64_t total = 0;
  if (n) KB_CUDA(cudaMemcpy(&total, msg_off + n, 8, cudaMemcpyDefault));
  need(total == 0 || (msg_ct && msgs_out), "kb_decrypt_batch: null message buffer");
  DevIn<uint32_t> dp(ctx, proofs_xy, n * 16), dc(ctx, ct_g2_xy, n * 32);
  DevIn<uint8_t> dpi(ctx, proofs_inf, n), dci(ctx, ct_inf, n), dm(ctx, msg_ct, total);
  DevIn<uint64_t> doff(ctx, msg_off, n + 1);
  DevOut<uint8_t> om(ctx, msgs_out, total);
  decrypt_batch(ctx, dp, dpi, dc, dci, dm, doff, n, om);
  om.finish();
  KB_API_END(ctx)
}

int32_t kb_pairing_batch(kb_ctx* ctx, const uint32_t* g1_xy, const uint8_t* g1_inf, const uint32_t* g2_xy, const uint8_t* g2_inf,
                         uint64_t n, uint8_t* gt_bytes) {
  KB_API_BEGIN(ctx)
  need(n == 0 || (g1_xy && g2_xy && gt_bytes), "kb_pairing_batch: null pointer");
  DevIn<uint32_t> a(ctx, g1_xy, n * 16), b(ctx, g2_xy, n * 32);
  DevIn<uint8_t> ai(ctx, g1_inf, n), bi(ctx, g2_inf, n);
  DevOut<uint8_t> o(ctx, gt_bytes, n * 384);
  pairing_batch(ctx, a, ai, b, bi, n, o);
  o.finish();
  KB_API_END(ctx)
}

int32_t kb_verify_batch(kb_ctx* ctx, const uint32_t* com_xy, const uint8_t* com_inf, const uint32_t* points, const uint32_t* values,
                        const uint32_t* proofs_xy, const uint8_t* proofs_inf, uint64_t n, uint8_t* ok) {
  KB_API_BEGIN(ctx)
  need(n == 0 || (com_xy && points && values && proofs_xy && ok), "kb_verify_batch: null pointer");
  need_srs(ctx);
  DevIn<uint32_t> c(ctx, com_xy, n * 16), z(ctx, points, n * 8), v(ctx, values, n * 8), p(ctx, proofs_xy, n * 16);
  DevIn<uint8_t> ci(ctx, com_inf, n), pi(ctx, proofs_inf, n);
  DevOut<uint8_t> o(ctx, ok, n);
  verify_batch(ctx, c, ci, z, v, p, pi, n, o);
  o.finish();
  KB_API_END(ctx)
}

int32_t kb_g1_serialize(kb_ctx* ctx, const uint32_t* xy, const uint8_t* inf, uint64_t n, int32_t compress, uint8_t* out) {
  KB_API_BEGIN(ctx)
  need(n == 0 || (xy && out), "kb_g1_serialize: null pointer");
  DevIn<uint32_t> p(ctx, xy, n * 16);
  DevIn<uint8_t> pi(ctx, inf, n);
  DevOut<uint8_t> o(ctx, out, n * (compress ? 32 : 64));
  g1_serialize(ctx, p, pi, n, compress ? 1 : 0, o);
  o.finish();
  KB_API_END(ctx)
}

int32_t kb_g2_serialize(kb_ctx* ctx, const uint32_t* xy, const uint8_t* inf, uint64_t n, int32_t compress, uint8_t* out) {
  KB_API_BEGIN(ctx)
  need(n == 0 || (xy && out), "kb_g2_serialize: null pointer");
  DevIn<uint32_t> p(ctx, xy, n * 32);
  DevIn<uint8_t> pi(ctx, inf, n);
  DevOut<uint8_t> o(ctx, out, n * (compress ? 64 : 128));
  g2_serialize(ctx, p, pi, n, compress ? 1 : 0, o);
  o.finish();
  KB_API_END(ctx)
}

int32_t kb_g1_deserialize(kb_ctx* ctx, const uint8_t* bytes, uint64_t n, int32_t compress, int32_t validate, uint32_t* xy, uint8_t* inf,
                          uint8_t* ok) {
  KB_API_BEGIN(ctx)
  need(n == 0 || (bytes && xy && inf && ok), "kb_g1_deserialize: null pointer");
  DevIn<uint8_t> b(ctx, bytes, n * (compress ? 32 : 64));
  DevOut<uint32_t> o(ctx, xy, n * 16);
  DevOut<uint8_t> oi(ctx, inf, n), ook(ctx, ok, n);
  g1_deserialize(ctx, b, n, compress ? 1 : 0, validate ? 1 : 0, o, oi, ook);
  o.finish(); oi.finish(); ook.finish();
  KB_API_END(ctx)
}

int32_t kb_g2_deserialize(kb_ctx* ctx, const uint8_t* bytes, uint64_t n, int32_t compress, int32_t validate, uint32_t* xy, uint8_t* inf,
                          uint8_t* ok) {
  KB_API_BEGIN(ctx)
  need(n == 0 || (bytes && xy && inf && ok), "kb_g2_deserialize: null pointer");
  DevIn<uint8_t> b(ctx, bytes, n * (compress ? 64 : 128));
  DevOut<uint32_t> o(ctx, xy, n * 32);
  DevOut<uint8_t> oi(ctx, inf, n), ook(ctx, ok, n);
  g2_deserialize(ctx, b, n, compress ? 1 : 0, validate ? 1 : 0, o, oi, ook);
  o.finish(); oi.finish(); ook.finish();
  KB_API_END(ctx)
}

int32_t kb_debug_fp_op(kb_ctx* ctx, int32_t field, int32_t op, const uint32_t* a, const uint32_t* b, uint32_t* out, uint64_t n) {
  KB_API_BEGIN(ctx)
  need(a && b && out, "kb_debug_fp_op: null pointer");
  DevIn<uint32_t> da(ctx, a, n * 8), db(ctx, b, n * 8);
  DevOut<uint32_t> o(ctx, out, n * 8);
  debug_fp_op(ctx, field, op, da, db, o, n);
  o.finish();
  KB_API_END(ctx)
}


// ------------------------------------------------------------------------------------------------------------------
// Multi-device contexts: ONE process drives all the GPUs of a box, one host thread + stream per device.
// Every device holds the whole SRS and its fixed-base tables (HBM capacity is what a B200 has to spare), so any call
// range splits evenly: `commit` by point range - each device returns one partial sum, the <= 8 partials (65 B each)
// are added on the first device - and the encrypt / decrypt batches by index with no exchange at all.  All other
// entry points run on the first device.  Peer access is enabled between the devices, so device-resident buffers of
// any of them are legal arguments (reads and writes go over NVLink).
// ------------------------------------------------------------------------------------------------------------------
}  // extern "C"
namespace {

template <class F>
int32_t on_all_devices(kb_ctx* ctx, F f) {   // f(k, sub-context) on one host thread per device; first failure wins
  const size_t nd = 1 + ctx->peers.size();
  std::vector<int32_t> rc(nd, KB_OK);
  for (size_t k = 1; k < nd; k++) ctx->workers[k - 1]->post([&rc, &f, ctx, k] { rc[k] = f(k, ctx->peers[k - 1]); });
  rc[0] = f(0, ctx);
  for (size_t k = 1; k < nd; k++) ctx->workers[k - 1]->wait();
  for (int i = 0; i < KB_T_COUNT; i++)
    for (kb_ctx* p : ctx->peers) if (p->last_ms[i] > ctx->last_ms[i]) ctx->last_ms[i] = p->last_ms[i];   // the slowest device
  for (size_t k = 1; k < nd; k++)
    if (rc[k] != KB_OK && rc[0] == KB_OK) { ctx->err = "device " + std::to_string(ctx->peers[k - 1]->device) + ": " + ctx->peers[k - 1]->err; return rc[k]; }
  return rc[0];
}
inline uint64_t shard_lo(uint64_t n, size_t k, size_t nd) { return n / nd * k + (k < n % nd ? k : n % nd); }
template <class T> inline const T* at(const T* p, uint64_t i) { return p ? p + i : nullptr; }
template <class T> inline T* at(T* p, uint64_t i) { return p ? p + i : nullptr; }
constexpr uint64_t MULTI_MIN_PER_DEVICE = 1024;   // below this a split costs more than it saves

}  // namespace
extern "C" {

int32_t kb_ctx_create_multi(const int32_t* devices, int32_t ndev, kb_ctx** out) {
  if (!out || !devices || ndev < 1 || ndev > 64) return KB_ERR_ARG;
  *out = nullptr;
  for (int a = 0; a < ndev; a++) for (int b = 0; b < a; b++) if (devices[a] == devices[b]) return KB_ERR_ARG;
  std::vector<kb_ctx*> subs(ndev, nullptr);
  std::vector<int32_t> rc(ndev, KB_OK);
  {
    std::vector<std::thread> th;   // the table builds of the devices run side by side
    for (int k = 1; k < ndev; k++) th.emplace_back([&, k] { rc[k] = kb_ctx_create(devices[k], &subs[k]); });
    rc[0] = kb_ctx_create(devices[0], &subs[0]);
    for (auto& t : th) t.join();
  }
  for (int k = 0; k < ndev; k++)
    if (rc[k] != KB_OK) { for (kb_ctx* c : subs) kb_ctx_destroy(c); return rc[k]; }
  for (int a = 0; a < ndev; a++) {   // peer access in both directions where the hardware offers it (NVLink / NVSwitch)
    cudaSetDevice(devices[a]);
    for (int b = 0; b < ndev; b++) {
      if (a == b) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) == cudaSuccess && can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) subs[0]->peer_access = false;
      } else subs[0]->peer_access = false;
      cudaGetLastError();
    }
  }
  subs[0]->peers.assign(subs.begin() + 1, subs.end());
  for (int k = 1; k < ndev; k++) subs[0]->workers.push_back(new DeviceWorker());
  *out = subs[0];
  return KB_OK;
}

int32_t kb_srs_upload(kb_ctx* ctx, const uint32_t* g1_aff_xy, uint64_t n, const uint32_t* tau_g2_xy) {
  if (!ctx) return KB_ERR_ARG;
  if (ctx->peers.empty()) return single_srs_upload(ctx, g1_aff_xy, n, tau_g2_xy);
  return on_all_devices(ctx, [&](size_t, kb_ctx* sub) { return single_srs_upload(sub, g1_aff_xy, n, tau_g2_xy); });
}

int32_t kb_srs_generate(kb_ctx* ctx, const uint32_t* tau, uint64_t first_power, uint64_t n, uint32_t* out_g1_xy, uint32_t* out_tau_g2_xy) {
  if (!ctx) return KB_ERR_ARG;
  if (ctx->peers.empty()) return single_srs_generate(ctx, tau, first_power, n, out_g1_xy, out_tau_g2_xy);
  return on_all_devices(ctx, [&](size_t k, kb_ctx* sub) {
    return single_srs_generate(sub, tau, first_power, n, k == 0 ? out_g1_xy : nullptr, k == 0 ? out_tau_g2_xy : nullptr);
  });
}

int32_t kb_msm_g1(kb_ctx* ctx, const uint32_t* scalars, uint64_t first, uint64_t n, uint32_t out_xy[16], uint8_t* out_inf) {
  if (!ctx) return KB_ERR_ARG;
  const size_t nd = 1 + ctx->peers.size();
  if (nd == 1 || n < MULTI_MIN_PER_DEVICE * nd || !out_xy || !scalars) return single_msm_g1(ctx, scalars, first, n, out_xy, out_inf);
  if (!ctx->peer_access && is_device_ptr(scalars)) return fail(ctx, KB_ERR_ARG, "kb_msm_g1: device buffers need peer access between the devices of a multi-device context");
  std::vector<uint32_t> part_xy(16 * nd);
  std::vector<uint8_t> part_inf(nd, 1);
  int32_t rc = on_all_devices(ctx, [&](size_t k, kb_ctx* sub) {
    const uint64_t lo = shard_lo(n, k, nd), hi = shard_lo(n, k + 1, nd);
    return single_msm_g1(sub, scalars + 8 * lo, first + lo, hi - lo, &part_xy[16 * k], &part_inf[k]);
  });
  if (rc != KB_OK) return rc;
  const float acc_ms = ctx->last_ms[KB_T_MSM_ACC], tot_ms = ctx->last_ms[KB_T_TOTAL];
  rc = kb_g1_sum(ctx, part_xy.data(), part_inf.data(), nd, out_xy, out_inf);
  ctx->last_ms[KB_T_MSM_ACC] = acc_ms;
  if (tot_ms >= 0 && ctx->last_ms[KB_T_TOTAL] >= 0) ctx->last_ms[KB_T_TOTAL] += tot_ms;   // slowest partial + the sum
  return rc;
}

int32_t kb_encrypt_batch(kb_ctx* ctx, const uint32_t com_xy[16], uint8_t com_inf, const uint32_t* points, const uint32_t* values,
                         const uint32_t* r, const uint8_t* msgs, const uint64_t* msg_off, uint64_t n,
                         uint32_t* ct_g2_xy, uint8_t* ct_inf, uint8_t* msg_ct) {
  if (!ctx) return KB_ERR_ARG;
  const size_t nd = 1 + ctx->peers.size();
  if (nd == 1 || n < MULTI_MIN_PER_DEVICE * nd || !msg_off || !com_xy)
    return single_encrypt_batch(ctx, com_xy, com_inf, points, values, r, msgs, msg_off, n, ct_g2_xy, ct_inf, msg_ct);
  if (!ctx->peer_access && (is_device_ptr(points) || is_device_ptr(ct_g2_xy)))
    return fail(ctx, KB_ERR_ARG, "kb_encrypt_batch: device buffers need peer access between the devices of a multi-device context");
  std::vector<uint64_t> off(n + 1);
  uint32_t com_host[16];
  cudaSetDevice(ctx->device);
  if (cudaMemcpy(off.data(), msg_off, (n + 1) * 8, cudaMemcpyDefault) != cudaSuccess || cudaMemcpy(com_host, com_xy, 64, cudaMemcpyDefault) != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, KB_ERR_CUDA, "kb_encrypt_batch: cannot read msg_off / commitment");
  }
  return on_all_devices(ctx, [&](size_t k, kb_ctx* sub) {
    const uint64_t lo = shard_lo(n, k, nd), hi = shard_lo(n, k + 1, nd), base = off[lo];
    std::vector<uint64_t> loc(hi - lo + 1);
    for (uint64_t i = lo; i <= hi; i++) loc[i - lo] = off[i] - base;
    return single_encrypt_batch(sub, com_host, com_inf, at(points, 8 * lo), at(values, 8 * lo), at(r, 8 * lo), at(msgs, base), loc.data(), hi - lo,
                                at(ct_g2_xy, 32 * lo), at(ct_inf, lo), at(msg_ct, base));
  });
}

int32_t kb_decrypt_batch(kb_ctx* ctx, const uint32_t* proofs_xy, const uint8_t* proofs_inf, const uint32_t* ct_g2_xy,
                         const uint8_t* ct_inf, const uint8_t* msg_ct, const uint64_t* msg_off, uint64_t n, uint8_t* msgs_out) {
  if (!ctx) return KB_ERR_ARG;
  const size_t nd = 1 + ctx->peers.size();
  if (nd == 1 || n < MULTI_MIN_PER_DEVICE * nd || !msg_off)
    return single_decrypt_batch(ctx, proofs_xy, proofs_inf, ct_g2_xy, ct_inf, msg_ct, msg_off, n, msgs_out);
  if (!ctx->peer_access && (is_device_ptr(proofs_xy) || is_device_ptr(msgs_out)))
    return fail(ctx, KB_ERR_ARG, "kb_decrypt_batch: device buffers need peer access between the devices of a multi-device context");
  std::vector<uint64_t> off(n + 1);
  cudaSetDevice(ctx->device);
  if (cudaMemcpy(off.data(), msg_off, (n + 1) * 8, cudaMemcpyDefault) != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, KB_ERR_CUDA, "kb_decrypt_batch: cannot read msg_off");
  }
  return on_all_devices(ctx, [&](size_t k, kb_ctx* sub) {
    const uint64_t lo = shard_lo(n, k, nd), hi = shard_lo(n, k + 1, nd), base = off[lo];
    std::vector<uint64_t> loc(hi - lo + 1);
    for (uint64_t i = lo; i <= hi; i++) loc[i - lo] = off[i] - base;
    return single_decrypt_batch(sub, at(proofs_xy, 16 * lo), at(proofs_inf, lo), at(ct_g2_xy, 32 * lo), at(ct_inf, lo), at(msg_ct, base), loc.data(),
                                hi - lo, at(msgs_out, base));
  });
}

}  // extern "C"
