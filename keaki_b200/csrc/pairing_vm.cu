// Batched pairing / decapsulation kernels built on the pairing VM (pairing_vm.cuh).
// `decapsulate` (src/kem.rs:55-72) + the XOR of `decrypt` (src/enc.rs:44-55) over the batch of
// `vec_decrypt` (src/vec.rs:72-81), and raw `E::pairing` (src/kem.rs:30,58; src/kzg.rs:148).
//
// Two adjacent lanes per pairing, lane t owning coordinate c_t of every Fq2 slot.  The slot file
// lives in shared memory as uint4 halves, [slot][half][thread], so a warp's access to one half is
// 512 contiguous bytes (conflict free, LDS.128 / STS.128).  Spilled values go to a global scratch
// laid out the same way ([gslot][half][thread]: coalesced, L2-resident).  GT never leaves the chip
// on the decrypt path: canonical bytes -> BLAKE3 XOF -> XOR happen in the epilogue.
#define KB_INLINE_ALL
#include "ctx.cuh"
#include "blake3.cuh"
#include "pairing_vm.cuh"
#include "pairing_prog_gen.cuh"
#include <cstdlib>

namespace kb {

template <int VM_BLOCK>   // threads per block = VM_BLOCK / 2 pairings (two lanes each)
struct DevLane {
  uint4* sm;        // shared slot file ([slot][half][thread]), already offset by threadIdx.x
  uint4* gl;        // global scratch ([gslot][half][thread]), already offset by the global thread index
  size_t gstride;   // threads in the launch (2 x pairings, padded to the block)
  uint32_t t;       // which Fq2 coordinate this lane owns

  __device__ __forceinline__ static Fq unpack(const uint4& q0, const uint4& q1) {
    Fq r;
    r.v[0] = q0.x; r.v[1] = q0.y; r.v[2] = q0.z; r.v[3] = q0.w;
    r.v[4] = q1.x; r.v[5] = q1.y; r.v[6] = q1.z; r.v[7] = q1.w;
    return r;
  }
  __device__ __forceinline__ Fq ld(uint32_t s) const {
    const uint4* p = sm + (size_t)s * 2 * VM_BLOCK;
    return unpack(p[0], p[VM_BLOCK]);
  }
  __device__ __forceinline__ Fq ld_partner(uint32_t s) const {   // the other coordinate of the same slot
    const uint4* p = sm + (size_t)s * 2 * VM_BLOCK + (t ? -1 : 1);
    return unpack(p[0], p[VM_BLOCK]);
  }
  __device__ __forceinline__ Fq ld_c(uint32_t s, uint32_t h) const {   // coordinate h of the slot, whichever lane asks
    const uint4* p = sm + (size_t)s * 2 * VM_BLOCK + ((int)h - (int)t);
    return unpack(p[0], p[VM_BLOCK]);
  }
  __device__ __forceinline__ void st(uint32_t s, const Fq& a) const {
    uint4* p = sm + (size_t)s * 2 * VM_BLOCK;
    p[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    p[VM_BLOCK] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
  }
  __device__ __forceinline__ Fq ldc(const uint32_t* p) const {
    const uint4* q = reinterpret_cast<const uint4*>(p + 8 * t);
    return unpack(__ldg(q), __ldg(q + 1));
  }
  __device__ __forceinline__ Fq ldg(uint32_t g) const {
    const uint4* p = gl + (size_t)g * 2 * gstride;
    return unpack(p[0], p[gstride]);
  }
  __device__ __forceinline__ void stg(uint32_t g, const Fq& a) const {
    uint4* p = gl + (size_t)g * 2 * gstride;
    p[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    p[gstride] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
  }
  __device__ __forceinline__ Fq xchg(const Fq& a) const {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, a.v[i], 1);
    return r;
  }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};

// mode 0: write the 96 canonical GT words; mode 1: key = BLAKE3-XOF(GT bytes), out = key XOR msg_ct.
// A warp serves 16 pairings: lane = pairing_in_warp * 2 + t.  MINB = resident blocks per SM the register
// allocation is held to (the slot budget decides how many fit in shared memory).  The launch shape is chosen so
// that the batch divides into whole rounds of resident pairings (vm_pick_shape).
template <int VM_BLOCK, int MINB>
__global__ void __launch_bounds__(VM_BLOCK, MINB) pairing_vm_kernel(const uint64_t* __restrict__ prog, const uint32_t* __restrict__ consts,
                                                                    uint64_t out_slots, const uint32_t* __restrict__ g1,
                                                                    const uint8_t* __restrict__ g1_inf, const uint32_t* __restrict__ g2,
                                                                    const uint8_t* __restrict__ g2_inf, uint64_t n, uint4* __restrict__ scratch,
                                                                    uint64_t gstride, int mode, uint32_t* __restrict__ gt_words,
                                                                    const uint8_t* __restrict__ msg_ct, const uint64_t* __restrict__ off,
                                                                    uint8_t* __restrict__ out) {
  extern __shared__ uint4 vm_smem[];
  const uint64_t gtid = blockIdx.x * (uint64_t)VM_BLOCK + threadIdx.x;
  const uint64_t pairing = gtid >> 1;
  const bool live = pairing < n;
  const uint64_t i = live ? pairing : n - 1;   // padding lanes recompute the last pairing (shuffles need all lanes)
  DevLane<VM_BLOCK> ln;
  ln.t = threadIdx.x & 1u;
  ln.sm = vm_smem + threadIdx.x;
  ln.gl = scratch + gtid;
  ln.gstride = gstride;

  // slot 0 = (xP, yP), slot 1 = Q.x, slot 2 = Q.y: lane t takes coordinate t of each
  Fq pc = fp_load<FqParams>(g1 + 16 * i + 8 * ln.t);
  Fq qx = fp_load<FqParams>(g2 + 32 * i + 8 * ln.t);
  Fq qy = fp_load<FqParams>(g2 + 32 * i + 16 + 8 * ln.t);
  uint32_t pz = pc.is_zero(), qz = qx.is_zero() && qy.is_zero();
  pz &= __shfl_xor_sync(0xffffffffu, pz, 1);
  qz &= __shfl_xor_sync(0xffffffffu, qz, 1);
  const bool trivial = (g1_inf && g1_inf[i]) || (g2_inf && g2_inf[i]) || pz || qz;
  ln.st(0, pc);
  ln.st(1, qx);
  ln.st(2, qy);

  vm::run(prog, ln, consts);
  __syncwarp();

  // canonical (non-Montgomery) limbs of this lane's six coordinates, in ark-serialize order
  uint32_t mine[48];
#pragma unroll 1
  for (int k = 0; k < 6; k++) {
    Fq c = ln.ld((uint32_t)(out_slots >> (8 * k)) & 255u);
    if (trivial) c = (k == 0 && ln.t == 0) ? Fq::one() : Fq::zero();   // arkworks skips pairs with an infinity: GT = 1
    Fq unit = Fq::zero(); unit.v[0] = 1;
    c = vm::mul1(c, unit);   // Montgomery form -> canonical integer
#pragma unroll
    for (int j = 0; j < 8; j++) mine[8 * k + j] = c.v[j];
  }
  if (mode == 0) {
    if (live) {
      uint4* o = reinterpret_cast<uint4*>(gt_words + 96 * i + 8 * ln.t);
#pragma unroll
      for (int k = 0; k < 6; k++) {
        o[4 * k] = make_uint4(mine[8 * k], mine[8 * k + 1], mine[8 * k + 2], mine[8 * k + 3]);
        o[4 * k + 1] = make_uint4(mine[8 * k + 4], mine[8 * k + 5], mine[8 * k + 6], mine[8 * k + 7]);
      }
    }
  } else {
    uint32_t w[96];
#pragma unroll
    for (int k = 0; k < 6; k++)
#pragma unroll
      for (int j = 0; j < 8; j++) {
        uint32_t other = __shfl_xor_sync(0xffffffffu, mine[8 * k + j], 1);
        w[16 * k + j] = ln.t ? other : mine[8 * k + j];
        w[16 * k + 8 + j] = ln.t ? mine[8 * k + j] : other;
      }
    if (live && ln.t == 0) {
      uint64_t lo = off[i], hi = off[i + 1];
      b3_gt_xof_xor(w, msg_ct + lo, out + lo, hi - lo);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// slot file of one block: slots x 2 halves x threads x 16 B
static int vm_smem_bytes(const kb_ctx* ctx, int block) { return ctx->vm_slots * 2 * block * 16; }

struct VmShape { int block, minb; };
// Launch shapes compiled below.  64-thread blocks allow 7 per SM (14 warps: 224 resident pairings), which turns
// 2^16 pairings on 148 SMs into 1.98 rounds instead of 2.3 rounds of 192.
static const VmShape VM_SHAPES[] = {{128, 2}, {128, 3}, {128, 4}, {64, 7}};

template <int BLOCK, int MINB>
static void vm_prepare(const kb_ctx* ctx) {
  KB_CUDA(cudaFuncSetAttribute(pairing_vm_kernel<BLOCK, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, vm_smem_bytes(ctx, BLOCK)));
}

// Program choice: KB_PAIRING_SLOTS / KB_PAIRING_SHAPE="block,minb" (defaults: the fastest measured variant)
void vm_init(kb_ctx* ctx) {
  using namespace vmprog;
  int slots = 18;
  ctx->vm_block = 128; ctx->vm_minb = 3;   // 12 warps / SM, 168 registers; measured fastest
  if (const char* e = getenv("KB_PAIRING_SLOTS")) slots = atoi(e);
  if (const char* e = getenv("KB_PAIRING_SHAPE")) sscanf(e, "%d,%d", &ctx->vm_block, &ctx->vm_minb);
  const Program* pr = nullptr;
  for (int k = 0; k < NUM_PROGRAMS; k++) if (PROGRAMS[k].slots == slots) pr = &PROGRAMS[k];
  if (!pr) throw CudaError("KB_PAIRING_SLOTS names a program variant that was not generated");
  bool shape_ok = false;
  for (const VmShape& sh : VM_SHAPES) shape_ok |= sh.block == ctx->vm_block && sh.minb == ctx->vm_minb;
  if (!shape_ok) throw CudaError("KB_PAIRING_SHAPE names a launch shape that was not compiled");
  ctx->vm_slots = pr->slots;
  ctx->vm_gslots = pr->gslots;
  if ((vm_smem_bytes(ctx, ctx->vm_block) + 1024) * ctx->vm_minb > 228 * 1024)
    throw CudaError("pairing VM: slot files of the requested blocks per SM do not fit shared memory");
  ctx->vm_out = 0;
  for (int k = 0; k < 6; k++) ctx->vm_out |= (uint64_t)pr->out[k] << (8 * k);
  const size_t bytes = (size_t)(pr->len + 2) * 8;   // one padding instruction after END (prefetch)
  KB_CUDA(cudaMalloc((void**)&ctx->d_vm_prog, bytes));
  KB_CUDA(cudaMemsetAsync(ctx->d_vm_prog, 0, bytes, ctx->stream));
  KB_CUDA(cudaMemcpyAsync(ctx->d_vm_prog, pr->words, (size_t)pr->len * 8, cudaMemcpyHostToDevice, ctx->stream));
  KB_CUDA(cudaMalloc((void**)&ctx->d_vm_consts, sizeof(CONSTS)));
  KB_CUDA(cudaMemcpyAsync(ctx->d_vm_consts, CONSTS, sizeof(CONSTS), cudaMemcpyHostToDevice, ctx->stream));
  vm_prepare<128, 2>(ctx); vm_prepare<128, 3>(ctx); vm_prepare<128, 4>(ctx); vm_prepare<64, 7>(ctx);
}

void vm_free(kb_ctx* ctx) {
  cudaFree(ctx->d_vm_prog); cudaFree(ctx->d_vm_consts);
  ctx->d_vm_prog = nullptr; ctx->d_vm_consts = nullptr;
}

static void vm_launch(kb_ctx* ctx, const uint32_t* d_g1, const uint8_t* d_g1_inf, const uint32_t* d_g2, const uint8_t* d_g2_inf,
                      uint64_t n, int mode, uint32_t* d_gt_words, const uint8_t* d_msg_ct, const uint64_t* d_off, uint8_t* d_out) {
  const int block = ctx->vm_block;
  const unsigned blocks = cdiv(2 * n, block);
  const uint64_t gstride = (uint64_t)blocks * block;
  DevBuf<uint4> scratch(ctx, (size_t)(ctx->vm_gslots ? ctx->vm_gslots : 1) * 2 * gstride);
  timer_start(ctx, KB_T_PAIRING);
#define KB_VM_GO(B, MB) KB_LAUNCH(ctx, (pairing_vm_kernel<B, MB>), blocks, B, vm_smem_bytes(ctx, B), ctx->d_vm_prog, ctx->d_vm_consts, \
            ctx->vm_out, d_g1, d_g1_inf, d_g2, d_g2_inf, n, scratch.p, gstride, mode, d_gt_words, d_msg_ct, d_off, d_out)
  if (block == 64) KB_VM_GO(64, 7);
  else if (ctx->vm_minb >= 4) KB_VM_GO(128, 4);
  else if (ctx->vm_minb == 3) KB_VM_GO(128, 3);
  else KB_VM_GO(128, 2);
#undef KB_VM_GO
  timer_stop(ctx, KB_T_PAIRING);
}

void pairing_batch(kb_ctx* ctx, const uint32_t* d_g1, const uint8_t* d_g1_inf, const uint32_t* d_g2, const uint8_t* d_g2_inf,
                   uint64_t n, uint8_t* d_gt_bytes) {
  if (!n) return;
  if (ctx->pairing_impl == 1) return st_pairing_launch(ctx, d_g1, d_g1_inf, d_g2, d_g2_inf, n, 0, reinterpret_cast<uint32_t*>(d_gt_bytes), nullptr, nullptr, nullptr);
  vm_launch(ctx, d_g1, d_g1_inf, d_g2, d_g2_inf, n, 0, reinterpret_cast<uint32_t*>(d_gt_bytes), nullptr, nullptr, nullptr);
}

void decrypt_batch(kb_ctx* ctx, const uint32_t* d_proofs, const uint8_t* d_pinf, const uint32_t* d_ct, const uint8_t* d_cinf,
                   const uint8_t* d_msg_ct, const uint64_t* d_off, uint64_t n, uint8_t* d_out) {
  if (!n) return;
  if (ctx->pairing_impl == 1) return st_pairing_launch(ctx, d_proofs, d_pinf, d_ct, d_cinf, n, 1, nullptr, d_msg_ct, d_off, d_out);
  vm_launch(ctx, d_proofs, d_pinf, d_ct, d_cinf, n, 1, nullptr, d_msg_ct, d_off, d_out);
}

}  // namespace kb
