// Test hook behind kb_debug_fp_op: elementwise field operations on the device, so the GPU tests can
// compare the PTX carry-chain primitives limb-for-limb with the oracle.
#include "ctx.cuh"

namespace kb {

template <class P>
__device__ __forceinline__ Fp<P> debug_apply(int op, const Fp<P>& x, const Fp<P>& y) {
  switch (op) {
    case 0: return fp_add<P>(x, y);
    case 1: return fp_sub<P>(x, y);
    case 2: return fp_mul_inl<P>(x, y);
    case 3: return fp_neg<P>(x);
    case 4: return fp_inv<P>(x);
    case 5: return fp_from_mont<P>(x);
    case 6: return fp_to_mont<P>(x);
    default: return fp_sqr<P>(x);
  }
}

__global__ void __launch_bounds__(128) debug_fp_kernel(int field, int op, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                       uint32_t* __restrict__ out, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (field == 0) fp_store<FqParams>(out + 8 * i, debug_apply<FqParams>(op, fp_load<FqParams>(a + 8 * i), fp_load<FqParams>(b + 8 * i)));
  else fp_store<FrParams>(out + 8 * i, debug_apply<FrParams>(op, fp_load<FrParams>(a + 8 * i), fp_load<FrParams>(b + 8 * i)));
}

void debug_fp_op(kb_ctx* ctx, int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, uint64_t n) {
  if (n) KB_LAUNCH(ctx, debug_fp_kernel, cdiv(n, 128), 128, 0, field, op, a, b, out, n);
}

}  // namespace kb
