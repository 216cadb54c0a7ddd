// Microbenchmark (round 2): Montgomery multiplication in a carry-free radix-2^29 representation (9 limbs, 64-bit column
// accumulators, plain IMAD.WIDE with no carry chain, R = 2^261) against the 32-bit-limb carry-chain multiplier of fp.cuh.
// Question: is the issue cost per product lower (plain IMAD.WIDE issues in 2.4 cycles, the carry-chained form in 4.4)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mb_f29 mb_f29.cu
#define KB_INLINE_ALL
#include "../../keaki_b200/csrc/fp.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace kb;

struct F29 { uint32_t v[9]; };
static constexpr uint32_t M29 = (1u << 29) - 1;
// q in radix 2^29 and -q^-1 mod 2^29 (filled by the host at start-up)
__constant__ uint32_t c_q29[9];
__constant__ uint32_t c_inv29;

__device__ __forceinline__ F29 f29_mul(const F29& a, const F29& b) {
  uint64_t c[18];
#pragma unroll
  for (int k = 0; k < 18; k++) c[k] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++)
#pragma unroll
    for (int j = 0; j < 9; j++) c[i + j] += (uint64_t)a.v[i] * b.v[j];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const uint32_t m = ((uint32_t)c[i] * c_inv29) & M29;
#pragma unroll
    for (int j = 0; j < 9; j++) c[i + j] += (uint64_t)m * c_q29[j];
    c[i + 1] += c[i] >> 29;
  }
  F29 r;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    r.v[k] = (uint32_t)c[9 + k] & M29;
    if (k < 8) c[10 + k] += c[9 + k] >> 29;
  }
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(512) k(const uint32_t* in, uint32_t* out, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (MODE == 0) {
    Fq x = fp_load<FqParams>(in + 8 * (t & 255)), y = fp_load<FqParams>(in + 8 * ((t + 1) & 255));
    Fq a = x, b = y;
    for (int i = 0; i < iters; i++) { a = fp_mul_inl<FqParams>(a, x); b = fp_mul_inl<FqParams>(b, y); }
    a = a + b;
    fp_store<FqParams>(out + 8 * t, a);
  } else {
    F29 x, y;
#pragma unroll
    for (int i = 0; i < 9; i++) { x.v[i] = in[(9 * t + i) & 2047] & M29; y.v[i] = in[(9 * t + i + 77) & 2047] & M29; }
    F29 a = x, b = y;
    for (int i = 0; i < iters; i++) { a = f29_mul(a, x); b = f29_mul(b, y); }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) s ^= a.v[i] + b.v[i];
    out[t] = s;
  }
}

int main() {
  // q and -q^-1 mod 2^29
  const uint64_t Q[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  uint32_t q29[9];
  for (int k = 0; k < 9; k++) {
    int bit = 29 * k, w = bit >> 6, sh = bit & 63;
    uint64_t v = Q[w] >> sh;
    if (sh > 35 && w + 1 < 4) v |= Q[w + 1] << (64 - sh);
    q29[k] = (uint32_t)v & M29;
  }
  uint32_t inv = 1;
  for (int i = 0; i < 6; i++) inv *= 2 - q29[0] * inv;   // q^-1 mod 2^32
  inv = (0u - inv) & M29;
  cudaMemcpyToSymbol(c_q29, q29, sizeof(q29));
  cudaMemcpyToSymbol(c_inv29, &inv, 4);
  uint32_t *in, *out;
  cudaMalloc(&in, 2048 * 4 * 4); cudaMalloc(&out, 148 * 4 * 512 * 32);
  cudaMemset(in, 0x5a, 2048 * 4 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  for (int mode = 0; mode < 2; mode++)
    for (int threads : {128, 256, 384, 512}) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148, threads>>>(in, out, iters); else k<1><<<148, threads>>>(in, out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      printf("%s  warps/scheduler %d: %.3f ms, %.3e products/s\n", mode == 0 ? "32-bit limbs, carry chains (fp_mul_inl)" : "29-bit limbs, carry-free columns      ",
             threads / 128, best, 2.0 * iters * 148 * threads / (best * 1e-3));
    }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return 0;
}
