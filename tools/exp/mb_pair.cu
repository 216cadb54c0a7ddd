// Multiplier microbenchmark (round 1): throughput of chains of Montgomery products per thread at 1..4 warps per scheduler,
// sequential vs. two products interleaved step by step.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mb_pair mb_pair.cu
// Result on B200: 4.38e10 / 6.53e10 / 6.72e10 / 6.78e10 Fq products/s at 1 / 2 / 3 / 4 warps per scheduler; interleaving changes nothing.
#define KB_INLINE_ALL
#include "../../keaki_b200/csrc/fp.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace kb;

template <class P>
__device__ __forceinline__ void fp_mul_pair(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d, Fp<P>& r1, Fp<P>& r2) {
  uint32_t e1[8], o1[8], e2[8], o2[8];
  mont_step<P, true>(e1, o1, a.v, b.v[0]);  mont_step<P, true>(e2, o2, c.v, d.v[0]);
  mont_step<P, false>(o1, e1, a.v, b.v[1]); mont_step<P, false>(o2, e2, c.v, d.v[1]);
  mont_step<P, false>(e1, o1, a.v, b.v[2]); mont_step<P, false>(e2, o2, c.v, d.v[2]);
  mont_step<P, false>(o1, e1, a.v, b.v[3]); mont_step<P, false>(o2, e2, c.v, d.v[3]);
  mont_step<P, false>(e1, o1, a.v, b.v[4]); mont_step<P, false>(e2, o2, c.v, d.v[4]);
  mont_step<P, false>(o1, e1, a.v, b.v[5]); mont_step<P, false>(o2, e2, c.v, d.v[5]);
  mont_step<P, false>(e1, o1, a.v, b.v[6]); mont_step<P, false>(e2, o2, c.v, d.v[6]);
  mont_step<P, false>(o1, e1, a.v, b.v[7]); mont_step<P, false>(o2, e2, c.v, d.v[7]);
  r1.v[0] = add_cc(e1[0], o1[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r1.v[i] = addc_cc(e1[i], o1[i + 1]);
  r1.v[7] = addc(e1[7], 0);
  r2.v[0] = add_cc(e2[0], o2[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r2.v[i] = addc_cc(e2[i], o2[i + 1]);
  r2.v[7] = addc(e2[7], 0);
  fp_reduce_once<P>(r1.v);
  fp_reduce_once<P>(r2.v);
}

template <int MODE>
__global__ void __launch_bounds__(512) k(const uint32_t* in, uint32_t* out, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fq x = fp_load<FqParams>(in + 8 * (t & 255)), y = fp_load<FqParams>(in + 8 * ((t + 1) & 255));
  Fq a = x, b = y;
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) { a = fp_mul_inl<FqParams>(a, x); b = fp_mul_inl<FqParams>(b, y); }
    else { Fq r1, r2; fp_mul_pair<FqParams>(a, x, b, y, r1, r2); a = r1; b = r2; }
  }
  a = a + b;
  fp_store<FqParams>(out + 8 * t, a);
}

int main() {
  uint32_t *in, *out;
  cudaMalloc(&in, 256 * 32); cudaMalloc(&out, 148 * 4 * 512 * 32);
  cudaMemset(in, 0x5a, 256 * 32);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  for (int mode = 0; mode < 2; mode++)
    for (int threads : {128, 256, 384, 512})
      for (int bps : {1}) {
        for (int rep = 0; rep < 2; rep++) {
          cudaEventRecord(e0);
          if (mode == 0) k<0><<<148 * bps, threads>>>(in, out, iters); else k<1><<<148 * bps, threads>>>(in, out, iters);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (rep) printf("mode %d threads/SM %d: %.3f ms, %.3e Fq-mul/s, %.1f cycles/mul/warp(@1.965GHz)\n", mode, threads * bps, ms,
                          2.0 * iters * 148 * bps * threads / (ms * 1e-3), ms * 1e-3 * 1.965e9 / (2.0 * iters));
        }
      }
  uint32_t h[8]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost); printf("%08x\n", h[0]);
  return 0;
}
