// Microbenchmark (round 2): how much instruction-level parallelism does ptxas extract from independent carry-chain
// streams?  The pairing kernel is issue/latency-bound (lone-warp CPI 3.8, 53 % "wait" stalls): if two independent lazy Fq2
// products placed in ONE function body ran in the time of one, a paired routine (f2mul2) would shorten every Fq6 product.
// Modes: 0 = K dependent rounds of ONE lazy Fq2 product (x = x * y);  1 = two independent streams per round through two
// noinline calls;  2 = the two streams inlined in one body (ptxas free to interleave);  3 = four streams in one body.
// Grid: 148 blocks x WARPS warps.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mb_ilp mb_ilp.cu
#include "../../keaki_b200/csrc/fpl.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace kb;

static __device__ __noinline__ Fq2 f2mul_call(Fq2 a, Fq2 b) { return lz::fq2_mul_lazy(a, b); }

template <int MODE>
__global__ void __launch_bounds__(256) k(const uint32_t* in, uint32_t* out, int iters) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fq2 x[4], y;
#pragma unroll
  for (int s = 0; s < 4; s++) { x[s].c0 = fp_load<FqParams>(in + 8 * ((t + s) & 255)); x[s].c1 = fp_load<FqParams>(in + 8 * ((t + s + 9) & 255)); }
  y.c0 = fp_load<FqParams>(in + 8 * ((t + 5) & 255)); y.c1 = fp_load<FqParams>(in + 8 * ((t + 6) & 255));
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) { x[0] = f2mul_call(x[0], y); }
    if (MODE == 1) { x[0] = f2mul_call(x[0], y); x[1] = f2mul_call(x[1], y); }
    if (MODE == 2) { const Fq2 r0 = lz::fq2_mul_lazy(x[0], y), r1 = lz::fq2_mul_lazy(x[1], y); x[0] = r0; x[1] = r1; }
    if (MODE == 3) {
      const Fq2 r0 = lz::fq2_mul_lazy(x[0], y), r1 = lz::fq2_mul_lazy(x[1], y), r2 = lz::fq2_mul_lazy(x[2], y), r3 = lz::fq2_mul_lazy(x[3], y);
      x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
    }
  }
  Fq2 s = x[0] + x[1] + x[2] + x[3];
  fp_store<FqParams>(out + 16 * t, s.c0); fp_store<FqParams>(out + 16 * t + 8, s.c1);
}

int main() {
  uint32_t h[8 * 256];
  for (int i = 0; i < 8 * 256; i++) h[i] = (i % 8 == 7) ? 0x10000000u + i : 0x9e3779b9u * (i + 1);
  uint32_t *din, *dout;
  cudaMalloc(&din, sizeof(h)); cudaMalloc(&dout, 16 * 4 * 148 * 256);
  cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps : {1, 4, 8}) {
    for (int mode = 0; mode < 4; mode++) {
      float best = 1e9f;
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148, 32 * warps>>>(din, dout, iters);
        if (mode == 1) k<1><<<148, 32 * warps>>>(din, dout, iters);
        if (mode == 2) k<2><<<148, 32 * warps>>>(din, dout, iters);
        if (mode == 3) k<3><<<148, 32 * warps>>>(din, dout, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      const int streams = mode == 0 ? 1 : mode == 3 ? 4 : 2;
      printf("warps/SM %d mode %d: %.3f ms, %.1f ns per product per warp, %.2f warp-products/us/SM (x 32 per thread)\n", warps, mode, best,
             best * 1e6 / iters / streams, (double)iters * streams * warps / (best * 1e3));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
