// Microbenchmark (round 2): the two ways to add a stream of points into bucket accumulators, under IDEAL operand supply
// (everything in registers, no sort, no memory traffic), to decide the MSM's accumulation form with measured numbers:
//   (a) XYZZ mixed addition, one accumulator per thread (what msm_accumulate_kernel runs): 8M + 2S, no inversion;
//   (b) batch-affine addition: every thread holds B affine accumulators and adds one affine point to each; the B x 32
//       denominators of a warp share ONE field inversion through Montgomery's trick - thread-local prefix products, a
//       prefix and a suffix product scan over the 32 lanes by shuffles, one inversion, back-substitution - then
//       lambda = dy / dx, x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1  (1S + 2M + the trick's 3M).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mb_affine mb_affine.cu        Run: ./mb_affine
// The coordinates are random field elements (not curve points): the instruction streams are the same and no special case
// (equal x) can trigger.
#define KB_INLINE_ALL
#include "../../keaki_b200/csrc/ec.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace kb;

__device__ __forceinline__ Fq shfl_fq(const Fq& a, int src) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
  return r;
}
__device__ __forceinline__ Fq shfl_up_fq(const Fq& a, int d) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_up_sync(0xffffffffu, a.v[i], d);
  return r;
}
__device__ __forceinline__ Fq shfl_down_fq(const Fq& a, int d) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], d);
  return r;
}

__global__ void __launch_bounds__(256, 2) k_xyzz(const uint32_t* in, uint32_t* out, int iters) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  G1 acc;
  acc.x = fp_load<FqParams>(in + 8 * (t & 1023)); acc.y = fp_load<FqParams>(in + 8 * ((t + 1) & 1023));
  acc.zz = fp_load<FqParams>(in + 8 * ((t + 2) & 1023)); acc.zzz = fp_load<FqParams>(in + 8 * ((t + 3) & 1023));
  G1Affine p;
  p.x = fp_load<FqParams>(in + 8 * ((t + 4) & 1023)); p.y = fp_load<FqParams>(in + 8 * ((t + 5) & 1023));
  for (int i = 0; i < iters; i++) {
    acc = ec_add_mixed(acc, p);
    p.x = p.x + acc.y;   // a new point every time (one modular addition, as cheap as the real kernel's negation)
  }
  fp_store<FqParams>(out + 8 * t, acc.x + acc.y + acc.zz + acc.zzz);
}

template <int B>
__global__ void __launch_bounds__(256, 2) k_affine(const uint32_t* in, uint32_t* out, int iters) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  Fq ax[B], ay[B], px[B], py[B];
#pragma unroll
  for (int b = 0; b < B; b++) {
    ax[b] = fp_load<FqParams>(in + 8 * ((t + 7 * b) & 1023)); ay[b] = fp_load<FqParams>(in + 8 * ((t + 7 * b + 1) & 1023));
    px[b] = fp_load<FqParams>(in + 8 * ((t + 7 * b + 2) & 1023)); py[b] = fp_load<FqParams>(in + 8 * ((t + 7 * b + 3) & 1023));
  }
  for (int it = 0; it < iters; it++) {
    Fq d[B], pre[B];
#pragma unroll
    for (int b = 0; b < B; b++) { d[b] = px[b] - ax[b]; pre[b] = b == 0 ? d[0] : pre[b - 1] * d[b]; }
    // inclusive prefix and suffix products of the thread totals over the warp
    Fq up = pre[B - 1], dn = pre[B - 1];
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      Fq u = shfl_up_fq(up, s), w = shfl_down_fq(dn, s);
      if (lane >= s) up = up * u;
      if (lane + s < 32) dn = dn * w;
    }
    const Fq inv_all = inv(shfl_fq(up, 31));             // every lane runs the same inversion (SIMT): one per warp in effect
    Fq left = shfl_up_fq(up, 1), right = shfl_down_fq(dn, 1);
    Fq mine = inv_all;                                     // inverse of this thread's total = inv_all * prod(others)
    if (lane > 0) mine = mine * left;
    if (lane < 31) mine = mine * right;
    // back-substitution inside the thread, then the additions
#pragma unroll
    for (int b = B - 1; b >= 0; b--) {
      const Fq di = b == 0 ? mine : mine * pre[b - 1];     // 1 / d[b]
      if (b > 0) mine = mine * d[b];
      const Fq lam = (py[b] - ay[b]) * di;
      const Fq x3 = sqr(lam) - ax[b] - px[b];
      ay[b] = lam * (ax[b] - x3) - ay[b];
      ax[b] = x3;
      px[b] = px[b] + ay[b];
    }
  }
  Fq s = Fq::zero();
#pragma unroll
  for (int b = 0; b < B; b++) s = s + ax[b] + ay[b];
  fp_store<FqParams>(out + 8 * t, s);
}

template <class K>
static double run(K kern, const uint32_t* in, uint32_t* out, int iters, int per_thread, const char* name) {
  const int blocks = 148 * 2, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(in, out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double adds = (double)blocks * threads * iters * per_thread;
  printf("%-44s %8.3f ms  %.3e additions/s\n", name, best, adds / (best * 1e-3));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return adds / (best * 1e-3);
}

int main() {
  uint32_t *in, *out;
  cudaMalloc(&in, 1024 * 32); cudaMalloc(&out, 148 * 2 * 256 * 32);
  uint32_t h[1024 * 8];
  uint64_t x = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < 1024 * 8; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = (uint32_t)x; if ((i & 7) == 7) h[i] &= 0x0fffffffu; }
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const double a = run(k_xyzz, in, out, 256, 1, "XYZZ mixed addition (8M + 2S)");
  const double b1 = run(k_affine<1>, in, out, 64, 1, "batch-affine, 1 per thread x 32 lanes / inversion");
  const double b2 = run(k_affine<2>, in, out, 64, 2, "batch-affine, 2 per thread x 32 lanes / inversion");
  const double b4 = run(k_affine<4>, in, out, 64, 4, "batch-affine, 4 per thread x 32 lanes / inversion");
  const double b8 = run(k_affine<8>, in, out, 32, 8, "batch-affine, 8 per thread x 32 lanes / inversion");
  printf("ratio to XYZZ: %.2f %.2f %.2f %.2f\n", b1 / a, b2 / a, b4 / a, b8 / a);
  return 0;
}
