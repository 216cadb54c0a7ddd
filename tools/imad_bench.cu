// Integer-pipe throughput microbenchmark for sm_100a (B200).
//
// Measures the issue rate of the instructions the BN254 field multiplier is built from, so that
// the IMAD roofline denominator in bench.py is a MEASURED number ("IMAD peak is not
// in MEASURED_PEAKS.json: the builder must add a microbenchmark").
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/imad_bench tools/imad_bench.cu
// Run:    tools/imad_bench [json-out]
//
// Every kernel runs CHAINS independent dependency chains per thread, ITER iterations, all SMs,
// 8 CTAs x 256 threads per SM, and is timed with CUDA events. "per_clk_sm" uses the SM clock
// sampled through clock64() deltas in the same kernel (cycles elapsed on one SM).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int CHAINS = 8;
constexpr int ITER = 32768;

enum Kind { K_IMAD_LO = 0, K_IMAD_HI, K_IMAD_WIDE, K_IMAD_WIDE_CC, K_IADD3, K_MIX_WIDE_IADD, K_DFMA, K_LOHI_PAIR, K_COUNT };
static const char* kind_name[K_COUNT] = {"imad_lo", "imad_hi", "imad_wide", "imad_wide_carry", "iadd3",
                                         "mix_wide+iadd3", "dfma", "imad_lo+imad_hi"};
// instructions of the measured class issued per chain per iteration
static const int ops_per_chain_iter[K_COUNT] = {1, 1, 1, 2, 1, 1, 1, 2};

template <int KIND>
__global__ void __launch_bounds__(256) bench_kernel(uint32_t* out, uint32_t seed, long long* cycles) {
  uint32_t a = seed * (threadIdx.x + 1u) | 1u, b = seed ^ 0x9e3779b9u;
  uint32_t lo[CHAINS], hi[CHAINS];
  uint64_t w[CHAINS];
  double d[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) { lo[c] = a + c; hi[c] = b + c; d[c] = (double)(c + 1) * 1e-3; w[c] = ((uint64_t)(b + c) << 32) | (a + c); }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      if (KIND == K_IMAD_LO) {
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[c]) : "r"(a), "r"(b));
      } else if (KIND == K_IMAD_HI) {
        asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo[c]) : "r"(a), "r"(b));
      } else if (KIND == K_IMAD_WIDE) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a), "r"(b));
      } else if (KIND == K_IMAD_WIDE_CC) {
        // two lo/hi pairs linked by the carry flag: what the Montgomery rows look like
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;\n\t"
                     "madc.lo.cc.u32 %0, %2, %4, %0;\n\tmadc.hi.u32 %1, %2, %4, %1;"
                     : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b), "r"(seed));
      } else if (KIND == K_IADD3) {
        asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[c]) : "r"(hi[c]));
      } else if (KIND == K_MIX_WIDE_IADD) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a), "r"(b));
      } else if (KIND == K_DFMA) {
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(1.0000001), "d"(1e-9));
      } else if (KIND == K_LOHI_PAIR) {
        asm volatile("mad.lo.u32 %0, %0, %2, %3;\n\tmad.hi.u32 %1, %1, %2, %3;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(a), "r"(b));
      }
    }
    if (KIND == K_MIX_WIDE_IADD) {
      // CHAINS extra independent adds on the ALU pipe per iteration (1 per wide mad)
      uint32_t x = a;
#pragma unroll
      for (int c = 0; c < CHAINS; c++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b + c));
      a = x | 1u;
    }
  }
  long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) r ^= lo[c] ^ hi[c] ^ (uint32_t)__double2loint(d[c]) ^ (uint32_t)w[c] ^ (uint32_t)(w[c] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

static double g_clock_khz = 1.0;
template <int KIND>
static void run(int sms, uint32_t* dout, long long* dcyc, FILE* js, bool last) {
  int ctas = sms * 8;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int wu = 0; wu < 3; wu++) bench_kernel<KIND><<<ctas, 256>>>(dout, 12345u + wu, dcyc);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  long long cyc = 0;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(e0));
    bench_kernel<KIND><<<ctas, 256>>>(dout, 777u + rep, dcyc);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) { best = ms; CK(cudaMemcpy(&cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost)); }
  }
  double ops = (double)ctas * 256.0 * CHAINS * ITER * ops_per_chain_iter[KIND];
  double per_s = ops / (best * 1e-3);
  // per-SM per-clock rate from the in-kernel cycle count of CTA 0 (8 CTAs of identical work share each SM)
  (void)cyc;
  double mhz = g_clock_khz / 1e3;  // max SM clock reported by the driver; the achieved clock is sampled by nvidia-smi outside
  double per_clk_sm = per_s / ((double)sms * g_clock_khz * 1e3);
  printf("%-18s  %8.3f ms  %10.4e ops/s  %7.2f ops/clk/SM  (~%.0f MHz)\n", kind_name[KIND], best, per_s, per_clk_sm, mhz);
  if (js) fprintf(js, "  \"%s\": {\"ops_per_s\": %.6e, \"ops_per_clk_sm\": %.3f, \"ms\": %.4f, \"approx_mhz\": %.0f}%s\n",
                  kind_name[KIND], per_s, per_clk_sm, best, mhz, last ? "" : ",");
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  { int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0)); g_clock_khz = khz; }
  printf("device: %s, %d SMs, cc %d.%d\n", prop.name, sms, prop.major, prop.minor);
  uint32_t* dout; long long* dcyc;
  CK(cudaMalloc(&dout, (size_t)sms * 8 * 256 * 4)); CK(cudaMalloc(&dcyc, 8));
  FILE* js = argc > 1 ? fopen(argv[1], "w") : nullptr;
  if (js) fprintf(js, "{\n  \"device\": \"%s\", \"sms\": %d,\n", prop.name, sms);
  run<K_IMAD_LO>(sms, dout, dcyc, js, false);
  run<K_IMAD_HI>(sms, dout, dcyc, js, false);
  run<K_LOHI_PAIR>(sms, dout, dcyc, js, false);
  run<K_IMAD_WIDE>(sms, dout, dcyc, js, false);
  run<K_IMAD_WIDE_CC>(sms, dout, dcyc, js, false);
  run<K_IADD3>(sms, dout, dcyc, js, false);
  run<K_MIX_WIDE_IADD>(sms, dout, dcyc, js, false);
  run<K_DFMA>(sms, dout, dcyc, js, true);
  if (js) { fprintf(js, "}\n"); fclose(js); }
  return 0;
}
