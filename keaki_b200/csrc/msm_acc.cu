// MSM bucket accumulation — the dominant kernel of `commit` (src/kzg.rs:98).  Hot translation unit:
// the field multiplier and the XYZZ mixed addition are fully inlined here (KB_INLINE_ALL).
#define KB_INLINE_ALL
#include "ctx.cuh"

namespace kb {

// ------------------------------------------------------------------------------------------
// bucket accumulation: one thread per bucket, XYZZ mixed additions, next base prefetched
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) msm_accumulate_kernel(const uint32_t* __restrict__ tab, uint64_t tab_n, uint64_t first,
                                                                const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ entries,
                                                                uint32_t nb, uint32_t* __restrict__ buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint32_t lo = offsets[b], hi = offsets[b + 1];
  G1 acc = G1::infinity();
  if (lo < hi) {
    uint32_t e = entries[lo];
    G1Affine nxt = ld_g1(tab + 16 * ((uint64_t)((e >> 26) & 31u) * tab_n + first + (e & 0x3ffffffu)));
    for (uint32_t k = lo; k < hi; k++) {
      G1Affine cur = nxt;
      bool negate = (e >> 31) != 0;
      if (k + 1 < hi) {
        e = entries[k + 1];
        nxt = ld_g1(tab + 16 * ((uint64_t)((e >> 26) & 31u) * tab_n + first + (e & 0x3ffffffu)));
      }
      if (negate) cur.y = -cur.y;
      acc = ec_add_mixed(acc, cur);
    }
  }
  st_g1x(buckets + 32 * (uint64_t)b, acc);
}

void launch_msm_accumulate(kb_ctx* ctx, const uint32_t* tab, uint64_t tab_n, uint64_t first, const uint32_t* offsets,
                           const uint32_t* entries, uint32_t nb, uint32_t* buckets) {
  KB_LAUNCH(ctx, msm_accumulate_kernel, cdiv(nb, 256), 256, 0, tab, tab_n, first, offsets, entries, nb, buckets);
}

}  // namespace kb
