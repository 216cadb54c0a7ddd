// Signed-digit window recoding for the bucketed Pippenger MSM.
// A canonical scalar k < r < 2^254 is written as sum_w d_w 2^(c w) with d_w in (-2^(c-1), 2^(c-1)],
// so a window needs only 2^(c-1) buckets (negative digits add the negated base).  With
// nwin = ceil(255 / c) the top window never produces a carry.
#pragma once
#include "fp.cuh"

namespace kb {

KB_HD int msm_num_windows(int c) { return (255 + c - 1) / c; }

// c-bit field starting at bit `bit` of a 256-bit little-endian integer (zero-extended past bit 255)
KB_HD uint32_t msm_window_bits(const uint32_t* k, int bit, int c) {
  int limb = bit >> 5, sh = bit & 31;
  if (limb >= 8) return 0;
  uint64_t v = k[limb];
  if (limb + 1 < 8) v |= (uint64_t)k[limb + 1] << 32;
  return (uint32_t)((v >> sh) & ((1u << c) - 1u));
}

KB_HD void msm_signed_digits(const uint32_t* k, int c, int nwin, int32_t* digits) {
  uint32_t carry = 0;
  const uint32_t half = 1u << (c - 1);
  for (int w = 0; w < nwin; w++) {
    uint32_t raw = msm_window_bits(k, w * c, c) + carry;
    if (raw > half) { digits[w] = (int32_t)raw - (int32_t)(1u << c); carry = 1; }
    else { digits[w] = (int32_t)raw; carry = 0; }
  }
}

}  // namespace kb
