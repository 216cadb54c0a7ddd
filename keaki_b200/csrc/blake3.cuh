// BLAKE3 (default hash mode, XOF output) specialised to the one shape the path needs: a single
// 384-byte input (the ark-serialize image of a GT element) hashed to `out_len` bytes.
// Replaces `blake3::Hasher::new().update(..).finalize_xof().fill(..)` (blake3 1.5.4, Cargo.lock:165-166)
// at src/kem.rs:42-46,65-69.  384 B = 6 blocks of 64 B in one chunk: five chained compressions, then
// the last block is compressed once per 64 B of output with the ROOT flag and an output-block counter.
#pragma once
#include "fp.cuh"

namespace kb {

KB_HD uint32_t b3_iv(int i) {
  constexpr uint32_t iv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  return iv[i];
}
enum { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_ROOT = 8 };

KB_HD uint32_t b3_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

#define KB_B3_G(a, b, c, d, mx, my) \
  do { a = a + b + (mx); d = b3_rotr(d ^ a, 16); c = c + d; b = b3_rotr(b ^ c, 12); \
       a = a + b + (my); d = b3_rotr(d ^ a, 8);  c = c + d; b = b3_rotr(b ^ c, 7); } while (0)

// Full 16-word compression output.
KB_HD void b3_compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter, uint32_t block_len,
                       uint32_t flags, uint32_t out[16]) {
  uint32_t s[16], m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) s[i] = cv[i];
#pragma unroll
  for (int i = 0; i < 4; i++) s[8 + i] = b3_iv(i);
  s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = block_len; s[15] = flags;
#pragma unroll
  for (int i = 0; i < 16; i++) m[i] = block[i];
#pragma unroll
  for (int r = 0; r < 7; r++) {
    KB_B3_G(s[0], s[4], s[8], s[12], m[0], m[1]);
    KB_B3_G(s[1], s[5], s[9], s[13], m[2], m[3]);
    KB_B3_G(s[2], s[6], s[10], s[14], m[4], m[5]);
    KB_B3_G(s[3], s[7], s[11], s[15], m[6], m[7]);
    KB_B3_G(s[0], s[5], s[10], s[15], m[8], m[9]);
    KB_B3_G(s[1], s[6], s[11], s[12], m[10], m[11]);
    KB_B3_G(s[2], s[7], s[8], s[13], m[12], m[13]);
    KB_B3_G(s[3], s[4], s[9], s[14], m[14], m[15]);
    if (r < 6) {
      // message permutation 2 6 3 10 7 0 4 13 1 11 12 5 9 14 15 8
      uint32_t t[16];
      t[0] = m[2]; t[1] = m[6]; t[2] = m[3]; t[3] = m[10]; t[4] = m[7]; t[5] = m[0]; t[6] = m[4]; t[7] = m[13];
      t[8] = m[1]; t[9] = m[11]; t[10] = m[12]; t[11] = m[5]; t[12] = m[9]; t[13] = m[14]; t[14] = m[15]; t[15] = m[8];
#pragma unroll
      for (int i = 0; i < 16; i++) m[i] = t[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) { out[i] = s[i] ^ s[i + 8]; out[i + 8] = s[i + 8] ^ cv[i]; }
}

// key[j] (^= msg[j] if msg != nullptr) for j < out_len, from the 96-word GT image.
// Bytes are little-endian within words, as in the BLAKE3 specification.
KB_HD void b3_gt_xof_xor(const uint32_t gt_words[96], const uint8_t* msg, uint8_t* out, uint64_t out_len) {
  uint32_t cv[8], o[16];
  for (int i = 0; i < 8; i++) cv[i] = b3_iv(i);
  for (int b = 0; b < 5; b++) {
    b3_compress(cv, gt_words + 16 * b, 0, 64, b == 0 ? B3_CHUNK_START : 0, o);
    for (int i = 0; i < 8; i++) cv[i] = o[i];
  }
  for (uint64_t off = 0, ctr = 0; off < out_len; off += 64, ctr++) {
    b3_compress(cv, gt_words + 80, ctr, 64, B3_CHUNK_END | B3_ROOT, o);
    uint64_t n = out_len - off < 64 ? out_len - off : 64;
    for (uint64_t j = 0; j < n; j++) {
      uint8_t k = (uint8_t)(o[j >> 2] >> (8 * (j & 3)));
      out[off + j] = msg ? (uint8_t)(k ^ msg[off + j]) : k;
    }
  }
}

}  // namespace kb
