// Wire format of group elements: ark-serialize 0.4.2 `CanonicalSerialize` / `CanonicalDeserialize` for
// short-Weierstrass affine points (ark-ec 0.4.2 models/short_weierstrass/{mod.rs, serialization_flags.rs}) — the bytes
// a `Ciphertext<E>` (src/enc.rs:13: `(E::G2, Vec<u8>)`) or an opening proof travels in between the sender and the
// receiver of tests/laconic_ot.rs.  The reference never serialises them itself; these kernels make
// the transport real without a host round trip through arkworks.
//
//   compressed   = x             flags in the two top bits of the last byte        G1 32 B, G2 64 B
//   uncompressed = x || y        flags in the last byte of y                       G1 64 B, G2 128 B
//   flags: 0 = y <= -y, 1 << 7 = y > -y, 1 << 6 = infinity (x = y = 0)
//   Fq = 32 B little-endian canonical integer; Fq2 = c0 || c1, ordered by (c1, c0).
// Decoding takes square roots by exponentiation (q = 3 mod 4; Fq2 by the norm method); `validate` adds the curve
// equation and, for G2, the r-torsion check [r]P = O (`deserialize_*` against `deserialize_*_unchecked`).
#include "ctx.cuh"
#include "consts_gen.cuh"

namespace kb {

struct WireConsts { Fq2 twist_b; Fq two_inv; uint32_t sqrt_exp[8]; uint32_t half[8]; };
__constant__ WireConsts c_wire;

void wire_upload_consts() {
  WireConsts w;
  memcpy(&w.twist_b, consts::TWIST_B, 64);
  memcpy(&w.two_inv, consts::FQ_TWO_INV, 32);
  memcpy(w.sqrt_exp, consts::FQ_SQRT_EXP, 32);
  memcpy(w.half, consts::FQ_HALF, 32);
  KB_CUDA(cudaMemcpyToSymbol(c_wire, &w, sizeof(w)));
}

// a > b as 256-bit integers
__device__ __forceinline__ bool u256_gt(const uint32_t* a, const uint32_t* b) {
  uint32_t t = sub_cc(b[0], a[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) t = subc_cc(b[i], a[i]);
  (void)t;
  return subc(0, 0) != 0;   // b - a borrowed
}
// y > -y for a canonical y in [0, q): y > (q - 1) / 2
__device__ __forceinline__ bool fq_is_larger(const Fq& y_canon) { return u256_gt(y_canon.v, c_wire.half); }
__device__ __forceinline__ bool fq2_is_larger(const Fq2& y_canon) {
  return y_canon.c1.is_zero() ? fq_is_larger(y_canon.c0) : fq_is_larger(y_canon.c1);
}

__device__ __forceinline__ void put_fq(uint8_t* out, const Fq& canon, uint32_t flags) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t w = canon.v[i];
    if (i == 7) w |= flags << 24;
    out[4 * i] = (uint8_t)w; out[4 * i + 1] = (uint8_t)(w >> 8); out[4 * i + 2] = (uint8_t)(w >> 16); out[4 * i + 3] = (uint8_t)(w >> 24);
  }
}
// reads 32 bytes; `strip` removes (and returns through *flags) the two top bits of the last byte.
// ok = false if the integer is not below q (arkworks: InvalidData).  Result in Montgomery form.
__device__ __forceinline__ Fq get_fq(const uint8_t* in, bool strip, uint32_t* flags, bool* ok) {
  Fq c;
#pragma unroll
  for (int i = 0; i < 8; i++)
    c.v[i] = (uint32_t)in[4 * i] | ((uint32_t)in[4 * i + 1] << 8) | ((uint32_t)in[4 * i + 2] << 16) | ((uint32_t)in[4 * i + 3] << 24);
  if (strip) { *flags = c.v[7] >> 30; c.v[7] &= 0x3fffffffu; }
  uint32_t q[8];
#pragma unroll
  for (int i = 0; i < 8; i++) q[i] = FqParams::mod(i);
  if (!u256_gt(q, c.v)) *ok = false;
  return fp_to_mont<FqParams>(c);
}

// a^e for a 256-bit exponent, square-and-multiply MSB first
__device__ __noinline__ Fq fq_pow(Fq a, const uint32_t* e) {
  Fq r = Fq::one();
  bool started = false;
  for (int bit = 255; bit >= 0; bit--) {
    if (started) r = sqr(r);
    if ((e[bit >> 5] >> (bit & 31)) & 1u) { r = started ? r * a : a; started = true; }
  }
  return r;
}
// square root in Fq (q = 3 mod 4): a^((q+1)/4); *ok = false if a is not a square
__device__ __forceinline__ Fq fq_sqrt(const Fq& a, bool* ok) {
  Fq s = fq_pow(a, c_wire.sqrt_exp);
  if (sqr(s) != a) *ok = false;
  return s;
}
// square root in Fq2 = Fq[u] / (u^2 + 1) by the norm method
__device__ __noinline__ Fq2 fq2_sqrt(Fq2 a, bool* ok) {
  Fq2 r;
  if (a.c1.is_zero()) {
    bool sq = true;
    Fq s = fq_sqrt(a.c0, &sq);
    if (sq) { r.c0 = s; r.c1 = Fq::zero(); return r; }
    bool sq2 = true;                                   // -1 is a non-residue, so -a0 is a square
    r.c0 = Fq::zero(); r.c1 = fq_sqrt(-a.c0, &sq2);
    if (!sq2) *ok = false;
    return r;
  }
  bool sq = true;
  Fq alpha = fq_sqrt(sqr(a.c0) + sqr(a.c1), &sq);
  if (!sq) { *ok = false; return Fq2::zero(); }
  bool sd = true;
  Fq c0 = fq_sqrt((a.c0 + alpha) * c_wire.two_inv, &sd);
  if (!sd) { sd = true; c0 = fq_sqrt((a.c0 - alpha) * c_wire.two_inv, &sd); }
  if (!sd || c0.is_zero()) { *ok = false; return Fq2::zero(); }
  r.c0 = c0;
  r.c1 = a.c1 * inv(dbl(c0));
  if (sqr(r) != a) *ok = false;
  return r;
}

// ------------------------------------------------------------------------------------------
// G1
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) g1_serialize_kernel(const uint32_t* __restrict__ xy, const uint8_t* __restrict__ inf, uint64_t n,
                                                           int compress, uint8_t* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t len = compress ? 32 : 64;
  uint8_t* o = out + len * i;
  G1Affine p = ld_g1(xy + 16 * i);
  const bool is_inf = (inf && inf[i]) || p.is_inf();
  Fq x = is_inf ? Fq::zero() : fp_from_mont<FqParams>(p.x), y = is_inf ? Fq::zero() : fp_from_mont<FqParams>(p.y);
  const uint32_t flags = is_inf ? 0x40u : (fq_is_larger(y) ? 0x80u : 0u);
  if (compress) put_fq(o, x, flags);
  else { put_fq(o, x, 0); put_fq(o + 32, y, flags); }
}

__global__ void __launch_bounds__(128) g1_deserialize_kernel(const uint8_t* __restrict__ in, uint64_t n, int compress, int validate,
                                                             uint32_t* __restrict__ xy, uint8_t* __restrict__ inf, uint8_t* __restrict__ okv) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t len = compress ? 32 : 64;
  const uint8_t* b = in + len * i;
  bool ok = true;
  uint32_t flags = 0;
  G1Affine p;
  if (compress) { p.x = get_fq(b, true, &flags, &ok); p.y = Fq::zero(); }
  else { p.x = get_fq(b, false, nullptr, &ok); p.y = get_fq(b + 32, true, &flags, &ok); }
  if (flags == 3u) ok = false;
  bool is_inf = (flags & 1u) != 0;   // bit 6 of the byte = bit 0 of the two stripped bits
  if (ok && !is_inf) {
    Fq three = Fq::one() + Fq::one() + Fq::one();
    Fq rhs = sqr(p.x) * p.x + three;
    if (compress) {
      Fq y = fq_sqrt(rhs, &ok);
      if (ok && fq_is_larger(fp_from_mont<FqParams>(y)) != ((flags & 2u) != 0)) y = -y;
      p.y = y;
    } else if (validate && sqr(p.y) != rhs) ok = false;
  }
  if (!ok || is_inf) p = G1Affine::infinity();
  st_g1(xy + 16 * i, p);
  inf[i] = (!ok || is_inf) ? 1 : 0;
  okv[i] = ok ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// G2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) g2_serialize_kernel(const uint32_t* __restrict__ xy, const uint8_t* __restrict__ inf, uint64_t n,
                                                           int compress, uint8_t* __restrict__ out) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t len = compress ? 64 : 128;
  uint8_t* o = out + len * i;
  G2Affine p = ld_g2(xy + 32 * i);
  const bool is_inf = (inf && inf[i]) || p.is_inf();
  Fq2 x, y;
  x.c0 = is_inf ? Fq::zero() : fp_from_mont<FqParams>(p.x.c0); x.c1 = is_inf ? Fq::zero() : fp_from_mont<FqParams>(p.x.c1);
  y.c0 = is_inf ? Fq::zero() : fp_from_mont<FqParams>(p.y.c0); y.c1 = is_inf ? Fq::zero() : fp_from_mont<FqParams>(p.y.c1);
  const uint32_t flags = is_inf ? 0x40u : (fq2_is_larger(y) ? 0x80u : 0u);
  put_fq(o, x.c0, 0);
  if (compress) put_fq(o + 32, x.c1, flags);
  else { put_fq(o + 32, x.c1, 0); put_fq(o + 64, y.c0, 0); put_fq(o + 96, y.c1, flags); }
}

__global__ void __launch_bounds__(128) g2_deserialize_kernel(const uint8_t* __restrict__ in, uint64_t n, int compress, int validate,
                                                             uint32_t* __restrict__ xy, uint8_t* __restrict__ inf, uint8_t* __restrict__ okv) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t len = compress ? 64 : 128;
  const uint8_t* b = in + len * i;
  bool ok = true;
  uint32_t flags = 0;
  G2Affine p;
  p.x.c0 = get_fq(b, false, nullptr, &ok);
  if (compress) { p.x.c1 = get_fq(b + 32, true, &flags, &ok); p.y = Fq2::zero(); }
  else { p.x.c1 = get_fq(b + 32, false, nullptr, &ok); p.y.c0 = get_fq(b + 64, false, nullptr, &ok); p.y.c1 = get_fq(b + 96, true, &flags, &ok); }
  if (flags == 3u) ok = false;
  bool is_inf = (flags & 1u) != 0;
  if (ok && !is_inf) {
    Fq2 rhs = sqr(p.x) * p.x + c_wire.twist_b;
    if (compress) {
      Fq2 y = fq2_sqrt(rhs, &ok);
      if (ok) {
        Fq2 yc; yc.c0 = fp_from_mont<FqParams>(y.c0); yc.c1 = fp_from_mont<FqParams>(y.c1);
        if (fq2_is_larger(yc) != ((flags & 2u) != 0)) y = -y;
      }
      p.y = y;
    } else if (validate && sqr(p.y) != rhs) ok = false;
    if (ok && validate) {   // [r] P == O
      uint32_t r[8];
#pragma unroll
      for (int k = 0; k < 8; k++) r[k] = FrParams::mod(k);
      if (!ec_mul(to_xyzz(p), r).is_inf()) ok = false;
    }
  }
  if (!ok || is_inf) p = G2Affine::infinity();
  st_g2(xy + 32 * i, p);
  inf[i] = (!ok || is_inf) ? 1 : 0;
  okv[i] = ok ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// SRS validation: the reference reads the ptau sections with `deserialize_uncompressed_unchecked`
// (src/kzg/ptau.rs:266,314) and never tests the points; here every resident G1 power must be a canonical-range,
// finite point of y^2 = x^3 + 3 (cofactor 1: on the curve = in the group) and [tau]_2 a point of the twist in the
// r-torsion.  first_bad = smallest failing G1 index, or n when only [tau]_2 fails.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool fq_in_range(const Fq& a) {
  uint32_t q[8];
#pragma unroll
  for (int i = 0; i < 8; i++) q[i] = FqParams::mod(i);
  return u256_gt(q, a.v);
}
__global__ void __launch_bounds__(256) srs_validate_g1_kernel(const uint32_t* __restrict__ srs, uint64_t n, unsigned long long* __restrict__ first_bad) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const G1Affine p = ld_g1(srs + 16 * i);
  const Fq three = Fq::one() + Fq::one() + Fq::one();
  const bool ok = fq_in_range(p.x) && fq_in_range(p.y) && !p.is_inf() && sqr(p.y) == sqr(p.x) * p.x + three;
  if (!ok) atomicMin(first_bad, (unsigned long long)i);
}
__global__ void srs_validate_tau2_kernel(const uint32_t* __restrict__ tau2, uint64_t n, unsigned long long* __restrict__ first_bad) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const G2Affine p = ld_g2(tau2);
  bool ok = fq_in_range(p.x.c0) && fq_in_range(p.x.c1) && fq_in_range(p.y.c0) && fq_in_range(p.y.c1) && !p.is_inf();
  if (ok) ok = sqr(p.y) == sqr(p.x) * p.x + c_wire.twist_b;
  if (ok) {
    uint32_t r[8];
#pragma unroll
    for (int k = 0; k < 8; k++) r[k] = FrParams::mod(k);
    ok = ec_mul(to_xyzz(p), r).is_inf();
  }
  if (!ok) atomicMin(first_bad, (unsigned long long)n);
}
// *d_first_bad = ~0 if the SRS is valid
void srs_validate(kb_ctx* ctx, unsigned long long* d_first_bad) {
  KB_CUDA(cudaMemsetAsync(d_first_bad, 0xff, sizeof(unsigned long long), ctx->stream));
  if (ctx->srs_n) KB_LAUNCH(ctx, srs_validate_g1_kernel, cdiv(ctx->srs_n, 256), 256, 0, ctx->d_srs, ctx->srs_n, d_first_bad);
  KB_LAUNCH(ctx, srs_validate_tau2_kernel, 1, 32, 0, ctx->d_tau2_tab, ctx->srs_n, d_first_bad);   // entry (window 0, digit 1) = tau_2
}

void g1_serialize(kb_ctx* ctx, const uint32_t* d_xy, const uint8_t* d_inf, uint64_t n, int compress, uint8_t* d_out) {
  if (n) KB_LAUNCH(ctx, g1_serialize_kernel, cdiv(n, 128), 128, 0, d_xy, d_inf, n, compress, d_out);
}
void g2_serialize(kb_ctx* ctx, const uint32_t* d_xy, const uint8_t* d_inf, uint64_t n, int compress, uint8_t* d_out) {
  if (n) KB_LAUNCH(ctx, g2_serialize_kernel, cdiv(n, 128), 128, 0, d_xy, d_inf, n, compress, d_out);
}
void g1_deserialize(kb_ctx* ctx, const uint8_t* d_in, uint64_t n, int compress, int validate, uint32_t* d_xy, uint8_t* d_inf, uint8_t* d_ok) {
  if (n) KB_LAUNCH(ctx, g1_deserialize_kernel, cdiv(n, 128), 128, 0, d_in, n, compress, validate, d_xy, d_inf, d_ok);
}
void g2_deserialize(kb_ctx* ctx, const uint8_t* d_in, uint64_t n, int compress, int validate, uint32_t* d_xy, uint8_t* d_inf, uint8_t* d_ok) {
  if (n) KB_LAUNCH(ctx, g2_deserialize_kernel, cdiv(n, 128), 128, 0, d_in, n, compress, validate, d_xy, d_inf, d_ok);
}

}  // namespace kb
