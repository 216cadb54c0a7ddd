// keaki_b200.hpp — C++ host side above the C ABI (keaki_b200.h): the keaki API with E = Bn254.
//
// The reference is a Rust crate and there is no Rust toolchain in this image, so the host layer a
// keaki user programs against is restated here in C++ with the reference's names, argument order
// (leading `rng`), return shapes and error behaviour:
//
//   KZGSetup::new_from_file / setup / g1_pow / g1_aff / tau_g2        src/kzg.rs:22-85
//   commit / open / verify / open_fk, KZGError::PolynomialTooLarge    src/kzg.rs:89-209
//   encapsulate / decapsulate                                         src/kem.rs:13-72
//   Ciphertext, encrypt / decrypt                                     src/enc.rs:13-55
//   PADDING_LEN, vec_commit / vec_encrypt / vec_decrypt               src/vec.rs:18-81
//   Receiver / Sender (laconic OT)                                    tests/laconic_ot.rs:15-113
//   Radix2EvaluationDomain::{new, size, elements, fft, ifft}          ark-poly, used at src/vec.rs:36-37, src/kzg.rs:163
//
// It holds NO curve or pairing arithmetic: every group / pairing / transform operation is one call
// into libkeaki_b200.so (CUDA, sm_100a; there is no CPU fallback — creating a setup without a GPU
// throws).  What does live here is what lives on the host in the reference too: the scalar field
// (ark-ff's `Fr`, needed to build polynomials, evaluate them and draw randomness) and container I/O.
// `Fr` keeps arkworks' in-RAM form (4 x u64 Montgomery limbs, R = 2^256), so values cross the ABI
// without conversion.  Header-only, C++17, links against -lkeaki_b200.
#ifndef KEAKI_B200_HPP
#define KEAKI_B200_HPP

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "keaki_b200.h"

namespace keaki {

// ---------------------------------------------------------------------------------------------
// Fr: the BN254 scalar field, Montgomery form (ark-bn254 `Fr` = Fp256<MontBackend<FrConfig, 4>>)
// ---------------------------------------------------------------------------------------------
struct Fr {
  uint64_t l[4];   // Montgomery limbs, little-endian: exactly `fe.0.0` of arkworks

  static constexpr uint64_t MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t INV = 0xc2e1f593efffffffull;   // -r^-1 mod 2^64
  static constexpr uint64_t R1[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};   // 2^256 mod r
  static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};   // 2^512 mod r

  static Fr zero() { return Fr{{0, 0, 0, 0}}; }
  static Fr one() { return Fr{{R1[0], R1[1], R1[2], R1[3]}}; }
  static Fr from_u64(uint64_t x) { Fr a{{x, 0, 0, 0}}; return mont_mul(a, Fr{{R2[0], R2[1], R2[2], R2[3]}}); }
  static Fr from_i64(int64_t x) { return x >= 0 ? from_u64((uint64_t)x) : -from_u64((uint64_t)(-x)); }
  /// canonical little-endian integer (4 limbs) -> field element; the integer must be below r
  static Fr from_canonical(const uint64_t c[4]) { Fr a{{c[0], c[1], c[2], c[3]}}; return mont_mul(a, Fr{{R2[0], R2[1], R2[2], R2[3]}}); }
  void to_canonical(uint64_t c[4]) const { Fr o = mont_mul(*this, Fr{{1, 0, 0, 0}}); std::memcpy(c, o.l, 32); }

  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool operator==(const Fr& b) const { return l[0] == b.l[0] && l[1] == b.l[1] && l[2] == b.l[2] && l[3] == b.l[3]; }
  bool operator!=(const Fr& b) const { return !(*this == b); }

  Fr operator+(const Fr& b) const {
    Fr r; unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) { c += (unsigned __int128)l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  Fr operator-(const Fr& b) const {
    Fr r; unsigned __int128 br = 0;
    for (int i = 0; i < 4; i++) { unsigned __int128 d = (unsigned __int128)l[i] - b.l[i] - (uint64_t)br; r.l[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { unsigned __int128 c = 0; for (int i = 0; i < 4; i++) { c += (unsigned __int128)r.l[i] + MOD[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
    return r;
  }
  Fr operator-() const { return zero() - *this; }
  Fr operator*(const Fr& b) const { return mont_mul(*this, b); }
  Fr& operator+=(const Fr& b) { return *this = *this + b; }
  Fr& operator*=(const Fr& b) { return *this = *this * b; }
  Fr pow(uint64_t e) const { Fr acc = one(), base = *this; for (; e; e >>= 1) { if (e & 1) acc = acc * base; base = base * base; } return acc; }
  Fr inverse() const {   // a^(r-2); zero maps to zero (callers check, as `inverse().unwrap()` would panic)
    uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
    Fr acc = one(), base = *this;
    for (int i = 0; i < 256; i++) { if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * base; base = base * base; }
    return acc;
  }

  /// ark-ff `Fr::rand`: four u64 from the generator, the two top bits cleared, rejected if not below r, and the bits
  /// are used AS the Montgomery representation (src/kem.rs:26, src/vec.rs:32 draw through this).
  template <class Rng> static Fr rand(Rng& rng) {
    for (;;) {
      Fr a;
      for (int i = 0; i < 4; i++) a.l[i] = rng.next_u64();
      a.l[3] &= 0x3fffffffffffffffull;
      if (!geq_mod(a.l)) return a;
    }
  }

 private:
  static bool geq_mod(const uint64_t* a) { for (int i = 3; i >= 0; i--) { if (a[i] > MOD[i]) return true; if (a[i] < MOD[i]) return false; } return true; }
  static void sub_mod(uint64_t* a) { unsigned __int128 b = 0; for (int i = 0; i < 4; i++) { unsigned __int128 t = (unsigned __int128)a[i] - MOD[i] - (uint64_t)b; a[i] = (uint64_t)t; b = (t >> 64) & 1; } }
  static Fr mont_mul(const Fr& a, const Fr& b) {   // CIOS
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      unsigned __int128 c = 0;
      for (int j = 0; j < 4; j++) { c += (unsigned __int128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * INV;
      c = ((unsigned __int128)m * MOD[0] + t[0]) >> 64;
      for (int j = 1; j < 4; j++) { c += (unsigned __int128)m * MOD[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq_mod(t)) sub_mod(t);
    Fr r; std::memcpy(r.l, t, 32); return r;
  }
};

/// Deterministic generator for tests (the reference's tests use `ark_std::test_rng()`); anything with `next_u64()` works.
struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next_u64() { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
  void fill_bytes(uint8_t* p, size_t n) { for (size_t i = 0; i < n; i++) p[i] = (uint8_t)next_u64(); }
};

// ---------------------------------------------------------------------------------------------
// Group elements: opaque affine Montgomery coordinates + infinity flag, as they cross the C ABI
// ---------------------------------------------------------------------------------------------
struct G1 {
  uint32_t xy[16] = {0};
  bool inf = true;
  bool operator==(const G1& b) const { return inf == b.inf && (inf || std::memcmp(xy, b.xy, 64) == 0); }
  bool operator!=(const G1& b) const { return !(*this == b); }
  static G1 zero() { return G1(); }
};
struct G2 {
  uint32_t xy[32] = {0};
  bool inf = true;
  bool operator==(const G2& b) const { return inf == b.inf && (inf || std::memcmp(xy, b.xy, 128) == 0); }
  bool operator!=(const G2& b) const { return !(*this == b); }
};
using Bytes = std::vector<uint8_t>;

// ---------------------------------------------------------------------------------------------
// Errors
// ---------------------------------------------------------------------------------------------
/// `KZGError::PolynomialTooLarge(usize, usize)` — the only recoverable error of the reference (src/kzg.rs:205-209)
struct KZGError : std::runtime_error {
  size_t len, max;
  KZGError(size_t l, size_t m) : std::runtime_error("PolynomialTooLarge(" + std::to_string(l) + ", " + std::to_string(m) + ")"), len(l), max(m) {}
};
/// `SetupFileError` (src/kzg/ptau.rs:360-376)
struct SetupFileError : std::runtime_error { using std::runtime_error::runtime_error; };
/// Any other non-zero status of the library (CUDA failure, bad argument): the reference panics in these places.
struct BackendError : std::runtime_error { using std::runtime_error::runtime_error; };

namespace detail {
struct CtxDeleter { void operator()(kb_ctx* c) const { if (c) kb_ctx_destroy(c); } };
using CtxPtr = std::shared_ptr<kb_ctx>;
inline CtxPtr make_ctx(int device) {
  kb_ctx* c = nullptr;
  int32_t rc = kb_ctx_create(device, &c);
  if (rc != KB_OK || !c) throw BackendError("keaki_b200: no usable sm_100 CUDA device (kb_ctx_create = " + std::to_string(rc) + "); there is no CPU fallback");
  return CtxPtr(c, CtxDeleter());
}
inline void check(kb_ctx* c, int32_t rc, const char* what) {
  if (rc == KB_OK) return;
  const char* m = kb_last_error(c);
  throw BackendError(std::string(what) + ": " + (m ? m : "error") + " (" + std::to_string(rc) + ")");
}
/// context for the calls of the reference that take no setup (`decapsulate`, `decrypt`, `vec_decrypt`, domains): the
/// context of the most recently created KZGSetup (same device, no second set of tables); one on device 0 is created
/// only if no setup exists yet
inline std::weak_ptr<kb_ctx>& latest_setup_ctx() { static std::weak_ptr<kb_ctx> w; return w; }
inline CtxPtr& default_ctx_slot() { static CtxPtr c; return c; }
inline CtxPtr default_ctx_ptr() {
  if (CtxPtr live = latest_setup_ctx().lock()) return live;
  if (!default_ctx_slot()) default_ctx_slot() = make_ctx(0);
  return default_ctx_slot();
}
inline kb_ctx* default_ctx() { return default_ctx_ptr().get(); }   // callers use it within one blocking call
inline const uint32_t* u32(const Fr* p) { return reinterpret_cast<const uint32_t*>(p); }
inline uint32_t* u32(Fr* p) { return reinterpret_cast<uint32_t*>(p); }
static_assert(sizeof(Fr) == 32, "Fr must be 4 x u64");
}  // namespace detail

// ---------------------------------------------------------------------------------------------
// Radix2EvaluationDomain (ark-poly): size = next power of two, generator 5^((r-1)/size)
// ---------------------------------------------------------------------------------------------
class Radix2EvaluationDomain {
 public:
  /// `Radix2EvaluationDomain::new(n)`: None (here: nullptr-like `valid() == false`) when the size exceeds 2^28
  explicit Radix2EvaluationDomain(size_t n) { size_ = 1; while (size_ < n) size_ <<= 1; valid_ = size_ <= (size_t(1) << 28); }
  bool valid() const { return valid_; }
  size_t size() const { return size_; }
  /// 1, w, w^2, ...: the forward transform of the unit vector e_1 (computed by the library's NTT)
  std::vector<Fr> elements() const { std::vector<Fr> e(size_, Fr::zero()); if (size_ > 1) e[1] = Fr::one(); else e[0] = Fr::one(); if (size_ > 1) ntt(e, false); return e; }
  std::vector<Fr> fft(std::vector<Fr> coeffs) const { coeffs.resize(size_, Fr::zero()); ntt(coeffs, false); return coeffs; }
  std::vector<Fr> ifft(std::vector<Fr> evals) const { evals.resize(size_, Fr::zero()); ntt(evals, true); return evals; }

 private:
  void ntt(std::vector<Fr>& v, bool inverse) const {
    kb_ctx* c = detail::default_ctx();
    detail::check(c, kb_fr_ntt(c, detail::u32(v.data()), v.size(), inverse ? 1 : 0), "kb_fr_ntt");
  }
  size_t size_;
  bool valid_;
};

// ---------------------------------------------------------------------------------------------
// ptau container (src/kzg/ptau.rs): header + TauG1 + TauG2 sections.  Coordinates are snarkjs Montgomery limbs and
// are passed through as such (DESIGN.md "Deliberate deviation").
// ---------------------------------------------------------------------------------------------
namespace ptau {
inline void get_powers_from_file(const std::string& path, std::vector<G1>& g1, std::vector<G2>& g2) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw SetupFileError("FileError(" + path + ")");
  std::vector<uint8_t> data;
  uint8_t buf[1 << 16];
  size_t k;
  while ((k = std::fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + k);
  std::fclose(f);
  auto rd32 = [&](size_t o) { uint32_t v; std::memcpy(&v, &data[o], 4); return v; };
  auto rd64 = [&](size_t o) { uint64_t v; std::memcpy(&v, &data[o], 8); return v; };
  if (data.size() < 12 || std::memcmp(data.data(), "ptau", 4) != 0) throw SetupFileError("InvalidFileType");
  if (rd32(8) != 11) throw SetupFileError("InvalidNumberOfSections(" + std::to_string(rd32(8)) + ")");
  size_t off = 12, sec_off[16] = {0}, sec_len[16] = {0};
  bool sec_seen[16] = {false};
  for (int s = 0; s < 11; s++) {
    if (off + 12 > data.size()) throw SetupFileError("UnexpectedEof");
    uint32_t id = rd32(off); uint64_t len = rd64(off + 4);
    if (!((id >= 1 && id <= 7) || (id >= 12 && id <= 15))) throw SetupFileError("UnknownSection(" + std::to_string(id) + ")");
    off += 12;
    if (len > data.size() - off) throw SetupFileError("UnexpectedEof");
    if (!sec_seen[id]) { sec_seen[id] = true; sec_off[id] = off; sec_len[id] = (size_t)len; }   // a repeated id keeps its first occurrence
    off += (size_t)len;
  }
  if (off != data.size()) throw SetupFileError("SectionsNotContiguous");
  // `FileSections::get` (src/kzg/ptau.rs:146-150): the three sections that are read must be present
  for (int id = 1; id <= 3; id++) if (!sec_seen[id]) throw SetupFileError("EmptySection(" + std::to_string(id) + ")");
  size_t h = sec_off[1];
  static const uint8_t QMOD[32] = {0x47, 0xfd, 0x7c, 0xd8, 0x16, 0x8c, 0x20, 0x3c, 0x8d, 0xca, 0x71, 0x68, 0x91, 0x6a, 0x81, 0x97,
                                   0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};
  if (sec_len[1] < 44) throw SetupFileError("ParseError(header section too short)");
  if (rd32(h) != 32 || std::memcmp(&data[h + 4], QMOD, 32) != 0) throw SetupFileError("InvalidFieldModulus");
  uint32_t power = rd32(h + 36);
  if (power > 28) throw SetupFileError("ParseError(power " + std::to_string(power) + " > 28)");   // 2-adicity of Fr; keeps the sizes below in range
  size_t n1 = 2 * (size_t(1) << power) - 1, n2 = size_t(1) << power;
  // src/kzg/ptau.rs:251-256, 299-304: exact section sizes
  if (sec_len[2] != n1 * 64) throw SetupFileError("ElementSizeMismatch(" + std::to_string(n1 * 64) + ", " + std::to_string(sec_len[2]) + ")");
  if (sec_len[3] != n2 * 128) throw SetupFileError("ElementSizeMismatch(" + std::to_string(n2 * 128) + ", " + std::to_string(sec_len[3]) + ")");
  g1.resize(n1); g2.resize(n2);
  for (size_t i = 0; i < n1; i++) { std::memcpy(g1[i].xy, &data[sec_off[2] + 64 * i], 64); g1[i].inf = false; }
  for (size_t i = 0; i < n2; i++) { std::memcpy(g2[i].xy, &data[sec_off[3] + 128 * i], 128); g2[i].inf = false; }
}
}  // namespace ptau

// ---------------------------------------------------------------------------------------------
// KZG (src/kzg.rs)
// ---------------------------------------------------------------------------------------------
class KZGSetup {
 public:
  /// src/kzg.rs:33-52
  static KZGSetup new_from_file(const std::string& file, int device = 0, bool validate = true) {
    std::vector<G1> g1; std::vector<G2> g2;
    ptau::get_powers_from_file(file, g1, g2);
    if (g2.size() < 2) throw SetupFileError("EmptySection(3)");
    KZGSetup s;
    s.ctx_ = detail::make_ctx(device);
    std::vector<uint32_t> flat(16 * g1.size());
    for (size_t i = 0; i < g1.size(); i++) std::memcpy(&flat[16 * i], g1[i].xy, 64);
    detail::check(s.ctx_.get(), kb_srs_upload(s.ctx_.get(), flat.data(), g1.size(), g2[1].xy), "kb_srs_upload");
    // unlike the reference (`deserialize_uncompressed_unchecked`, src/kzg/ptau.rs:266,314) the points are validated (on the GPU)
    if (validate) {
      uint64_t bad = 0;
      int32_t rc = kb_srs_validate(s.ctx_.get(), &bad);
      if (rc == KB_ERR_INVALID_POINT) throw SetupFileError(std::string("ParseError(") + kb_last_error(s.ctx_.get()) + ")");
      detail::check(s.ctx_.get(), rc, "kb_srs_validate");
    }
    s.g1_ = std::move(g1); s.tau_g2_ = g2[1];
    detail::latest_setup_ctx() = s.ctx_;
    return s;
  }
  /// src/kzg.rs:55-70 ("Don't use this"): g1_pow[i] = tau^i G1, tau_g2 = tau G2, generated on the device
  static KZGSetup setup(const Fr& secret, size_t max_d, int device = 0) {
    KZGSetup s;
    s.ctx_ = detail::make_ctx(device);
    std::vector<uint32_t> flat(16 * (max_d ? max_d : 1));
    s.tau_g2_.inf = false;
    detail::check(s.ctx_.get(), kb_srs_generate(s.ctx_.get(), detail::u32(&secret), 0, max_d, flat.data(), s.tau_g2_.xy), "kb_srs_generate");
    s.g1_.resize(max_d);
    for (size_t i = 0; i < max_d; i++) { std::memcpy(s.g1_[i].xy, &flat[16 * i], 64); s.g1_[i].inf = false; }
    detail::latest_setup_ctx() = s.ctx_;
    return s;
  }
  const std::vector<G1>& g1_pow() const { return g1_; }
  const std::vector<G1>& g1_aff() const { return g1_; }
  const G2& tau_g2() const { return tau_g2_; }
  kb_ctx* ctx() const { return ctx_.get(); }

 private:
  detail::CtxPtr ctx_;
  std::vector<G1> g1_;
  G2 tau_g2_;
};

/// `DensePolynomial::from_coefficients_vec` strips trailing zeros; the functions below take coefficient vectors.
inline std::vector<Fr> dense_polynomial(std::vector<Fr> c) { while (!c.empty() && c.back().is_zero()) c.pop_back(); return c; }
inline Fr evaluate(const std::vector<Fr>& p, const Fr& x) { Fr acc = Fr::zero(); for (size_t i = p.size(); i-- > 0;) acc = acc * x + p[i]; return acc; }

/// src/kzg.rs:89-101
inline G1 commit(const KZGSetup& setup, const std::vector<Fr>& p_in) {
  std::vector<Fr> p = dense_polynomial(p_in);
  if (p.size() > setup.g1_pow().size()) throw KZGError(p.size(), setup.g1_pow().size());
  G1 out; uint8_t inf = 1;
  detail::check(setup.ctx(), kb_msm_g1(setup.ctx(), detail::u32(p.data()), 0, p.size(), out.xy, &inf), "kb_msm_g1");
  out.inf = inf != 0;
  return out;
}
/// src/kzg.rs:104-124
inline G1 open(const KZGSetup& setup, const std::vector<Fr>& p_in, const Fr& point) {
  std::vector<Fr> p = dense_polynomial(p_in);
  if (p.size() >= 1 && p.size() - 1 > setup.g1_pow().size()) throw KZGError(p.size() - 1, setup.g1_pow().size());
  G1 out; uint8_t inf = 1;
  detail::check(setup.ctx(), kb_open_batch(setup.ctx(), detail::u32(p.data()), p.size(), detail::u32(&point), 1, out.xy, &inf), "kb_open_batch");
  out.inf = inf != 0;
  return out;
}
/// src/kzg.rs:127-151
inline bool verify(const KZGSetup& setup, const G1& commitment, const Fr& point, const Fr& value, const G1& proof) {
  uint8_t ci = commitment.inf, pi = proof.inf, ok = 0;
  detail::check(setup.ctx(), kb_verify_batch(setup.ctx(), commitment.xy, &ci, detail::u32(&point), detail::u32(&value), proof.xy, &pi, 1, &ok), "kb_verify_batch");
  return ok != 0;
}
/// src/kzg.rs:157-203: all d openings at the d-th roots of unity; `p` has exactly d = domain.size() coefficients.
/// The reference panics when d exceeds the SRS (slice at :169) or 2d > 2^28 (unwrap at :163): std::out_of_range here.
inline std::vector<G1> open_fk(const KZGSetup& setup, const std::vector<Fr>& p, const Radix2EvaluationDomain& domain_d) {
  const size_t d = domain_d.size();
  if (p.size() != d) throw std::invalid_argument("open_fk: p must have domain.size() coefficients");
  if (d > setup.g1_pow().size() || 2 * d > (size_t(1) << 28)) throw std::out_of_range("open_fk: d exceeds the SRS / the 2-adicity of Fr");
  std::vector<uint32_t> xy(16 * d); std::vector<uint8_t> inf(d);
  detail::check(setup.ctx(), kb_open_all_fk(setup.ctx(), detail::u32(p.data()), d, xy.data(), inf.data()), "kb_open_all_fk");
  std::vector<G1> out(d);
  for (size_t i = 0; i < d; i++) { std::memcpy(out[i].xy, &xy[16 * i], 64); out[i].inf = inf[i] != 0; }
  return out;
}

// ---------------------------------------------------------------------------------------------
// KEM (src/kem.rs) and encryption (src/enc.rs)
// ---------------------------------------------------------------------------------------------
using Ciphertext = std::pair<G2, Bytes>;   // src/enc.rs:13

namespace detail {
inline void encrypt_batch(const KZGSetup& s, const G1& com, const Fr* points, const Fr* values, const Fr* rs, const std::vector<Bytes>& msgs,
                          std::vector<Ciphertext>& out) {
  const size_t n = msgs.size();
  std::vector<uint64_t> off(n + 1, 0);
  for (size_t i = 0; i < n; i++) off[i + 1] = off[i] + msgs[i].size();
  std::vector<uint8_t> flat(off[n] ? off[n] : 1), ct_msg(off[n] ? off[n] : 1), ct_inf(n ? n : 1);
  for (size_t i = 0; i < n; i++) if (!msgs[i].empty()) std::memcpy(&flat[off[i]], msgs[i].data(), msgs[i].size());
  std::vector<uint32_t> ct(32 * (n ? n : 1));
  check(s.ctx(), kb_encrypt_batch(s.ctx(), com.xy, com.inf ? 1 : 0, u32(points), u32(values), u32(rs), flat.data(), off.data(), n,
                                  ct.data(), ct_inf.data(), ct_msg.data()), "kb_encrypt_batch");
  out.resize(n);
  for (size_t i = 0; i < n; i++) {
    std::memcpy(out[i].first.xy, &ct[32 * i], 128); out[i].first.inf = ct_inf[i] != 0;
    out[i].second.assign(ct_msg.begin() + off[i], ct_msg.begin() + off[i + 1]);
  }
}
inline std::vector<Bytes> decrypt_batch(kb_ctx* c, const G1* proofs, const Ciphertext* const* cts, size_t n) {
  std::vector<uint64_t> off(n + 1, 0);
  for (size_t i = 0; i < n; i++) off[i + 1] = off[i] + cts[i]->second.size();
  std::vector<uint8_t> flat(off[n] ? off[n] : 1), outb(off[n] ? off[n] : 1), pinf(n ? n : 1), cinf(n ? n : 1);
  std::vector<uint32_t> pxy(16 * (n ? n : 1)), cxy(32 * (n ? n : 1));
  for (size_t i = 0; i < n; i++) {
    std::memcpy(&pxy[16 * i], proofs[i].xy, 64); pinf[i] = proofs[i].inf;
    std::memcpy(&cxy[32 * i], cts[i]->first.xy, 128); cinf[i] = cts[i]->first.inf;
    if (!cts[i]->second.empty()) std::memcpy(&flat[off[i]], cts[i]->second.data(), cts[i]->second.size());
  }
  check(c, kb_decrypt_batch(c, pxy.data(), pinf.data(), cxy.data(), cinf.data(), flat.data(), off.data(), n, outb.data()), "kb_decrypt_batch");
  std::vector<Bytes> out(n);
  for (size_t i = 0; i < n; i++) out[i].assign(outb.begin() + off[i], outb.begin() + off[i + 1]);
  return out;
}
}  // namespace detail

/// src/enc.rs:19-40
template <class Rng>
inline Ciphertext encrypt(Rng& rng, const KZGSetup& kzg_setup, const G1& com, const Fr& point, const Fr& value, const Bytes& msg) {
  Fr r = Fr::rand(rng);                                   // src/kem.rs:26
  std::vector<Ciphertext> out;
  detail::encrypt_batch(kzg_setup, com, &point, &value, &r, std::vector<Bytes>{msg}, out);
  return out[0];
}
/// src/enc.rs:44-55
inline Bytes decrypt(const G1& proof, const Ciphertext& ct) {
  const Ciphertext* p = &ct;
  return detail::decrypt_batch(detail::default_ctx(), &proof, &p, 1)[0];
}
/// src/kem.rs:13-50 -> (ciphertext point, key of key_len bytes): encrypting zeros leaves the key in the masked message
template <class Rng>
inline std::pair<G2, Bytes> encapsulate(Rng& rng, const KZGSetup& kzg_setup, const G1& com, const Fr& point, const Fr& value, size_t key_len) {
  return encrypt(rng, kzg_setup, com, point, value, Bytes(key_len, 0));
}
/// src/kem.rs:55-72
inline Bytes decapsulate(const G1& proof, const G2& ct, size_t key_len) { return decrypt(proof, Ciphertext(ct, Bytes(key_len, 0))); }

// ---------------------------------------------------------------------------------------------
// Vector commitments (src/vec.rs)
// ---------------------------------------------------------------------------------------------
static constexpr size_t PADDING_LEN = 1;   // src/vec.rs:18

/// src/vec.rs:22-49 -> (commitment, proofs); one random element is appended, then iFFT, open_fk, commit
template <class Rng>
inline std::pair<G1, std::vector<G1>> vec_commit(Rng& rng, const KZGSetup& kzg_setup, const std::vector<Fr>& values) {
  std::vector<Fr> padded = values;
  for (size_t i = 0; i < PADDING_LEN; i++) padded.push_back(Fr::rand(rng));      // :29-33
  Radix2EvaluationDomain domain(padded.size());                                    // :36
  std::vector<Fr> p = domain.ifft(padded);                                         // :37
  std::vector<G1> proofs = open_fk(kzg_setup, p, domain);                          // :40-43
  G1 com = commit(kzg_setup, p);                                                   // :46
  return {com, proofs};
}
/// src/vec.rs:52-69: points[i], values[i], messages[i]; out-of-range indexing panics in the reference (std::out_of_range)
template <class Rng>
inline std::vector<Ciphertext> vec_encrypt(Rng& rng, const KZGSetup& kzg_setup, const G1& com, const std::vector<Fr>& points,
                                           const std::vector<Fr>& values, const std::vector<Bytes>& messages) {
  const size_t n = messages.size();
  if (points.size() < n || values.size() < n) throw std::out_of_range("vec_encrypt: index out of bounds");
  std::vector<Fr> rs(n);
  for (size_t i = 0; i < n; i++) rs[i] = Fr::rand(rng);      // same draws, same order as the loop at :63-66
  std::vector<Ciphertext> out;
  detail::encrypt_batch(kzg_setup, com, points.data(), values.data(), rs.data(), messages, out);
  return out;
}
/// src/vec.rs:72-81
inline std::vector<Bytes> vec_decrypt(const std::vector<G1>& proofs, const std::vector<const Ciphertext*>& cts) {
  if (proofs.size() < cts.size()) throw std::out_of_range("vec_decrypt: index out of bounds");
  return detail::decrypt_batch(detail::default_ctx(), proofs.data(), cts.data(), cts.size());
}

// ---------------------------------------------------------------------------------------------
// Laconic OT (tests/laconic_ot.rs:15-113)
// ---------------------------------------------------------------------------------------------
class Receiver {
 public:
  /// tests/laconic_ot.rs:26-39: vec_commit over the choices gives the commitment and the precomputed openings
  template <class Rng>
  Receiver(const KZGSetup& kzg_setup, Rng& rng, std::vector<Fr> choices) : choices_(std::move(choices)) {
    auto cp = vec_commit(rng, kzg_setup, choices_);
    commitment_ = cp.first; proofs_ = std::move(cp.second);
  }
  const G1& commitment() const { return commitment_; }
  const std::vector<G1>& proofs() const { return proofs_; }
  /// tests/laconic_ot.rs:41-59: position i takes the ciphertext of set 0 when choices[i] == 0, of set 1 otherwise
  std::vector<Bytes> receive(const std::vector<std::vector<Ciphertext>>& encrypted_sets) const {
    const size_t n_choices = encrypted_sets.at(0).size();
    std::vector<const Ciphertext*> chosen_cts;
    chosen_cts.reserve(n_choices);
    for (size_t i = 0; i < n_choices; i++) chosen_cts.push_back(&encrypted_sets.at(choices_.at(i).is_zero() ? 0 : 1).at(i));
    return vec_decrypt(proofs_, chosen_cts);
  }

 private:
  std::vector<Fr> choices_;
  G1 commitment_;
  std::vector<G1> proofs_;
};

class Sender {
 public:
  Sender(const KZGSetup& kzg_setup, const G1& commitment) : setup_(kzg_setup), commitment_(commitment) {}   // :70-75
  /// tests/laconic_ot.rs:77-112: the points are the domain elements of size n + PADDING_LEN; set 0 is encrypted under
  /// value 0, then set 1 under value 1 (the generator is consumed in that order)
  template <class Rng>
  std::vector<std::vector<Ciphertext>> send(Rng& rng, const std::vector<std::vector<Bytes>>& private_set) const {
    const size_t n_values = private_set.at(0).size();
    std::vector<Fr> elements = Radix2EvaluationDomain(n_values + PADDING_LEN).elements();
    std::vector<std::vector<Ciphertext>> encrypted_messages;
    encrypted_messages.push_back(vec_encrypt(rng, setup_, commitment_, elements, std::vector<Fr>(n_values, Fr::zero()), private_set.at(0)));
    encrypted_messages.push_back(vec_encrypt(rng, setup_, commitment_, elements, std::vector<Fr>(n_values, Fr::one()), private_set.at(1)));
    return encrypted_messages;
  }

 private:
  const KZGSetup& setup_;
  G1 commitment_;
};

}  // namespace keaki
#endif  // KEAKI_B200_HPP
