// The reference's own tests (src/kzg.rs:212-505, src/kem.rs:74-224, src/enc.rs:58-125, src/kzg/ptau.rs:476-514,
// tests/laconic_ot.rs:114-200) re-instantiated on BN254 against the C++ host layer (include/keaki_b200.hpp) —
// every group, pairing and transform operation below runs in libkeaki_b200.so on the GPU.
//
//   test_keaki_host [path/to/ppot_0080_01_mini.ptau [path/to/oracle_vectors.txt]]
// exit codes: 0 all passed, 1 a check failed, 3 no usable GPU (the product path has no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include "../../include/keaki_b200.hpp"

using namespace keaki;

static int g_failed = 0, g_checks = 0;
#define CHECK(cond) do { g_checks++; if (!(cond)) { g_failed++; std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } } while (0)
#define RUN(t) do { std::fprintf(stderr, "[ RUN  ] %s\n", #t); t(); std::fprintf(stderr, "[ DONE ] %s\n", #t); } while (0)

// k * G1 through the library (the `G1Affine::generator().mul(k)` of the reference tests)
static G1 g1_gen_mul(const KZGSetup& s, const Fr& k) {
  G1 out; uint8_t inf = 1;
  detail::check(s.ctx(), kb_g1_mul_gen_batch(s.ctx(), detail::u32(&k), 1, out.xy, &inf), "kb_g1_mul_gen_batch");
  out.inf = inf != 0;
  return out;
}
static Bytes pairing_bytes(const KZGSetup& s, const G1& p, const G2& q) {
  Bytes gt(384); uint8_t pi = p.inf, qi = q.inf;
  detail::check(s.ctx(), kb_pairing_batch(s.ctx(), p.xy, &pi, q.xy, &qi, 1, gt.data()), "kb_pairing_batch");
  return gt;
}
static std::vector<Fr> poly(std::initializer_list<long long> c) { std::vector<Fr> p; for (long long x : c) p.push_back(Fr::from_i64(x)); return p; }

// ---------------------------------------------------------------------------------------------- Fr (host arithmetic)
static void test_fr_host_arithmetic() {
  CHECK(Fr::from_u64(3) * Fr::from_u64(5) == Fr::from_u64(15));
  CHECK(Fr::from_i64(-24) + Fr::from_u64(24) == Fr::zero());
  CHECK(Fr::from_u64(7).inverse() * Fr::from_u64(7) == Fr::one());
  CHECK(Fr::from_u64(2).pow(10) == Fr::from_u64(1024));
  uint64_t c[4]; Fr::from_u64(113562).to_canonical(c);
  CHECK(c[0] == 113562 && c[1] == 0 && c[2] == 0 && c[3] == 0);
  CHECK(evaluate(poly({-24, -25, -5, 9, 7}), Fr::from_u64(11)) == Fr::from_u64(113562));   // src/kzg.rs:352
}

// ---------------------------------------------------------------------------------------------- src/kzg.rs tests
static void test_kzg_setup() {                                         // :218-239
  SplitMix64 rng(1);
  Fr secret = Fr::rand(rng);
  const size_t max_degree = 4;
  KZGSetup s = KZGSetup::setup(secret, max_degree);
  CHECK(s.g1_pow().size() == max_degree);
  for (size_t i = 0; i < max_degree; i++) CHECK(s.g1_pow()[i] == g1_gen_mul(s, secret.pow(i)));
  // [tau]_2 == tau * G2: e(G1, [tau]_2) == e(tau G1, G2), with G2's generator taken from a setup with secret 1
  KZGSetup unit = KZGSetup::setup(Fr::one(), 1);
  CHECK(pairing_bytes(s, s.g1_pow()[0], s.tau_g2()) == pairing_bytes(s, g1_gen_mul(s, secret), unit.tau_g2()));
}
static void test_kzg_commit() {                                        // :241-258
  SplitMix64 rng(2);
  Fr secret = Fr::rand(rng);
  KZGSetup s = KZGSetup::setup(secret, 4);
  std::vector<Fr> p = poly({1, 3, 2});
  G1 commitment = commit(s, p);
  // sum_i coeff_i [tau^i]_1 = [p(tau)]_1
  CHECK(commitment == g1_gen_mul(s, evaluate(p, secret)));
  CHECK(commit(s, poly({0, 0, 0})).inf);                               // the zero polynomial commits to the identity
}
static void test_kzg_commit_polynomial_too_large() {                   // :260-277
  SplitMix64 rng(3);
  KZGSetup s = KZGSetup::setup(Fr::rand(rng), 2);
  bool threw = false;
  try { commit(s, poly({1, 3, 2, 4})); } catch (const KZGError& e) { threw = e.len == 4 && e.max == 2 && std::string(e.what()) == "PolynomialTooLarge(4, 2)"; }
  CHECK(threw);
}
static void test_kzg_open_polynomial_too_large() {                     // :279-308 (the quotient has 5 coefficients)
  SplitMix64 rng(4);
  KZGSetup s = KZGSetup::setup(Fr::rand(rng), 2);
  bool threw = false;
  try { open(s, poly({1, 2, 3, 4, 5, 6}), Fr::from_u64(5)); } catch (const KZGError& e) { threw = e.len == 5 && e.max == 2; }
  CHECK(threw);
}
static void test_kzg_open_and_verify() {                               // :310-331
  SplitMix64 rng(5);
  KZGSetup s = KZGSetup::setup(Fr::rand(rng), 4);
  std::vector<Fr> p = poly({1, 3, 2});
  G1 commitment = commit(s, p);
  Fr point = Fr::from_u64(5), expected_value = Fr::from_u64(66);
  G1 proof = open(s, p, point);
  CHECK(verify(s, commitment, point, expected_value, proof));
}
static void test_kzg_verify_negative_cases() {                         // :333-468
  SplitMix64 rng(6);
  KZGSetup s = KZGSetup::setup(Fr::rand(rng), 8);
  std::vector<Fr> p = poly({-24, -25, -5, 9, 7});
  G1 commitment = commit(s, p);
  Fr point = Fr::from_u64(11), value = Fr::from_u64(113562);
  G1 proof = open(s, p, point);
  CHECK(verify(s, commitment, point, value, proof));
  CHECK(!verify(s, commitment, Fr::from_u64(99), value, proof));                       // wrong alpha
  CHECK(!verify(s, commitment, point, Fr::from_u64(12345), proof));                    // wrong beta
  CHECK(!verify(s, commitment, point, value, open(s, p, Fr::from_u64(12))));           // wrong proof
  CHECK(!verify(s, commit(s, poly({1, 2, 3})), point, value, proof));                  // wrong commitment
}
static void test_kzg_open_fk() {                                       // :470-505
  SplitMix64 rng(7);
  KZGSetup s = KZGSetup::setup(Fr::rand(rng), 16);
  std::vector<Fr> p = poly({1, 2, 3, 4});
  Radix2EvaluationDomain domain(p.size());
  std::vector<G1> proofs = open_fk(s, p, domain);
  std::vector<Fr> roots = domain.elements();
  CHECK(roots.size() == 4 && roots[0] == Fr::one() && roots[1].pow(4) == Fr::one() && roots[1].pow(2) != Fr::one());
  for (size_t i = 0; i < p.size(); i++) CHECK(proofs[i] == open(s, p, roots[i]));
  // fft / ifft are inverse transforms and agree with evaluation
  std::vector<Fr> ev = domain.fft(p);
  for (size_t i = 0; i < 4; i++) CHECK(ev[i] == evaluate(p, roots[i]));
  CHECK(domain.ifft(ev) == p);
}

// ---------------------------------------------------------------------------------------------- src/kem.rs, src/enc.rs tests
static void test_encapsulation_decapsulation() {                       // src/kem.rs:87-224
  SplitMix64 rng(8);
  KZGSetup s = KZGSetup::setup(Fr::rand(rng), 10);
  std::vector<Fr> p = poly({-24, -25, -5, 9, 7});
  Fr point = Fr::rand(rng), eval = evaluate(p, point);
  G1 commitment = commit(s, p);
  auto ck = encapsulate(rng, s, commitment, point, eval, 32);
  G1 proof = open(s, p, point);
  CHECK(ck.second.size() == 32 && decapsulate(proof, ck.first, 32) == ck.second);
  CHECK(decapsulate(open(s, p, Fr::rand(rng)), ck.first, 32) != ck.second);              // invalid proof
  auto wrong_val = encapsulate(rng, s, commitment, point, eval + Fr::one(), 32);           // wrong value
  CHECK(decapsulate(proof, wrong_val.first, 32) != wrong_val.second);
  auto wrong_pt = encapsulate(rng, s, commitment, point + Fr::one(), eval, 32);            // wrong point
  CHECK(decapsulate(proof, wrong_pt.first, 32) != wrong_pt.second);
  // key length is free (XOF): prefixes agree
  Bytes k64 = decapsulate(proof, ck.first, 64);
  CHECK(Bytes(k64.begin(), k64.begin() + 32) == ck.second);
}
static void test_encrypt_decrypt() {                                   // src/enc.rs:70-125
  SplitMix64 rng(9);
  KZGSetup s = KZGSetup::setup(Fr::rand(rng), 10);
  std::vector<Fr> p = poly({-24, -25, -5, 9, 7});
  Fr point = Fr::rand(rng), val = evaluate(p, point);
  G1 commitment = commit(s, p);
  const char* hw = "helloworld";
  Bytes msg(hw, hw + 10);
  Ciphertext ct = encrypt(rng, s, commitment, point, val, msg);
  CHECK(decrypt(open(s, p, point), ct) == msg);
  CHECK(decrypt(open(s, p, Fr::rand(rng)), ct) != msg);
}
static void test_vec_commit_encrypt_decrypt() {                        // src/vec.rs
  SplitMix64 rng(10);
  KZGSetup s = KZGSetup::setup(Fr::rand(rng), 16);
  std::vector<Fr> values;
  for (int i = 0; i < 7; i++) values.push_back(Fr::rand(rng));
  auto cp = vec_commit(rng, s, values);
  CHECK(cp.second.size() == 8);                                        // domain of size next_pow2(7 + PADDING_LEN)
  std::vector<Fr> points = Radix2EvaluationDomain(values.size() + PADDING_LEN).elements();
  for (size_t i = 0; i < values.size(); i++) CHECK(verify(s, cp.first, points[i], values[i], cp.second[i]));
  std::vector<Bytes> msgs;
  for (int i = 0; i < 7; i++) { Bytes m(5 + 9 * i); rng.fill_bytes(m.data(), m.size()); msgs.push_back(m); }   // ragged lengths, one > 64
  std::vector<Ciphertext> cts = vec_encrypt(rng, s, cp.first, points, values, msgs);
  std::vector<const Ciphertext*> refs;
  for (auto& c : cts) refs.push_back(&c);
  CHECK(vec_decrypt(cp.second, refs) == msgs);
}

// ---------------------------------------------------------------------------------------------- tests/laconic_ot.rs
static void laconic_ot_on(const KZGSetup& kzg, size_t n_choices, uint64_t seed) {
  SplitMix64 rng(seed);
  std::vector<Fr> choices;
  for (size_t i = 0; i < n_choices; i++) choices.push_back((rng.next_u64() & 1) ? Fr::one() : Fr::zero());
  Receiver receiver(kzg, rng, choices);
  std::vector<std::vector<Bytes>> private_set(2);
  for (int b = 0; b < 2; b++)
    for (size_t i = 0; i < n_choices; i++) { Bytes v(32); rng.fill_bytes(v.data(), 32); private_set[b].push_back(v); }
  Sender sender(kzg, receiver.commitment());
  std::vector<std::vector<Ciphertext>> encrypted = sender.send(rng, private_set);
  std::vector<Bytes> dec = receiver.receive(encrypted);
  CHECK(dec.size() == n_choices);
  for (size_t i = 0; i < n_choices; i++) CHECK(dec[i] == private_set[choices[i].is_zero() ? 0 : 1][i]);
}
static void test_laconic_ot() {                                        // :126-200 (SETUP_DEGREE 16, N_CHOICES 8)
  SplitMix64 rng(11);
  KZGSetup kzg = KZGSetup::setup(Fr::rand(rng), 16);
  laconic_ot_on(kzg, 8, 12);
}

// ---------------------------------------------------------------------------------------------- src/kzg/ptau.rs:476-514 + BASELINE config 1
static std::string g_ptau;
static void test_new_from_file_and_config1() {
  KZGSetup s = KZGSetup::new_from_file(g_ptau);
  CHECK(s.g1_pow().size() == 3);                                       // 2 * 2^1 - 1 powers in the reference's fixture
  // the decoded powers are a genuine SRS: e([tau^(i+1)]_1, G2) == e([tau^i]_1, [tau]_2)
  KZGSetup unit = KZGSetup::setup(Fr::one(), 1);
  for (int i = 0; i < 2; i++) CHECK(pairing_bytes(s, s.g1_pow()[i + 1], unit.tau_g2()) == pairing_bytes(s, s.g1_pow()[i], s.tau_g2()));
  laconic_ot_on(s, 1, 13);                                             // 1 choice + 1 pad -> domain of size 2 (config 1)
  bool threw = false;
  try { KZGSetup::new_from_file(g_ptau + ".does-not-exist"); } catch (const SetupFileError&) { threw = true; }
  CHECK(threw);
}

// ---------------------------------------------------------------------------------------------- committed golden bytes
// tests/golden/oracle_vectors.txt (written by tests/golden/make_golden.py from the oracle): the C++ layer must
// reproduce the oracle's commitment, FK proofs, ciphertext points and masked messages byte for byte.
static std::string g_golden;
static Bytes unhex(const std::string& h) {
  Bytes b;
  if (h == "-") return b;
  for (size_t i = 0; i + 1 < h.size(); i += 2) b.push_back((uint8_t)std::stoul(h.substr(i, 2), nullptr, 16));
  return b;
}
static Fr fr_from_le(const Bytes& b) { uint64_t c[4]; std::memcpy(c, b.data(), 32); return Fr::from_canonical(c); }
struct FixedRng {   // hands out prepared field elements through the `Fr::rand` protocol (four limbs of the Montgomery form)
  std::vector<Fr> xs; size_t i = 0, k = 0;
  uint64_t next_u64() { uint64_t v = xs.at(i).l[k]; if (++k == 4) { k = 0; i++; } return v; }
};
static Bytes g1_bytes(const KZGSetup& s, const G1& p) {
  Bytes out(64); uint8_t inf = p.inf;
  detail::check(s.ctx(), kb_g1_serialize(s.ctx(), p.xy, &inf, 1, 0, out.data()), "kb_g1_serialize");
  return out;
}
static Bytes g2_bytes(const KZGSetup& s, const G2& p) {
  Bytes out(128); uint8_t inf = p.inf;
  detail::check(s.ctx(), kb_g2_serialize(s.ctx(), p.xy, &inf, 1, 0, out.data()), "kb_g2_serialize");
  return out;
}
static void test_golden_vectors() {
  std::ifstream f(g_golden);
  CHECK(f.good());
  std::map<std::string, std::vector<Bytes>> v;
  for (std::string line; std::getline(f, line);) {
    std::istringstream ss(line);
    std::string key, tok;
    ss >> key;
    while (ss >> tok) v[key].push_back(unhex(tok));
  }
  Fr tau = fr_from_le(v["tau"].at(0));
  std::vector<Fr> p, points, values;
  for (auto& b : v["coeffs"]) p.push_back(fr_from_le(b));
  for (auto& b : v["points"]) points.push_back(fr_from_le(b));
  for (auto& b : v["values"]) values.push_back(fr_from_le(b));
  KZGSetup s = KZGSetup::setup(tau, p.size());
  G1 com = commit(s, p);
  CHECK(g1_bytes(s, com) == v["commitment"].at(0));
  std::vector<G1> proofs = open_fk(s, p, Radix2EvaluationDomain(p.size()));
  CHECK(proofs.size() == v["proofs"].size());
  for (size_t i = 0; i < proofs.size(); i++) CHECK(g1_bytes(s, proofs[i]) == v["proofs"][i]);
  FixedRng rng;
  for (auto& b : v["r"]) rng.xs.push_back(fr_from_le(b));
  std::vector<Ciphertext> cts = vec_encrypt(rng, s, com, points, values, v["messages"]);
  CHECK(cts.size() == v["ct"].size());
  for (size_t i = 0; i < cts.size(); i++) { CHECK(g2_bytes(s, cts[i].first) == v["ct"][i]); CHECK(cts[i].second == v["msg_ct"][i]); }
  std::vector<const Ciphertext*> refs;
  for (auto& c : cts) refs.push_back(&c);
  CHECK(vec_decrypt(proofs, refs) == v["messages"]);
}

int main(int argc, char** argv) {
  g_ptau = argc > 1 ? argv[1] : "tests/golden/ppot_0080_01_mini.ptau";
  g_golden = argc > 2 ? argv[2] : "tests/golden/oracle_vectors.txt";
  RUN(test_fr_host_arithmetic);
  try {
    RUN(test_kzg_setup);
  } catch (const BackendError& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return g_failed ? 1 : 3;
  }
  try {
    RUN(test_kzg_commit);
    RUN(test_kzg_commit_polynomial_too_large);
    RUN(test_kzg_open_polynomial_too_large);
    RUN(test_kzg_open_and_verify);
    RUN(test_kzg_verify_negative_cases);
    RUN(test_kzg_open_fk);
    RUN(test_encapsulation_decapsulation);
    RUN(test_encrypt_decrypt);
    RUN(test_vec_commit_encrypt_decrypt);
    RUN(test_laconic_ot);
    RUN(test_new_from_file_and_config1);
    RUN(test_golden_vectors);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "unexpected exception: %s\n", e.what());
    return 1;
  }
  std::fprintf(stderr, "%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
