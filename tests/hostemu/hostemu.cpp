// TEST-ONLY host build of the device arithmetic headers (keaki_b200/csrc/*.cuh).
//
// The headers compile for the host with the PTX carry-chain primitives replaced by an emulated
// carry flag, so the exact algorithm text that runs on the GPU (Montgomery multiplier, tower, curve
// formulas, Miller loop, final exponentiation, BLAKE3) can be checked against oracle/ on a box
// without a GPU.  This library is built into tests/hostemu/_build/ and is loaded ONLY by tests/;
// the product library (libkeaki_b200.so) contains no host arithmetic path.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>
#include "../../keaki_b200/csrc/pairing.cuh"
#include "../../keaki_b200/csrc/glv.cuh"
#include "../../keaki_b200/csrc/blake3.cuh"
#include "../../keaki_b200/csrc/consts_gen.cuh"
#include "../../keaki_b200/csrc/msm_digits.cuh"
#include "../../keaki_b200/csrc/pairing_vm.cuh"
#include "../../keaki_b200/csrc/pairing_prog_gen.cuh"
#include "../../keaki_b200/csrc/pairing_st.cuh"
#include "../../keaki_b200/csrc/pairing_warp.cuh"
#include "../../keaki_b200/csrc/pairing_warp_gen.cuh"

using namespace kb;

static Fq ldq(const uint32_t* p) { return fp_load<FqParams>(p); }
static Fr ldr(const uint32_t* p) { return fp_load<FrParams>(p); }
static Fq2 ldq2(const uint32_t* p) { Fq2 r; r.c0 = ldq(p); r.c1 = ldq(p + 8); return r; }
static void stq(uint32_t* p, const Fq& a) { fp_store<FqParams>(p, a); }
static void stq2(uint32_t* p, const Fq2& a) { stq(p, a.c0); stq(p + 8, a.c1); }
static Fq12 ldq12(const uint32_t* p) {
  Fq12 r;
  Fq6* h[2] = {&r.c0, &r.c1};
  for (int j = 0; j < 2; j++) { h[j]->c0 = ldq2(p + (j * 3 + 0) * 16); h[j]->c1 = ldq2(p + (j * 3 + 1) * 16); h[j]->c2 = ldq2(p + (j * 3 + 2) * 16); }
  return r;
}
static void stq12(uint32_t* p, const Fq12& a) {
  const Fq6* h[2] = {&a.c0, &a.c1};
  for (int j = 0; j < 2; j++) { stq2(p + (j * 3 + 0) * 16, h[j]->c0); stq2(p + (j * 3 + 1) * 16, h[j]->c1); stq2(p + (j * 3 + 2) * 16, h[j]->c2); }
}
static PairingConsts make_consts() {
  PairingConsts pc;
  for (int k = 0; k < 3; k++) for (int i = 0; i < 6; i++) pc.frob.g[k][i] = ldq2(consts::FROB_GAMMA + (k * 6 + i) * 16);
  pc.tw_x = ldq2(consts::TW_X);
  pc.tw_y = ldq2(consts::TW_Y);
  return pc;
}
static const PairingConsts& pcs() { static PairingConsts pc = make_consts(); return pc; }

extern "C" {

// op: 0 add, 1 sub, 2 mul, 3 neg(a), 4 inv(a), 5 from_mont(a), 6 to_mont(a), 7 sqr(a); field: 0 Fq, 1 Fr
void he_fp_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, int n) {
  for (int i = 0; i < n; i++) {
    if (field == 0) {
      Fq x = ldq(a + 8 * i), y = ldq(b + 8 * i), r;
      switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break; case 3: r = -x; break;
        case 4: r = inv(x); break; case 5: r = fp_from_mont<FqParams>(x); break; case 6: r = fp_to_mont<FqParams>(x); break; default: r = sqr(x); }
      stq(out + 8 * i, r);
    } else {
      Fr x = ldr(a + 8 * i), y = ldr(b + 8 * i), r;
      switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break; case 3: r = -x; break;
        case 4: r = inv(x); break; case 5: r = fp_from_mont<FrParams>(x); break; case 6: r = fp_to_mont<FrParams>(x); break; default: r = sqr(x); }
      fp_store<FrParams>(out + 8 * i, r);
    }
  }
}

// op: 0 mul, 1 sqr(a), 2 inv(a), 3 mul_xi(a)
void he_fq2_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fq2 x = ldq2(a), y = ldq2(b), r;
  switch (op) { case 0: r = x * y; break; case 1: r = sqr(x); break; case 2: r = inv(x); break; default: r = mul_xi(x); }
  stq2(out, r);
}

// op: 0 mul, 1 sqr(a), 2 inv(a), 3 cyclotomic_sqr(a), 4..6 frobenius k=1..3, 7 conj
void he_fq12_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fq12 x = ldq12(a), y = ldq12(b), r;
  switch (op) {
    case 0: r = x * y; break; case 1: r = sqr(x); break; case 2: r = inv(x); break; case 3: r = cyclotomic_sqr(x); break;
    case 4: case 5: case 6: r = frobenius(x, op - 3, pcs().frob); break; default: r = conj(x);
  }
  stq12(out, r);
}

// easy part of the final exponentiation (maps into the cyclotomic subgroup) - to build inputs for op 3 above
void he_fq12_easy(const uint32_t* a, uint32_t* out) {
  Fq12 f = ldq12(a);
  Fq12 r = conj(f) * inv(f);
  r = frobenius(r, 2, pcs().frob) * r;
  stq12(out, r);
}

// sparse line product: out = a * (l0 + l1 w + l3 w^3)
void he_mul_by_line(const uint32_t* a, const uint32_t* l0, const uint32_t* l1, const uint32_t* l3, uint32_t* out) {
  stq12(out, mul_by_line(ldq12(a), ldq2(l0), ldq2(l1), ldq2(l3)));
}

// G1: out_xy (affine, Montgomery; zeros = infinity) = k * P (k canonical 8 limbs) [+ Q if q != null]
void he_g1_mul_add(const uint32_t* p_xy, const uint32_t* k, const uint32_t* q_xy, uint32_t* out_xy) {
  G1Affine p; p.x = ldq(p_xy); p.y = ldq(p_xy + 8);
  G1 r = ec_mul(to_xyzz(p), k);
  if (q_xy) { G1Affine q; q.x = ldq(q_xy); q.y = ldq(q_xy + 8); r = ec_add_mixed(r, q); }
  G1Affine a = to_affine(r);
  stq(out_xy, a.x); stq(out_xy + 8, a.y);
}
// full (non-mixed) addition path: out = k1*P + k2*Q computed separately then ec_add
void he_g1_lincomb(const uint32_t* p_xy, const uint32_t* k1, const uint32_t* q_xy, const uint32_t* k2, uint32_t* out_xy) {
  G1Affine p, q; p.x = ldq(p_xy); p.y = ldq(p_xy + 8); q.x = ldq(q_xy); q.y = ldq(q_xy + 8);
  G1Affine a = to_affine(ec_add(ec_mul(to_xyzz(p), k1), ec_mul(to_xyzz(q), k2)));
  stq(out_xy, a.x); stq(out_xy + 8, a.y);
}
// GLV path: out_xy = k * P by g1_mul_glv; split (optional, 12 words) = m1[5], m2[5], neg1, neg2 of glv_decompose(k)
void he_g1_mul_glv(const uint32_t* p_xy, const uint32_t* k, uint32_t* out_xy, uint32_t* split) {
  G1Affine p; p.x = ldq(p_xy); p.y = ldq(p_xy + 8);
  G1Affine a = to_affine(g1_mul_glv(to_xyzz(p), k));
  stq(out_xy, a.x); stq(out_xy + 8, a.y);
  if (split) {
    GlvSplit s = glv_decompose(k);
    for (int i = 0; i < 5; i++) { split[i] = s.m1[i]; split[5 + i] = s.m2[i]; }
    split[10] = s.neg1; split[11] = s.neg2;
  }
}
void he_g2_mul_add(const uint32_t* p_xy, const uint32_t* k, const uint32_t* q_xy, uint32_t* out_xy) {
  G2Affine p; p.x = ldq2(p_xy); p.y = ldq2(p_xy + 16);
  G2 r = ec_mul(to_xyzz(p), k);
  if (q_xy) { G2Affine q; q.x = ldq2(q_xy); q.y = ldq2(q_xy + 16); r = ec_add_mixed(r, q); }
  G2Affine a = to_affine(r);
  stq2(out_xy, a.x); stq2(out_xy + 16, a.y);
}

// Miller loop only (Montgomery Fq12 out), full pairing as 384 canonical bytes, and key derivation
void he_miller(const uint32_t* p_xy, const uint32_t* q_xy, uint32_t* out) {
  G1Affine p; p.x = ldq(p_xy); p.y = ldq(p_xy + 8);
  G2Affine q; q.x = ldq2(q_xy); q.y = ldq2(q_xy + 16);
  stq12(out, miller_loop(p, q, pcs()));
}
void he_final_exp(const uint32_t* f, uint32_t* out) { stq12(out, final_exponentiation(ldq12(f), pcs())); }
void he_pairing_bytes(const uint32_t* p_xy, const uint32_t* q_xy, uint8_t* out384) {
  G1Affine p; p.x = ldq(p_xy); p.y = ldq(p_xy + 8);
  G2Affine q; q.x = ldq2(q_xy); q.y = ldq2(q_xy + 16);
  Fq12 e = final_exponentiation(miller_loop(p, q, pcs()), pcs());
  uint32_t w[96];
  gt_to_words(e, w);
  memcpy(out384, w, 384);
}
void he_gt_key(const uint8_t* gt384, const uint8_t* msg, uint8_t* out, uint64_t len) {
  uint32_t w[96];
  memcpy(w, gt384, 384);
  b3_gt_xof_xor(w, msg, out, len);
}

}  // extern "C"

// The pairing VM (pairing_vm.cuh) interpreting the shipped program for `slots` slots: GT as 384 canonical bytes.
// Exactly the interpreter + program the GPU kernel runs; the two lanes of a pairing are two host threads, the
// shuffle exchange and sync() are rendezvous of the two.
struct Barrier {
  std::atomic<int> arrived{0};
  std::atomic<int> gen{0};
  void wait() {
    int g0 = gen.load();
    if (arrived.fetch_add(1) == 1) { arrived.store(0); gen.fetch_add(1); }
    else while (gen.load() == g0) std::this_thread::yield();
  }
};
struct PairShared {
  Fq s[2][32];
  Fq g[2][256];
  Fq box[2];
  Barrier bar;
};
struct HostLane {
  PairShared* sh;
  uint32_t t;
  Fq ld(uint32_t i) const { return sh->s[t][i]; }
  Fq ld_partner(uint32_t i) const { return sh->s[1 - t][i]; }
  Fq ld_c(uint32_t i, uint32_t h) const { return sh->s[h][i]; }
  void st(uint32_t i, const Fq& v) { sh->s[t][i] = v; }
  Fq ldg(uint32_t i) const { return sh->g[t][i]; }
  void stg(uint32_t i, const Fq& v) { sh->g[t][i] = v; }
  Fq ldc(const uint32_t* p) const { return ldq(p + 8 * t); }
  Fq xchg(const Fq& v) {
    sh->box[t] = v;
    sh->bar.wait();
    Fq r = sh->box[1 - t];
    sh->bar.wait();
    return r;
  }
  void sync() { sh->bar.wait(); }
};
static void vm_host_lane(PairShared* sh, uint32_t t, const uint64_t* prog, const uint8_t* outs,
                         const uint32_t* p_xy, const uint32_t* q_xy, uint32_t* w) {
  HostLane ln{sh, t};
  ln.st(0, ldq(p_xy + 8 * t));
  ln.st(1, ldq(q_xy + 8 * t));
  ln.st(2, ldq(q_xy + 16 + 8 * t));
  vm::run(prog, ln, vmprog::CONSTS);
  ln.sync();
  for (int k = 0; k < 6; k++) {
    Fq c = fp_from_mont<FqParams>(ln.ld(outs[k]));
    for (int j = 0; j < 8; j++) w[16 * k + 8 * t + j] = c.v[j];
  }
}

extern "C" {

int he_vm_pairing_bytes(int slots, const uint32_t* p_xy, const uint32_t* q_xy, uint8_t* out384) {
  const vmprog::Program* pr = nullptr;
  for (int k = 0; k < vmprog::NUM_PROGRAMS; k++) if (vmprog::PROGRAMS[k].slots == slots) pr = &vmprog::PROGRAMS[k];
  if (!pr) return -1;
  std::vector<uint64_t> padded(pr->words, pr->words + pr->len);
  padded.push_back(0); padded.push_back(0);
  PairShared* sh = new PairShared();
  uint32_t w[96];
  std::thread other(vm_host_lane, sh, 1u, padded.data(), pr->out, p_xy, q_xy, w);
  vm_host_lane(sh, 0u, padded.data(), pr->out, p_xy, q_xy, w);
  other.join();
  memcpy(out384, w, 384);
  delete sh;
  return pr->len;
}

// (9 x + y) mod p helper of the pairing VM (x < p, y <= p; canonical or Montgomery alike: it is linear)
void he_mul9_add(const uint32_t* x, const uint32_t* y, uint32_t* out) { stq(out, vm::mul9_add(ldq(x), ldq(y))); }

// signed-digit recoding used by the MSM (window c bits): digits[w] in [-2^(c-1), 2^(c-1)]
void he_msm_digits(const uint32_t* scalar_canonical, int c, int nwin, int32_t* digits) {
  msm_signed_digits(scalar_canonical, c, nwin, digits);
}

}  // extern "C"

// The compiled single-thread pairing (pairing_st.cuh) on a host slot store: GT as 384 canonical bytes.
struct StState { Fq2 on[16]; Fq2 sc[st::SCRATCH_SLOTS]; };   // on-chip slots: all 12 (the split is a device-memory detail)
struct StMem {
  StState* p;
  Fq2 ld(int a) const { return a < 16 ? p->on[a] : p->sc[a - 16]; }
  void st(int a, const Fq2& v) const { if (a < 16) p->on[a] = v; else p->sc[a - 16] = v; }
};

extern "C" {

void he_st_pairing_bytes(const uint32_t* p_xy, const uint32_t* q_xy, uint8_t* out384) {
  StState* s = new StState();
  StMem m{s};
  m.st(st::G_P, ldq2(p_xy));
  m.st(st::G_QX, ldq2(q_xy));
  m.st(st::G_QY, ldq2(q_xy + 16));
  uint32_t tw[32];
  memcpy(tw, consts::TW_X, 64); memcpy(tw + 16, consts::TW_Y, 64);
  st::miller(m, tw);
  st::final_exp(m, consts::FROB_GAMMA);
  uint32_t w[96];
  st::gt_words(m, w);
  memcpy(out384, w, 384);
  delete s;
}
// lazy-reduction pieces: op 0: out16 = mul_wide(a, b); op 1: out8 = redc(a16); op 2: out16 = fq2_mul_lazy(a16, b16)
void he_lazy_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (op == 0) { lz::W t = lz::mul_wide(ldq(a), ldq(b)); memcpy(out, t.v, 64); }
  else if (op == 1) { lz::W t; memcpy(t.v, a, 64); stq(out, lz::redc<FqParams>(t)); }
  else stq2(out, lz::fq2_mul_lazy(ldq2(a), ldq2(b)));
}
// Fq12-level routines of pairing_st.cuh on F: op 0 F*B, 1 F*conj(B), 2 F^2, 3 cyclotomic F^2, 4 F * line(b[0..47]), 5..7 frobenius 1..3, 8 inverse
void he_st_f12_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  StState* s = new StState();
  StMem m{s};
  for (int i = 0; i < 6; i++) m.st(st::F + i, ldq2(a + 16 * i));
  if (op <= 1) { for (int i = 0; i < 6; i++) m.st(st::G12(0) + i, ldq2(b + 16 * i)); st::f12mul(m, st::G12(0), op); }
  else if (op == 2) st::f12sqr(m);
  else if (op == 3) st::cycsqr(m);
  else if (op == 4) { for (int i = 0; i < 3; i++) m.st(st::S + i, ldq2(b + 16 * i)); st::f12mul_line(m); }
  else if (op <= 7) st::f12frob(m, op - 4, consts::FROB_GAMMA);
  else st::f12inv(m);
  for (int i = 0; i < 6; i++) stq2(out + 16 * i, m.ld(st::F + i));
  delete s;
}

}  // extern "C"

// The warp-cooperative pairing (pairing_warp.cuh): the shipped schedules run step by step, the 32 lanes of a step one
// after the other - all of them compute (read) before any of them stores, as on the GPU.
struct WpMem {
  Fq2* s;
  Fq2 ld(uint32_t a) const { return s[a]; }
  Fq ld_half(uint32_t a, uint32_t h) const { return h ? s[a].c1 : s[a].c0; }
  void st(uint32_t a, const Fq2& v) const { s[a] = v; }
  void st_half(uint32_t a, uint32_t h, const Fq& v) const { if (h) s[a].c1 = v; else s[a].c0 = v; }
};

extern "C" {

// prog 0: pairing (inputs P, Qx, Qy, (xP, 0), (yP, 0)), 1: window bases of a GT table (input: a cyclotomic Fq12, tower order),
// 2: product of 48 Fq12 values (tower order each).
// Inputs / outputs as 16 Montgomery limbs per Fq2.
int he_wp_run(int prog, const uint32_t* inputs, uint32_t* outs) {
  const wpprog::Program& p = prog == 0 ? wpprog::PAIRING : prog == 1 ? wpprog::GT_BASES : wpprog::GT_PROD;
  std::vector<Fq2> slots(p.nslots, Fq2::zero());
  for (int i = 0; i < p.ninputs; i++) slots[p.inputs[i]] = ldq2(inputs + 16 * i);
  for (int i = 0; i < p.nconsts; i++) slots[p.const_slot[i]] = ldq2(wpprog::CONSTS + 16 * p.const_idx[i]);
  WpMem m{slots.data()};
  std::vector<uint32_t> words((size_t)p.nsteps * 32 * 8);
  wpprog::expand(p, words.data());
  for (int s = 0; s < p.nsteps; s++) {
    Fq2 r[32];
    for (int lane = 0; lane < 32; lane++) r[lane] = wp::lane_compute(m, words.data() + 8 * (32 * s + lane));
    for (int lane = 0; lane < 32; lane++) wp::lane_store(m, words.data() + 8 * (32 * s + lane), r[lane]);
  }
  for (int i = 0; i < p.nouts; i++) stq2(outs + 16 * i, slots[p.outs[i]]);
  return p.nsteps;
}
// v9: 288-bit two's complement integer in (-128 q, 128 q)  ->  out8 = v mod q
void he_wp_reduce9(const uint32_t* v9, uint32_t* out8) {
  uint32_t v[9];
  memcpy(v, v9, 36);
  stq(out8, wp::reduce9(v));
}

}  // extern "C"
