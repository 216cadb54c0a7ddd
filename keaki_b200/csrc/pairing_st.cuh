// Compiled pairing: optimal-ate Miller loop + arkworks final exponentiation as ordinary structured code, ONE thread
// per pairing.  Replaces `E::pairing` (ark-ec 0.4.2 models/bn/mod.rs) at src/kem.rs:30,58 and src/kzg.rs:148 for the
// batched paths (decapsulate / vec_decrypt, verify).
//
// Round 1 interpreted a generated straight-line program with two lanes per pairing (pairing_vm.cuh): 53 % of the
// executed instructions were interpreter body.  Here the control flow is the (data-independent) loop structure itself
// and nothing is decoded at run time:
//   * the Fq12 accumulator F and one Fq12 of scratch S live in SHARED memory (12 Fq2 slots, 768 B per thread,
//     [slot][quarter][thread] so a warp's access is 512 contiguous bytes); everything that is touched once per loop
//     iteration (the G2 accumulator, P, Q) or once per final-exponentiation step (saved powers, the wNAF table) lives
//     in a per-thread global scratch laid out the same way (coalesced, L2-resident);
//   * Fq2 products use the lazy-reduction form of fpl.cuh (3 integer products + 2 reductions), squarings and
//     Fq2 x Fq products the fused multiplier;  Fq6 / Fq12 level: Karatsuba, complex squaring, Granger-Scott
//     cyclotomic squaring, 13-product sparse line multiplication, width-4 wNAF for the three powers of z;
//   * the field routines are out-of-line calls (operands by value in registers), the tower routines out-of-line
//     calls working on slot addresses, so the code stays small while no value ever goes through local memory.
// Line functions may be scaled by elements of proper subfields (they die in the final exponentiation), so the
// Jacobian formulas of pairing.cuh are reused; the result is arkworks' GT value bit for bit.
//
// Compiles for the host too (tests/hostemu, TEST ONLY).
#pragma once
#include "fpl.cuh"
#include "pairing.cuh"

namespace kb {
namespace st {

#if defined(__CUDACC__)
#define KB_ST_CALL static __device__ __noinline__
#define KB_ST_INL __device__ __forceinline__
#else
#define KB_ST_CALL inline
#define KB_ST_INL inline
#endif

// slot addresses: 0..5 = F (tower order c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2), 6..11 = S, 16 + k = scratch slot k.
// The memory object decides how many of the addresses below 12 are on chip (12: F and S; 9: F and half of S; 6: F
// only); the others fall to scratch slot `address` (6..11 are reserved for them).
enum : int { F = 0, S = 6, G = 16,
             G_TX = G + 0, G_TY = G + 1, G_TZ = G + 2, G_P = G + 3, G_QX = G + 4, G_QY = G + 5, G_CONJ = G + 12 };
KB_HD constexpr int G12(int k) { return G + 18 + 6 * k; }   // k-th saved Fq12 of the scratch
static constexpr int SCRATCH_SLOTS = 18 + 6 * 15;           // Fq2 slots of global scratch per thread
static constexpr int G_SAVE_F = G + SCRATCH_SLOTS;          // 6 more slots: F between two launches of the segmented form (pairing_st.cu)

// ------------------------------------------------------------------------------------------ Fq2 level
KB_ST_CALL Fq2 f2mul(Fq2 a, Fq2 b) { return lz::fq2_mul_lazy(a, b); }
KB_ST_CALL Fq2 f2sqr(Fq2 a) {   // (a0 + a1)(a0 - a1) + 2 a0 a1 u; the unreduced sums are valid multiplier inputs (< 2q)
  Fq2 r;
  r.c0 = fp_mul_inl<FqParams>(vm::add_nr(a.c0, a.c1), a.c0 - a.c1);
  r.c1 = fp_mul_inl<FqParams>(vm::add_nr(a.c0, a.c0), a.c1);
  return r;
}
KB_ST_CALL Fq2 f2mulfq(Fq2 a, Fq k) {
  Fq2 r;
  r.c0 = fp_mul_inl<FqParams>(a.c0, k);
  r.c1 = fp_mul_inl<FqParams>(a.c1, k);
  return r;
}
KB_ST_CALL Fq f1mul(Fq a, Fq b) { return fp_mul_inl<FqParams>(a, b); }
// (9 + u) a = (9 a0 - a1) + (9 a1 + a0) u
KB_ST_INL Fq2 xi(const Fq2& a) {
  Fq2 r;
  r.c0 = vm::mul9_add(a.c0, vm::neg_nz(a.c1));
  r.c1 = vm::mul9_add(a.c1, a.c0);
  return r;
}
KB_ST_INL Fq2 tpl(const Fq2& a) { return dbl(a) + a; }

struct F6 { Fq2 c0, c1, c2; };
KB_ST_INL Fq2 ld_const2(const uint32_t* p) {   // 16 limbs of a constant table
  Fq2 r;
#pragma unroll
  for (int i = 0; i < 8; i++) { r.c0.v[i] = p[i]; r.c1.v[i] = p[8 + i]; }
  return r;
}

// ------------------------------------------------------------------------------------------ Fq6 level (slot addresses)
// a + b without reduction (< 2q per coordinate): a valid operand of f2mul when the other operand is fully reduced
KB_ST_INL Fq2 add_loose(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = vm::add_nr(a.c0, b.c0); r.c1 = vm::add_nr(a.c1, b.c1); return r; }

// slots d..d+2 = (a..a+2) * (b..b+2) in Fq6 = Fq2[v]/(v^3 - xi) (Karatsuba: 6 products; the a-side sums stay
// unreduced).  All loads precede the stores, so d may alias a or b.
template <class M>
KB_ST_CALL void f6mul(M m, int d, int a, int b) {
  const Fq2 v0 = f2mul(m.ld(a), m.ld(b));
  const Fq2 v1 = f2mul(m.ld(a + 1), m.ld(b + 1));
  const Fq2 v2 = f2mul(m.ld(a + 2), m.ld(b + 2));
  Fq2 c0 = f2mul(add_loose(m.ld(a + 1), m.ld(a + 2)), m.ld(b + 1) + m.ld(b + 2)) - v1 - v2;
  c0 = v0 + xi(c0);
  const Fq2 c1 = f2mul(add_loose(m.ld(a), m.ld(a + 1)), m.ld(b) + m.ld(b + 1)) - v0 - v1 + xi(v2);
  const Fq2 c2 = f2mul(add_loose(m.ld(a), m.ld(a + 2)), m.ld(b) + m.ld(b + 2)) - v0 - v2 + v1;
  m.st(d, c0); m.st(d + 1, c1); m.st(d + 2, c2);
}
// (a..a+2) * (b0 + b1 v): 5 products
template <class M>
KB_ST_INL F6 f6m01(M& m, int a, const Fq2& b0, const Fq2& b1) {
  const Fq2 aa = f2mul(m.ld(a), b0), bb = f2mul(m.ld(a + 1), b1);
  F6 r;
  r.c0 = xi(f2mul(add_loose(m.ld(a + 1), m.ld(a + 2)), b1) - bb) + aa;
  r.c1 = f2mul(add_loose(m.ld(a), m.ld(a + 1)), b0 + b1) - aa - bb;
  r.c2 = f2mul(add_loose(m.ld(a), m.ld(a + 2)), b0) - aa + bb;
  return r;
}

// ------------------------------------------------------------------------------------------ Fq12 level (F, S on chip)
// F <- F * B, B = the Fq12 at slots b..b+5 (conjb: its conjugate, materialised at G_CONJ first).  Karatsuba over
// Fq6; S is scratch.
template <class M>
KB_ST_CALL void f12mul(M m, int b, int conjb) {
  if (conjb) {
    for (int i = 0; i < 3; i++) { m.st(G_CONJ + i, m.ld(b + i)); m.st(G_CONJ + 3 + i, -m.ld(b + 3 + i)); }
    b = G_CONJ;
  }
  f6mul(m, S, F, b);                     // t0 = a0 b0
  f6mul(m, S + 3, F + 3, b + 3);         // t1 = a1 b1
  for (int i = 0; i < 3; i++) m.st(F + i, m.ld(F + i) + m.ld(F + 3 + i));   // a0 + a1
  for (int i = 0; i < 3; i++) m.st(F + 3 + i, m.ld(b + i) + m.ld(b + 3 + i));   // b0 + b1
  f6mul(m, F + 3, F, F + 3);             // t2
  for (int i = 0; i < 3; i++) m.st(F + 3 + i, m.ld(F + 3 + i) - m.ld(S + i) - m.ld(S + 3 + i));   // c1 = t2 - t0 - t1
  m.st(F, m.ld(S) + xi(m.ld(S + 5)));    // c0 = t0 + v t1
  m.st(F + 1, m.ld(S + 1) + m.ld(S + 3));
  m.st(F + 2, m.ld(S + 2) + m.ld(S + 4));
}
// F <- F^2 (complex method: 2 Fq6 products)
template <class M>
KB_ST_CALL void f12sqr(M m) {
  f6mul(m, S, F, F + 3);                 // t = a0 a1
  for (int i = 0; i < 3; i++) m.st(S + 3 + i, m.ld(F + i) + m.ld(F + 3 + i));   // a0 + a1
  {                                      // a0 + v a1
    const Fq2 x0 = m.ld(F) + xi(m.ld(F + 5)), x1 = m.ld(F + 1) + m.ld(F + 3), x2 = m.ld(F + 2) + m.ld(F + 4);
    m.st(F, x0); m.st(F + 1, x1); m.st(F + 2, x2);
  }
  f6mul(m, F, S + 3, F);                 // u = (a0 + a1)(a0 + v a1)
  {                                      // c0 = u - t - v t,  c1 = 2 t
    const Fq2 t0 = m.ld(S), t1 = m.ld(S + 1), t2 = m.ld(S + 2);
    m.st(F, m.ld(F) - t0 - xi(t2));
    m.st(F + 1, m.ld(F + 1) - t1 - t0);
    m.st(F + 2, m.ld(F + 2) - t2 - t1);
    m.st(F + 3, dbl(t0)); m.st(F + 4, dbl(t1)); m.st(F + 5, dbl(t2));
  }
}
// (a + b s)^2 in Fq4 = Fq2[s]/(s^2 - xi)
KB_ST_INL void fq4sqr(const Fq2& a, const Fq2& b, Fq2& r0, Fq2& r1) {
  const Fq2 a2 = f2sqr(a), b2 = f2sqr(b);
  r1 = f2sqr(a + b) - a2 - b2;
  r0 = a2 + xi(b2);
}
// F <- F^2 for F in the cyclotomic subgroup (Granger-Scott, 9 Fq2 squarings; arrangement of tower.cuh).
// w-basis coefficient g_i sits at slot F + (i even ? i / 2 : 3 + i / 2).
template <class M>
KB_ST_CALL void cycsqr(M m) {
  Fq2 A0, A1, B0, B1, C0, C1;
  {
    const Fq2 g0 = m.ld(F), g3 = m.ld(F + 4);
    fq4sqr(g0, g3, A0, A1);
    m.st(F, dbl(A0 - g0) + A0);
    m.st(F + 4, dbl(A1 + g3) + A1);
  }
  fq4sqr(m.ld(F + 3), m.ld(F + 2), B0, B1);   // (g1, g4)
  fq4sqr(m.ld(F + 1), m.ld(F + 5), C0, C1);   // (g2, g5)
  const Fq2 sC0 = xi(C1);
  const Fq2 h1 = dbl(sC0 + m.ld(F + 3)) + sC0, h4 = dbl(C0 - m.ld(F + 2)) + C0;
  const Fq2 h2 = dbl(B0 - m.ld(F + 1)) + B0, h5 = dbl(B1 + m.ld(F + 5)) + B1;
  m.st(F + 3, h1); m.st(F + 2, h4); m.st(F + 1, h2); m.st(F + 5, h5);
}
// F <- F * (l0 + l1 w + l3 w^3) with (l0, l1, l3) at S, S + 1, S + 2: the line is c0 = (l0, 0, 0), c1 = (l1, l3, 0) in the
// tower; 13 Fq2 products (3 + 5 + 5).  S + 3..5 is scratch.
template <class M>
KB_ST_CALL void f12mul_line(M m) {
  {
    const Fq2 l0 = m.ld(S);
    for (int i = 0; i < 3; i++) m.st(S + 3 + i, f2mul(m.ld(F + i), l0));   // a = f0 l0
  }
  const F6 b = f6m01(m, F + 3, m.ld(S + 1), m.ld(S + 2));                  // b = f1 (l1 + l3 v)
  for (int i = 0; i < 3; i++) m.st(F + i, m.ld(F + i) + m.ld(F + 3 + i));   // f0 + f1
  const F6 e = f6m01(m, F, m.ld(S) + m.ld(S + 1), m.ld(S + 2));
  m.st(F + 3, e.c0 - m.ld(S + 3) - b.c0);                                   // c1 = e - a - b
  m.st(F + 4, e.c1 - m.ld(S + 4) - b.c1);
  m.st(F + 5, e.c2 - m.ld(S + 5) - b.c2);
  m.st(F, m.ld(S + 3) + xi(b.c2));                                          // c0 = a + v b
  m.st(F + 1, m.ld(S + 4) + b.c0);
  m.st(F + 2, m.ld(S + 5) + b.c1);
}
// F <- F^(q^k), k = 1..3: w-basis coefficient i -> conj^k(g_i) * gamma_{k,i};  gamma = FROB_GAMMA table (consts_gen.cuh)
template <class M>
KB_ST_CALL void f12frob(M m, int k, const uint32_t* gamma) {
  for (int s = 0; s < 6; s++) {
    const int i = s < 3 ? 2 * s : 2 * (s - 3) + 1;
    Fq2 x = m.ld(F + s);
    if (k & 1) x = conj(x);
    if (i != 0) {
      const uint32_t* g = gamma + ((k - 1) * 6 + i) * 16;
      const Fq2 c = ld_const2(g);
      x = (k == 2) ? f2mulfq(x, c.c0) : f2mul(x, c);   // gamma_{2,i} lies in Fq
    }
    m.st(F + s, x);
  }
}
template <class M> KB_ST_INL void f12conj(M& m) { for (int i = 3; i < 6; i++) m.st(F + i, -m.ld(F + i)); }
template <class M> KB_ST_INL void f12copy(M& m, int d, int a) { for (int i = 0; i < 6; i++) m.st(d + i, m.ld(a + i)); }

KB_ST_INL Fq2 f2inv(const Fq2& a) {
  const Fq d = inv(f1mul(a.c0, a.c0) + f1mul(a.c1, a.c1));
  Fq2 r; r.c0 = f1mul(a.c0, d); r.c1 = -f1mul(a.c1, d); return r;
}
// F <- 1 / F: d = 1 / (a0^2 - v a1^2) in Fq6, result (a0 d, -a1 d)
template <class M>
KB_ST_CALL void f12inv(M m) {
  f6mul(m, S, F, F);
  f6mul(m, S + 3, F + 3, F + 3);
  const Fq2 d0 = m.ld(S) - xi(m.ld(S + 5)), d1 = m.ld(S + 1) - m.ld(S + 3), d2 = m.ld(S + 2) - m.ld(S + 4);
  const Fq2 t0 = f2sqr(d0) - xi(f2mul(d1, d2));
  const Fq2 t1 = xi(f2sqr(d2)) - f2mul(d0, d1);
  const Fq2 t2 = f2sqr(d1) - f2mul(d0, d2);
  const Fq2 n = f2inv(f2mul(d0, t0) + xi(f2mul(d2, t1) + f2mul(d1, t2)));
  m.st(S + 3, f2mul(t0, n)); m.st(S + 4, f2mul(t1, n)); m.st(S + 5, f2mul(t2, n));
  f6mul(m, F, F, S + 3);
  f6mul(m, F + 3, F + 3, S + 3);
  for (int i = 3; i < 6; i++) m.st(F + i, -m.ld(F + i));
}

// ------------------------------------------------------------------------------------------ Miller loop
// Tangent at T, T <- 2T (formulas of pairing.cuh line_dbl); leaves (l0 yP, l1 xP, l3) at S..S+2.
template <class M>
KB_ST_CALL void line_dbl(M m) {
  const Fq2 X = m.ld(G_TX), Y = m.ld(G_TY), Z = m.ld(G_TZ);
  const Fq2 a = f2sqr(X), b = f2sqr(Y), c = f2sqr(b);
  const Fq2 d = dbl(f2sqr(X + b) - a - c);
  const Fq2 e = tpl(a);
  const Fq2 zz = f2sqr(Z);
  const Fq2 z3 = dbl(f2mul(Y, Z));
  const Fq2 P = m.ld(G_P);
  m.st(S + 2, f2mul(e, X) - dbl(b));
  m.st(S + 1, f2mulfq(-f2mul(e, zz), P.c0));
  m.st(S, f2mulfq(f2mul(z3, zz), P.c1));
  const Fq2 x3 = f2sqr(e) - dbl(d);
  m.st(G_TY, f2mul(e, d - x3) - dbl(dbl(dbl(c))));
  m.st(G_TX, x3);
  m.st(G_TZ, z3);
}
// Chord through T and the affine point (x2, y2), T <- T + (x2, y2) (pairing.cuh line_add)
template <class M>
KB_ST_CALL void line_add(M m, Fq2 x2, Fq2 y2) {
  const Fq2 X = m.ld(G_TX), Y = m.ld(G_TY), Z = m.ld(G_TZ);
  const Fq2 zz = f2sqr(Z);
  const Fq2 h = f2mul(x2, zz) - X;
  const Fq2 r = f2mul(y2, f2mul(Z, zz)) - Y;
  const Fq2 z3 = f2mul(Z, h);
  const Fq2 P = m.ld(G_P);
  m.st(S + 2, f2mul(r, x2) - f2mul(z3, y2));
  m.st(S + 1, f2mulfq(-r, P.c0));
  m.st(S, f2mulfq(z3, P.c1));
  const Fq2 h2 = f2sqr(h), h3 = f2mul(h, h2), v = f2mul(X, h2);
  const Fq2 x3 = f2sqr(r) - h3 - dbl(v);
  m.st(G_TY, f2mul(r, v - x3) - f2mul(Y, h3));
  m.st(G_TX, x3);
  m.st(G_TZ, z3);
}

// The pairing as a RESUMABLE sequence of steps: everything that is alive between two steps is in the memory object (F
// on chip, the G2 accumulator, P, Q and the saved powers in the scratch), so any range of steps can run in one kernel
// launch and the next range in another (pairing_st.cu packs ranges of equal cost into rounds of resident warps).
//
// Miller loop F <- f_{6z+2,Q}(P) l_{pi(Q)} l_{-pi^2(Q)} for P at G_P (x, y packed as one Fq2), Q at G_QX, G_QY (both
// finite): steps 0..63 are the loop iterations i = 63 - step (step 0 initialises F and T), step 64 the two Frobenius
// additions.  tw = TW_X || TW_Y (consts_gen.cuh).
static constexpr int MILLER_STEPS = 65;
template <class M>
KB_ST_INL void miller_steps(M& m, const uint32_t* tw, int lo, int hi) {
  for (int s = lo; s < hi; s++) {
    if (s == 0) {
      m.st(F, Fq2::one());
      for (int i = 1; i < 6; i++) m.st(F + i, Fq2::zero());
      m.st(G_TX, m.ld(G_QX)); m.st(G_TY, m.ld(G_QY)); m.st(G_TZ, Fq2::one());
    }
    if (s < 64) {
      const int i = 63 - s;
      if (i != 63) f12sqr(m);
      line_dbl(m);
      f12mul_line(m);
      const int d = ate_digit(i);
      if (d != 0) {
        const Fq2 qy = m.ld(G_QY);
        line_add(m, m.ld(G_QX), d > 0 ? qy : -qy);
        f12mul_line(m);
      }
    } else {
      const Fq2 twx = ld_const2(tw), twy = ld_const2(tw + 16);
      const Fq2 q1x = f2mul(conj(m.ld(G_QX)), twx), q1y = f2mul(conj(m.ld(G_QY)), twy);
      line_add(m, q1x, q1y);
      f12mul_line(m);
      const Fq2 q2x = f2mul(conj(q1x), twx), q2y = -f2mul(conj(q1y), twy);
      line_add(m, q2x, q2y);
      f12mul_line(m);
    }
  }
}
template <class M>
KB_ST_INL void miller(M& m, const uint32_t* tw) { miller_steps(m, tw, 0, MILLER_STEPS); }

// ------------------------------------------------------------------------------------------ final exponentiation
// width-4 wNAF of z (LSB first; 14 non-zero digits in {+-1, +-3, +-5, +-7}; conjugation is the free inverse)
KB_HD int z_wnaf(int i) {
  constexpr signed char d[63] = {1, 0, 0, 0, -1, 0, 0, 0, 0, 5, 0, 0, 0, 0, 0, 0, -7, 0, 0, 0, 7, 0, 0, 0, 0, 5, 0, 0, 0, 0, 1, 0,
                                 0, 0, -3, 0, 0, 0, -5, 0, 0, 0, 5, 0, 0, 0, 0, 3, 0, 0, 0, -3, 0, 0, 0, 0, 5, 0, 0, 0, 0, 0, 1};
  return d[i];
}
// F <- F^z (F cyclotomic) in 63 steps: step 0 builds the table a, a^3, a^5, a^7 at scratch Fq12 slots tb..tb+3 (a^2 at
// tb+4) and reloads a; step k = 1..62 is the square(-and-multiply) for digit i = 62 - k
template <class M>
KB_ST_INL void cyc_exp_z_step(M& m, int tb, int k) {
  if (k == 0) {
    f12copy(m, G12(tb), F);
    cycsqr(m);
    f12copy(m, G12(tb + 4), F);
    for (int j = 1; j < 4; j++) {
      f12mul(m, j == 1 ? G12(tb) : G12(tb + 4), 0);   // a^2 a, then (a^(2j-1)) a^2
      f12copy(m, G12(tb + j), F);
    }
    f12copy(m, F, G12(tb));
    return;
  }
  cycsqr(m);
  const int d = z_wnaf(62 - k);
  if (d != 0) f12mul(m, G12(tb + ((d < 0 ? -d : d) >> 1)), d < 0 ? 1 : 0);
}

// F <- F^((q^6 - 1)(q^2 + 1) lambda): easy part, then the y0..y16 arrangement of the Fuentes-Castaneda hard part that
// arkworks uses (pairing.cuh final_exponentiation), on one on-chip accumulator with saved values in the scratch.
// Steps: 0 easy part; 1..63 r^z; 64 glue; 65..127 y3^z; 128 glue; 129..191 y5^z; 192 the closing products.
static constexpr int FE_STEPS = 193;
template <class M>
KB_ST_INL void final_exp_steps(M& m, const uint32_t* gamma, int lo, int hi) {
  enum { K_T = 0, K_R = 1, K_Y1 = 2, K_Y3 = 3, K_Y4 = 4, K_Y8 = 5, K_Y9 = 6, K_Y11 = 7, K_Y13 = 8, K_Y14 = 9, K_TAB = 10 };
  for (int s = lo; s < hi; s++) {
    if (s == 0) {
      f12copy(m, G12(K_T), F);
      f12inv(m);
      f12mul(m, G12(K_T), 1);              // conj(f) / f
      f12copy(m, G12(K_T), F);
      f12frob(m, 2, gamma);
      f12mul(m, G12(K_T), 0);              // r: cyclotomic from here on
      f12copy(m, G12(K_R), F);
    } else if (s < 64) {
      cyc_exp_z_step(m, K_TAB, s - 1);     // ... y0 = r^-z after the conjugation below
    } else if (s == 64) {
      f12conj(m);                          // y0
      cycsqr(m); f12copy(m, G12(K_Y1), F); // y1
      cycsqr(m);                           // y2
      f12mul(m, G12(K_Y1), 0); f12copy(m, G12(K_Y3), F);   // y3 = y2 y1
    } else if (s < 128) {
      cyc_exp_z_step(m, K_TAB, s - 65);
    } else if (s == 128) {
      f12conj(m); f12copy(m, G12(K_Y4), F);   // y4 = y3^-z
      cycsqr(m);                           // y5
    } else if (s < 192) {
      cyc_exp_z_step(m, K_TAB, s - 129);   // y6 = y5^z
    } else {
      f12mul(m, G12(K_Y4), 0);             // y7 = y6 y4
      f12mul(m, G12(K_Y3), 1); f12copy(m, G12(K_Y8), F);   // y8 = y7 conj(y3)
      f12mul(m, G12(K_Y1), 0); f12copy(m, G12(K_Y9), F);   // y9 = y8 y1
      f12copy(m, F, G12(K_Y8));
      f12mul(m, G12(K_Y4), 0);             // y10 = y8 y4
      f12mul(m, G12(K_R), 0); f12copy(m, G12(K_Y11), F);   // y11 = y10 r
      f12copy(m, F, G12(K_Y9));
      f12frob(m, 1, gamma);                // y12
      f12mul(m, G12(K_Y11), 0); f12copy(m, G12(K_Y13), F); // y13 = y12 y11
      f12copy(m, F, G12(K_Y8));
      f12frob(m, 2, gamma);
      f12mul(m, G12(K_Y13), 0); f12copy(m, G12(K_Y14), F); // y14 = y8^(q^2) y13
      f12copy(m, F, G12(K_Y9));
      f12mul(m, G12(K_R), 1);              // conj(r) y9
      f12frob(m, 3, gamma);                // y15
      f12mul(m, G12(K_Y14), 0);            // y15 y14
    }
  }
}
template <class M>
KB_ST_INL void final_exp(M& m, const uint32_t* gamma) { final_exp_steps(m, gamma, 0, FE_STEPS); }

// canonical (non-Montgomery) little-endian words of F in ark-serialize order (src/kem.rs:31-32,60-61)
template <class M>
KB_ST_INL void gt_words(M& m, uint32_t w[96]) {
  Fq unit = Fq::zero(); unit.v[0] = 1;
  for (int s = 0; s < 6; s++) {
    const Fq2 x = m.ld(F + s);
    const Fq lo = f1mul(x.c0, unit), hi = f1mul(x.c1, unit);
    for (int k = 0; k < 8; k++) { w[16 * s + k] = lo.v[k]; w[16 * s + 8 + k] = hi.v[k]; }
  }
}

}  // namespace st
}  // namespace kb
