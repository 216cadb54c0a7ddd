// Short-Weierstrass (a = 0) group arithmetic, generic over the coordinate field F (Fq for G1,
// Fq2 for G2).  Replaces ark-ec 0.4.2's SW `Projective`/`Affine` group law (Cargo.lock:41-42) at
// src/kzg.rs:57,60,98,135,144,190 and src/kem.rs:22,30,36,37.
//
// Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// a mixed addition costs 8M + 2S and needs no field inversion.  Infinity is ZZ = 0.  Affine
// points use (0, 0) for infinity (not on either curve since b != 0).  Results leave the device
// as canonical affine coordinates, so the coordinate system is free.
#pragma once
#include "tower.cuh"

namespace kb {

template <class F>
struct Affine {
  F x, y;
  static KB_HD Affine infinity() { Affine p; p.x = F::zero(); p.y = F::zero(); return p; }
  KB_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

template <class F>
struct XYZZ {
  F x, y, zz, zzz;
  static KB_HD XYZZ infinity() { XYZZ p; p.x = F::one(); p.y = F::one(); p.zz = F::zero(); p.zzz = F::zero(); return p; }
  KB_HD bool is_inf() const { return zz.is_zero(); }
};

template <class F>
KB_HD XYZZ<F> to_xyzz(const Affine<F>& a) {
  if (a.is_inf()) return XYZZ<F>::infinity();
  XYZZ<F> p; p.x = a.x; p.y = a.y; p.zz = F::one(); p.zzz = F::one(); return p;
}

template <class F>
KB_HD Affine<F> neg(const Affine<F>& a) { Affine<F> r; r.x = a.x; r.y = -a.y; return r; }
template <class F>
KB_HD XYZZ<F> neg(const XYZZ<F>& a) { XYZZ<F> r = a; r.y = -a.y; return r; }

// dbl-2008-s-1
template <class F>
KB_FN XYZZ<F> ec_dbl(const XYZZ<F>& p) {
  if (p.is_inf()) return p;
  F u = dbl(p.y), v = sqr(u), w = u * v, s = p.x * v;
  F x2 = sqr(p.x);
  F m = dbl(x2) + x2;
  XYZZ<F> r;
  r.x = sqr(m) - dbl(s);
  r.y = m * (s - r.x) - w * p.y;
  r.zz = v * p.zz;
  r.zzz = w * p.zzz;
  return r;
}

// doubling of an affine point (mdbl-2008-s-1)
template <class F>
KB_FN XYZZ<F> ec_dbl_affine(const Affine<F>& p) {
  if (p.is_inf()) return XYZZ<F>::infinity();
  F u = dbl(p.y), v = sqr(u), w = u * v, s = p.x * v;
  F x2 = sqr(p.x);
  F m = dbl(x2) + x2;
  XYZZ<F> r;
  r.x = sqr(m) - dbl(s);
  r.y = m * (s - r.x) - w * p.y;
  r.zz = v;
  r.zzz = w;
  return r;
}

// madd-2008-s with the exceptional cases (P = +-Q, either operand at infinity) handled.
template <class F>
KB_FN XYZZ<F> ec_add_mixed(const XYZZ<F>& p, const Affine<F>& q) {
  if (q.is_inf()) return p;
  if (p.is_inf()) return to_xyzz(q);
  F u2 = q.x * p.zz, s2 = q.y * p.zzz;
  F pp_ = u2 - p.x, r_ = s2 - p.y;
  if (pp_.is_zero()) {
    if (r_.is_zero()) return ec_dbl_affine(q);
    return XYZZ<F>::infinity();
  }
  F pp = sqr(pp_), ppp = pp_ * pp, qv = p.x * pp;
  XYZZ<F> r;
  r.x = sqr(r_) - ppp - dbl(qv);
  r.y = r_ * (qv - r.x) - p.y * ppp;
  r.zz = p.zz * pp;
  r.zzz = p.zzz * ppp;
  return r;
}

// add-2008-s
template <class F>
KB_FN XYZZ<F> ec_add(const XYZZ<F>& p, const XYZZ<F>& q) {
  if (q.is_inf()) return p;
  if (p.is_inf()) return q;
  F u1 = p.x * q.zz, u2 = q.x * p.zz, s1 = p.y * q.zzz, s2 = q.y * p.zzz;
  F pp_ = u2 - u1, r_ = s2 - s1;
  if (pp_.is_zero()) {
    if (r_.is_zero()) return ec_dbl(p);
    return XYZZ<F>::infinity();
  }
  F pp = sqr(pp_), ppp = pp_ * pp, qv = u1 * pp;
  XYZZ<F> r;
  r.x = sqr(r_) - ppp - dbl(qv);
  r.y = r_ * (qv - r.x) - s1 * ppp;
  r.zz = p.zz * q.zz * pp;
  r.zzz = p.zzz * q.zzz * ppp;
  return r;
}

// Affine normalisation with a single inversion.
template <class F>
KB_FN Affine<F> to_affine(const XYZZ<F>& p) {
  if (p.is_inf()) return Affine<F>::infinity();
  F i = inv(p.zz * p.zzz);
  Affine<F> r;
  r.x = p.x * (i * p.zzz);  // X / ZZ
  r.y = p.y * (i * p.zz);   // Y / ZZZ
  return r;
}

// Equality of the points represented (cross-multiplied, like arkworks' Projective::eq)
template <class F>
KB_HD bool ec_eq(const XYZZ<F>& p, const XYZZ<F>& q) {
  if (p.is_inf() || q.is_inf()) return p.is_inf() && q.is_inf();
  return p.x * q.zz == q.x * p.zz && p.y * q.zzz == q.y * p.zzz;
}

// k * P for a canonical (non-Montgomery) 256-bit scalar, 4-bit fixed windows, MSB first.
template <class F>
KB_HD_NOINLINE XYZZ<F> ec_mul(const XYZZ<F>& p, const uint32_t k[8]) {
  XYZZ<F> tab[16];
  tab[0] = XYZZ<F>::infinity();
  tab[1] = p;
  for (int i = 2; i < 16; i++) tab[i] = (i & 1) ? ec_add(tab[i - 1], p) : ec_dbl(tab[i >> 1]);
  XYZZ<F> acc = XYZZ<F>::infinity();
  for (int w = 63; w >= 0; w--) {
    acc = ec_dbl(ec_dbl(ec_dbl(ec_dbl(acc))));
    uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d) acc = ec_add(acc, tab[d]);
  }
  return acc;
}

typedef Affine<Fq> G1Affine;
typedef XYZZ<Fq> G1;
typedef Affine<Fq2> G2Affine;
typedef XYZZ<Fq2> G2;

}  // namespace kb
