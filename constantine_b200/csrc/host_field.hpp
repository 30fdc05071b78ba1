// Host-side (CPU) field / point arithmetic used by the product's serial tail.
//
// The device produces one bucket-sum point per window; combining W <= ~40 window sums is an inherently serial
// chain of ~bits doublings (r = 2^c * r + S_w).  One GPU thread needs ~4 us per doubling, a host core ~0.3 us,
// so this tail -- and only this tail -- runs on the calling host thread (see DESIGN.md "serial tail").
// It is part of the product, not the oracle: there is no alternative CPU path for the bucket work.
//
// Follows the same value conventions as the reference (Montgomery residues, R = 2^(64*limbs);
// reference constantine/math/arithmetic/limbs_montgomery.nim:180-217 for CIOS), written for
// 64-bit limbs with unsigned __int128.
#pragma once
#include <cstdint>
#include <cstring>
#include "field_constants.cuh"

namespace b200 {
namespace host {

typedef unsigned __int128 u128;

template <class F>
struct HFp {
  static constexpr int N = F::N64;
  uint64_t l[N];

  static HFp zero() { HFp r; for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
  static HFp one() { HFp r; for (int i = 0; i < N; i++) r.l[i] = F::ONE64(i); return r; }
  bool is_zero() const { uint64_t o = 0; for (int i = 0; i < N; i++) o |= l[i]; return o == 0; }
  bool operator==(const HFp& b) const { uint64_t o = 0; for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i]; return o == 0; }

  static bool geq_p(const uint64_t* a) {
    for (int i = N - 1; i >= 0; i--) {
      if (a[i] > F::P64(i)) return true;
      if (a[i] < F::P64(i)) return false;
    }
    return true;
  }
  static void sub_p(uint64_t* a) {
    u128 borrow = 0;
    for (int i = 0; i < N; i++) {
      u128 d = (u128)a[i] - F::P64(i) - borrow;
      a[i] = (uint64_t)d;
      borrow = (d >> 64) & 1;
    }
  }
  HFp operator+(const HFp& b) const {
    HFp r; u128 c = 0;
    for (int i = 0; i < N; i++) { c += (u128)l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (c || geq_p(r.l)) sub_p(r.l);
    return r;
  }
  HFp operator-(const HFp& b) const {
    HFp r; u128 borrow = 0;
    for (int i = 0; i < N; i++) {
      u128 d = (u128)l[i] - b.l[i] - borrow;
      r.l[i] = (uint64_t)d;
      borrow = (d >> 64) & 1;
    }
    if (borrow) {
      u128 c = 0;
      for (int i = 0; i < N; i++) { c += (u128)r.l[i] + F::P64(i); r.l[i] = (uint64_t)c; c >>= 64; }
    }
    return r;
  }
  HFp operator*(const HFp& b) const {
    uint64_t t[N + 2];
    for (int i = 0; i < N + 2; i++) t[i] = 0;
    for (int i = 0; i < N; i++) {
      u128 c = 0;
      for (int j = 0; j < N; j++) {
        c += (u128)l[j] * b.l[i] + t[j];
        t[j] = (uint64_t)c; c >>= 64;
      }
      c += t[N]; t[N] = (uint64_t)c; t[N + 1] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * F::INV64;
      c = (u128)m * F::P64(0) + t[0];
      c >>= 64;
      for (int j = 1; j < N; j++) {
        c += (u128)m * F::P64(j) + t[j];
        t[j - 1] = (uint64_t)c; c >>= 64;
      }
      c += t[N]; t[N - 1] = (uint64_t)c; c >>= 64;
      t[N] = t[N + 1] + (uint64_t)c;
    }
    HFp r;
    for (int i = 0; i < N; i++) r.l[i] = t[i];
    if (t[N] || geq_p(r.l)) sub_p(r.l);
    return r;
  }
  HFp sqr() const { return (*this) * (*this); }
  HFp dbl() const { return (*this) + (*this); }
  HFp neg() const { if (is_zero()) return *this; return zero() - *this; }
  // a^(p-2) (Fermat); only used off the hot path (test/bench input generation, affine normalisation helpers)
  HFp inv() const {
    uint64_t e[N];
    for (int i = 0; i < N; i++) e[i] = F::P64(i);
    e[0] -= 2;  // p is odd and > 2: no borrow
    HFp r = one(), b = *this;
    for (int i = 0; i < 64 * N; i++) {
      if ((e[i >> 6] >> (i & 63)) & 1) r = r * b;
      b = b.sqr();
    }
    return r;
  }
};

template <class F>
struct HFp2 {
  typedef HFp<F> Base;
  Base c0, c1;
  static HFp2 zero() { HFp2 r; r.c0 = Base::zero(); r.c1 = Base::zero(); return r; }
  static HFp2 one() { HFp2 r; r.c0 = Base::one(); r.c1 = Base::zero(); return r; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const HFp2& b) const { return c0 == b.c0 && c1 == b.c1; }
  HFp2 operator+(const HFp2& b) const { HFp2 r; r.c0 = c0 + b.c0; r.c1 = c1 + b.c1; return r; }
  HFp2 operator-(const HFp2& b) const { HFp2 r; r.c0 = c0 - b.c0; r.c1 = c1 - b.c1; return r; }
  HFp2 operator*(const HFp2& b) const {
    Base v0 = c0 * b.c0, v1 = c1 * b.c1, s = (c0 + c1) * (b.c0 + b.c1);
    HFp2 r; r.c0 = v0 - v1; r.c1 = (s - v0) - v1; return r;
  }
  HFp2 sqr() const { Base t = c0 * c1; HFp2 r; r.c0 = (c0 + c1) * (c0 - c1); r.c1 = t + t; return r; }
  HFp2 dbl() const { HFp2 r; r.c0 = c0.dbl(); r.c1 = c1.dbl(); return r; }
  HFp2 neg() const { HFp2 r; r.c0 = c0.neg(); r.c1 = c1.neg(); return r; }
  HFp2 inv() const {
    Base n = (c0.sqr() + c1.sqr()).inv();
    HFp2 r; r.c0 = c0 * n; r.c1 = (c1 * n).neg(); return r;
  }
};

// Extended-Jacobian ("XYZZ") point: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity iff ZZ == 0.
// Same coordinate system the device buckets use (formulas: EFD shortw "xyzz", add-2008-s / dbl-2008-s-1, a = 0;
// the reference carries the same system in constantine/math/elliptic/ec_shortweierstrass_jacobian_extended.nim:30-40).
template <class T>
struct HXyzz {
  T x, y, zz, zzz;
  static HXyzz inf() { HXyzz r; r.x = T::zero(); r.y = T::zero(); r.zz = T::zero(); r.zzz = T::zero(); return r; }
  bool is_inf() const { return zz.is_zero(); }
};

template <class T>
inline HXyzz<T> xyzz_dbl(const HXyzz<T>& p) {
  if (p.is_inf()) return p;
  T U = p.y.dbl();
  T V = U.sqr();
  T W = U * V;
  T S = p.x * V;
  T X2 = p.x.sqr();
  T M = X2.dbl() + X2;
  HXyzz<T> r;
  r.x = M.sqr() - S.dbl();
  r.y = M * (S - r.x) - W * p.y;
  r.zz = V * p.zz;
  r.zzz = W * p.zzz;
  return r;
}

template <class T>
inline HXyzz<T> xyzz_add(const HXyzz<T>& p, const HXyzz<T>& q) {
  if (p.is_inf()) return q;
  if (q.is_inf()) return p;
  T U1 = p.x * q.zz, U2 = q.x * p.zz;
  T S1 = p.y * q.zzz, S2 = q.y * p.zzz;
  T P = U2 - U1, R = S2 - S1;
  if (P.is_zero()) {
    if (R.is_zero()) return xyzz_dbl(p);
    return HXyzz<T>::inf();
  }
  T PP = P.sqr();
  T PPP = P * PP;
  T Q = U1 * PP;
  HXyzz<T> r;
  r.x = R.sqr() - PPP - Q.dbl();
  r.y = R * (Q - r.x) - S1 * PPP;
  r.zz = p.zz * q.zz * PP;
  r.zzz = p.zzz * q.zzz * PPP;
  return r;
}

// XYZZ -> Jacobian (X, Y, Z) with x = X/Z^2, y = Y/Z^3, no inversion:  Z := ZZ*ZZZ  (Z^2 = ZZ^2 ZZZ^2, Z^3 = ZZ^3 ZZZ^3)
//   X = x Z^2 = X1 * ZZ * ZZZ^2 ;  Y = y Z^3 = Y1 * ZZ^3 * ZZZ^2.
// Infinity is written as (1, 1, 0) like the reference does (ec_shortweierstrass_jacobian.nim:46-63).
template <class T>
inline void xyzz_to_jac(const HXyzz<T>& p, T& X, T& Y, T& Z) {
  if (p.is_inf()) { X = T::one(); Y = T::one(); Z = T::zero(); return; }
  T zzz2 = p.zzz.sqr();
  T zz2 = p.zz.sqr();
  X = p.x * p.zz * zzz2;
  Y = p.y * zz2 * p.zz * zzz2;
  Z = p.zz * p.zzz;
}

// XYZZ -> homogeneous projective (X, Y, Z) with x = X/Z, y = Y/Z:  Z := ZZ*ZZZ, X = X1*ZZZ, Y = Y1*ZZ.
// Infinity is (0, 1, 0) (reference ec_shortweierstrass_projective.nim:46-62).
template <class T>
inline void xyzz_to_prj(const HXyzz<T>& p, T& X, T& Y, T& Z) {
  if (p.is_inf()) { X = T::zero(); Y = T::one(); Z = T::zero(); return; }
  X = p.x * p.zzz;
  Y = p.y * p.zz;
  Z = p.zz * p.zzz;
}

}  // namespace host
}  // namespace b200
