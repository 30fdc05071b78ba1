// Short-Weierstrass (a = 0) point arithmetic for the bucket method, generic over the coordinate field
// (Fp for G1 / Pasta, Fp2 for G2).
//
// Replaces on the reference's hot path:
//   mixedSum_vartime / sum_vartime / double   reference constantine/math/elliptic/ec_shortweierstrass_jacobian.nim:798-896, 681-796, 564-610
// The reference accumulates buckets in Jacobian (8M+3S mixed add) or batched-affine form; on the GPU every
// bucket lives in extended-Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): the mixed add is
// 8M+2S and never needs a Z that is not already squared/cubed. The reference ships the same system
// (ec_shortweierstrass_jacobian_extended.nim:30-40, 258-310) but does not use it in its MSM. Any coordinate
// system yields the same group element, which is what parity is defined on (affine-normalised equality,
// reference tests/parallel/t_ec_template_parallel.nim:188).
//
// All special cases are exact: infinity operands (affine infinity is (0,0), reference
// ec_shortweierstrass_affine.nim:52-62; XYZZ infinity is ZZ == 0), P + P (doubling), P + (-P).
#pragma once
#include "field.cuh"

namespace b200 {

template <class T>
struct Aff {
  T x, y;
  B200_DEV bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

template <class T>
struct Xyzz {
  T x, y, zz, zzz;
  B200_DEV static Xyzz inf() { Xyzz r; r.x = T::zero(); r.y = T::zero(); r.zz = T::zero(); r.zzz = T::zero(); return r; }
  B200_DEV bool is_inf() const { return zz.is_zero(); }
  B200_DEV static Xyzz from_affine(const Aff<T>& p) {
    Xyzz r;
    if (p.is_inf()) return inf();
    r.x = p.x; r.y = p.y; r.zz = T::one(); r.zzz = T::one();
    return r;
  }
};

// ---- global-memory (de)serialisation: an ABI struct is a flat array of 32-bit words (little-endian limbs) ----
template <class T>
B200_DEV void load_words(T& v, const uint32_t* __restrict__ p) {
  static_assert(T::WORDS % 4 == 0, "16-byte multiple");
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int k = 0; k < T::WORDS / 4; k++) {
    uint4 w = __ldg(q + k);
    v.set_word(4 * k + 0, w.x); v.set_word(4 * k + 1, w.y); v.set_word(4 * k + 2, w.z); v.set_word(4 * k + 3, w.w);
  }
}
template <class T>
B200_DEV void load_words_rw(T& v, const uint32_t* p) {  // plain (coherent) loads for buffers written by earlier kernels
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int k = 0; k < T::WORDS / 4; k++) {
    uint4 w = q[k];
    v.set_word(4 * k + 0, w.x); v.set_word(4 * k + 1, w.y); v.set_word(4 * k + 2, w.z); v.set_word(4 * k + 3, w.w);
  }
}
template <class T>
B200_DEV void store_words(uint32_t* p, const T& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int k = 0; k < T::WORDS / 4; k++) {
    uint4 w;
    w.x = v.word(4 * k + 0); w.y = v.word(4 * k + 1); w.z = v.word(4 * k + 2); w.w = v.word(4 * k + 3);
    q[k] = w;
  }
}

template <class T>
B200_DEV Aff<T> load_affine(const uint32_t* __restrict__ base, uint32_t idx) {
  const uint32_t* p = base + (size_t)idx * (2 * T::WORDS);
  Aff<T> a;
  load_words(a.x, p);
  load_words(a.y, p + T::WORDS);
  return a;
}
template <class T>
B200_DEV Xyzz<T> load_xyzz(const uint32_t* base, size_t idx) {
  const uint32_t* p = base + idx * (4 * T::WORDS);
  Xyzz<T> a;
  load_words_rw(a.x, p);
  load_words_rw(a.y, p + T::WORDS);
  load_words_rw(a.zz, p + 2 * T::WORDS);
  load_words_rw(a.zzz, p + 3 * T::WORDS);
  return a;
}
template <class T>
B200_DEV void store_xyzz(uint32_t* base, size_t idx, const Xyzz<T>& a) {
  uint32_t* p = base + idx * (4 * T::WORDS);
  store_words(p, a.x);
  store_words(p + T::WORDS, a.y);
  store_words(p + 2 * T::WORDS, a.zz);
  store_words(p + 3 * T::WORDS, a.zzz);
}

// ---- group law ------------------------------------------------------------------------------------------

// 2*(x,y) for an affine, finite point: mdbl-2008-s-1 with a = 0.  (y == 0 gives ZZ = 0 = infinity, correct.)
template <class T>
B200_DEV Xyzz<T> xyzz_dbl_affine(const Aff<T>& p) {
  T U = p.y.dbl();
  T V = U.sqr();
  T W = U * V;
  T S = p.x * V;
  T X2 = p.x.sqr();
  T M = X2.dbl() + X2;
  Xyzz<T> r;
  r.x = M.sqr() - S.dbl();
  r.y = M * (S - r.x) - W * p.y;
  r.zz = V;
  r.zzz = W;
  return r;
}

// 2*P, dbl-2008-s-1 with a = 0
template <class T>
B200_DEV Xyzz<T> xyzz_dbl(const Xyzz<T>& p) {
  if constexpr (T::HAS_MUL2) {
    // the same formulas with independent products paired (T::mul2_call): 9 multiplications in 5 dependent steps
    T U = p.y.dbl();
    T V = U * U;
    auto ws = T::mul2_call(U, V, p.x, V);                 // W = U V, S = X V
    auto xz = T::mul2_call(p.x, p.x, V, p.zz);            // X^2, ZZ3 = V ZZ
    T M = xz.a.dbl() + xz.a;
    auto mz = T::mul2_call(M, M, ws.a, p.zzz);            // M^2, ZZZ3 = W ZZZ
    Xyzz<T> r;
    r.x = mz.a - ws.b.dbl();
    auto yy = T::mul2_call(M, ws.b - r.x, ws.a, p.y);     // M (S - X3), W Y
    r.y = yy.a - yy.b;
    r.zz = xz.b;
    r.zzz = mz.b;
    return r;
  }
  T U = p.y.dbl();
  T V = U.sqr();
  T W = U * V;
  T S = p.x * V;
  T X2 = p.x.sqr();
  T M = X2.dbl() + X2;
  Xyzz<T> r;
  r.x = M.sqr() - S.dbl();
  r.y = M * (S - r.x) - W * p.y;
  r.zz = V * p.zz;     // infinity (ZZ = 0) stays infinity
  r.zzz = W * p.zzz;
  return r;
}

// Out-of-line copies for the rare branches of the hot kernel and for every kernel that is not the hot one
// (keeps ptxas time and code size bounded: each is compiled once per coordinate type instead of once per call site).
template <class T>
__device__ __noinline__ void xyzz_dbl_affine_ni(Xyzz<T>& r, const Aff<T>& p) { r = xyzz_dbl_affine(p); }
// by-value flavour for the hot kernel: the accumulator stays in registers, only the (rare) call site copies
template <class T>
__device__ __noinline__ Xyzz<T> xyzz_dbl_affine_val(Aff<T> p) { return xyzz_dbl_affine(p); }
template <class T>
__device__ __noinline__ Xyzz<T> xyzz_dbl_val(Xyzz<T> p) { return xyzz_dbl(p); }
template <class T>
__device__ __noinline__ void xyzz_dbl_ni(Xyzz<T>& r) { r = xyzz_dbl(r); }

// acc += q  (q affine, finite or infinity): madd-2008-s, 8M + 2S on the generic path.
template <class T>
B200_DEV void xyzz_madd(Xyzz<T>& acc, const Aff<T>& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) {
    acc.x = q.x; acc.y = q.y; acc.zz = T::one(); acc.zzz = T::one();
    return;
  }
#ifdef B200_MADD_ROLLED
#define mul_u operator*
#define sqr_u sqr
#endif
  T U2 = q.x.mul_u(acc.zz);
  T S2 = q.y.mul_u(acc.zzz);
  T P = U2 - acc.x;
  T R = S2 - acc.y;
  if (P.is_zero()) {
    if (R.is_zero()) {
      if constexpr (T::WORDS <= 12) acc = xyzz_dbl_affine_val(q);   // registers stay registers; the rare call copies
      else xyzz_dbl_affine_ni(acc, q);                              // Fp2: by reference (large by-value frames misbehave)
    }
    else acc = Xyzz<T>::inf();
    return;
  }
  T PP = P.sqr_u();
  T PPP = P.mul_u(PP);
  T Q = acc.x.mul_u(PP);
  T X3 = R.sqr_u() - PPP - Q.dbl();
#if defined(B200_MADD_ROLLED) || defined(B200_NO_DOT2)
  T Y3 = R.mul_u(Q - X3) - acc.y.mul_u(PPP);
#else
  T Y3 = T::dot2_u(R, Q - X3, acc.y.neg(), PPP);   // R (Q - X3) - Y1 PPP with one Montgomery reduction
#endif
  acc.x = X3;
  acc.y = Y3;
  acc.zz = acc.zz.mul_u(PP);
  acc.zzz = acc.zzz.mul_u(PPP);
#ifdef B200_MADD_ROLLED
#undef mul_u
#undef sqr_u
#endif
}

// acc += q  (both XYZZ): add-2008-s, 12M + 2S on the generic path.
template <class T>
B200_DEV void xyzz_add(Xyzz<T>& acc, const Xyzz<T>& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  if constexpr (T::HAS_MUL2) {
    // independent products paired: 14 multiplications in 8 dependent steps
    auto u = T::mul2_call(acc.x, q.zz, q.x, acc.zz);       // U1, U2
    auto sv = T::mul2_call(acc.y, q.zzz, q.y, acc.zzz);    // S1, S2
    T P = u.b - u.a;
    T R = sv.b - sv.a;
    if (P.is_zero()) {
      if (R.is_zero()) acc = xyzz_dbl_val(acc);
      else acc = Xyzz<T>::inf();
      return;
    }
    T PP = P * P;
    auto pq = T::mul2_call(P, PP, u.a, PP);                // PPP, Q
    auto rz = T::mul2_call(R, R, acc.zz, q.zz);            // R^2, ZZ1 ZZ2
    T X3 = rz.a - pq.a - pq.b.dbl();
    auto yz = T::mul2_call(R, pq.b - X3, sv.a, pq.a);      // R (Q - X3), S1 PPP
    auto zz = T::mul2_call(rz.b, PP, acc.zzz, q.zzz);      // ZZ3, ZZZ1 ZZZ2
    acc.x = X3;
    acc.y = yz.a - yz.b;
    acc.zz = zz.a;
    acc.zzz = zz.b * pq.a;
    return;
  }
  T U1 = acc.x * q.zz;
  T U2 = q.x * acc.zz;
  T S1 = acc.y * q.zzz;
  T S2 = q.y * acc.zzz;
  T P = U2 - U1;
  T R = S2 - S1;
  if (P.is_zero()) {
    if (R.is_zero()) {
      if constexpr (T::WORDS <= 12) acc = xyzz_dbl_val(acc);
      else xyzz_dbl_ni(acc);
    }
    else acc = Xyzz<T>::inf();
    return;
  }
  T PP = P.sqr();
  T PPP = P * PP;
  T Q = U1 * PP;
  T X3 = R.sqr() - PPP - Q.dbl();
  T Y3 = R * (Q - X3) - S1 * PPP;
  acc.x = X3;
  acc.y = Y3;
  acc.zz = acc.zz * q.zz * PP;
  acc.zzz = acc.zzz * q.zzz * PPP;
}

// Fully unrolled multipliers (mul_u / sqr_u) for single chains of dependent point operations, where nothing hides the
// latency of the rolled-loop multiplier (k_batch_tail: one thread per MSM runs Horner over its windows).
template <class T>
B200_DEV Xyzz<T> xyzz_dbl_u(const Xyzz<T>& p) {
  T U = p.y.dbl();
  T V = U.sqr_u();
  T W = U.mul_u(V);
  T S = p.x.mul_u(V);
  T X2 = p.x.sqr_u();
  T M = X2.dbl() + X2;
  Xyzz<T> r;
  r.x = M.sqr_u() - S.dbl();
  r.y = M.mul_u(S - r.x) - W.mul_u(p.y);
  r.zz = V.mul_u(p.zz);
  r.zzz = W.mul_u(p.zzz);
  return r;
}
template <class T>
B200_DEV void xyzz_add_u(Xyzz<T>& acc, const Xyzz<T>& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  T U1 = acc.x.mul_u(q.zz);
  T U2 = q.x.mul_u(acc.zz);
  T S1 = acc.y.mul_u(q.zzz);
  T S2 = q.y.mul_u(acc.zzz);
  T P = U2 - U1;
  T R = S2 - S1;
  if (P.is_zero()) {
    if (R.is_zero()) xyzz_dbl_ni(acc);
    else acc = Xyzz<T>::inf();
    return;
  }
  T PP = P.sqr_u();
  T PPP = P.mul_u(PP);
  T Q = U1.mul_u(PP);
  T X3 = R.sqr_u() - PPP - Q.dbl();
  T Y3 = R.mul_u(Q - X3) - S1.mul_u(PPP);
  acc.x = X3;
  acc.y = Y3;
  acc.zz = acc.zz.mul_u(q.zz).mul_u(PP);
  acc.zzz = acc.zzz.mul_u(q.zzz).mul_u(PPP);
}

template <class T>
__device__ __noinline__ void xyzz_add_ni(Xyzz<T>& acc, const Xyzz<T>& q) { xyzz_add(acc, q); }
template <class T>
__device__ __noinline__ void xyzz_madd_ni(Xyzz<T>& acc, const Aff<T>& q) { xyzz_madd(acc, q); }
template <class T>
__device__ __noinline__ void mul_ni(T& r, const T& a, const T& b) { r = a * b; }

}  // namespace b200
