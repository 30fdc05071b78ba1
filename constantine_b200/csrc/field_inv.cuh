// Modular inversion for the batched-affine bucket additions: Bernstein-Yang "safegcd" divsteps, per thread.
//
// Replaces on the reference's hot path:
//   inv_vartime -> invmod_vartime        reference constantine/math/arithmetic/finite_fields.nim:386-396,
//                                        constantine/math/arithmetic/limbs_exgcd.nim:708-876
// The reference runs the same family of algorithm (Bernstein-Yang divsteps, transition matrices batched per machine word,
// 62 divsteps per batch on 64-bit words). This is an independent formulation for one 32-bit SIMT lane:
//   * values f, g, d, e live in signed 30-bit limbs (N30 of them, room for the range (-2p, p));
//   * a batch is 30 branch-free divsteps on the low 32 bits of f and g, giving a 2x2 transition matrix with entries
//     |.| <= 2^30; the matrix is applied to (f, g) exactly and to (d, e) modulo p with 64-bit accumulators;
//   * e starts at R^2 mod p, so the result is R^2 * a^-1: for a Montgomery residue a = x R that is x^-1 R, the
//     Montgomery form of the inverse, with no extra multiplication;
//   * the batch loop stops as soon as g = 0 (variable time, like the reference's inv_vartime); INV_BATCHES batches always
//     suffice (half-delta divsteps bound floor((45907 bits + 26313) / 19929)).
// Why it is built this way: every lane of a warp inverts its own batch product at the same time, so the inversion has to be
// cheap in ISSUED instructions and free of lane-dependent branches: ~30 x (30 x 18 + ~350) = 27 k instructions for 381 bits,
// against ~100 k for a limb-wise binary extended Euclid under SIMT divergence.
// Limb-level model with register-width assertions: tests/safegcd_emulation.py (run by the CPU suite).
#pragma once
#include "field.cuh"

namespace b200 {

// bits [30 i, 30 i + 30) of an N-limb little-endian integer
template <int N>
B200_DEV int32_t limb30_of(const uint32_t* a, int i) {
  const int bit = 30 * i, w = bit >> 5, s = bit & 31;
  uint32_t lo = (w < N) ? a[w] : 0u;
  uint32_t hi = (w + 1 < N) ? a[w + 1] : 0u;
  uint32_t v = s ? ((lo >> s) | (hi << (32 - s))) : lo;
  return (int32_t)(v & 0x3FFFFFFFu);
}

template <class F>
__device__ __noinline__ void fe_inv_safegcd(uint32_t* r, const uint32_t* a) {
  constexpr int N = F::N, L = F::N30;
  constexpr int32_t M30 = 0x3FFFFFFF;
  int32_t f[L], g[L], d[L], e[L];
  {
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < N; i++) nz |= a[i];
    if (nz == 0u) {   // 0 has no inverse: return 0 (callers never feed it; the batch kernels skip zero denominators)
#pragma unroll
      for (int i = 0; i < N; i++) r[i] = 0u;
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < L; i++) {
    g[i] = limb30_of<N>(a, i);
    f[i] = F::P30(i);
    d[i] = 0;
    e[i] = F::R2_30(i);
  }
  int32_t zeta = -1;   // -(delta + 1/2)
#pragma unroll 1
  for (int batch = 0; batch < F::INV_BATCHES; batch++) {
    // ---- 30 divsteps on the low words: transition matrix [u v; q rr] (scaled by 2^30)
    int32_t u = 1, v = 0, q = 0, rr = 1;
    uint32_t fl = (uint32_t)f[0] | ((uint32_t)f[1] << 30), gl = (uint32_t)g[0] | ((uint32_t)g[1] << 30);
#pragma unroll 6
    for (int i = 0; i < 30; i++) {
      int32_t c1 = zeta >> 31;                        // all-ones if delta > 0
      const uint32_t x = (fl ^ (uint32_t)c1) - (uint32_t)c1;
      const int32_t y = (u ^ c1) - c1, z = (v ^ c1) - c1;
      const int32_t c2 = -(int32_t)(gl & 1u);         // all-ones if g is odd
      gl += x & (uint32_t)c2; q += y & c2; rr += z & c2;
      c1 &= c2;
      zeta = (zeta ^ c1) - 1;
      fl += gl & (uint32_t)c1; u += q & c1; v += rr & c1;
      gl >>= 1; u <<= 1; v <<= 1;
    }
    // ---- (d, e) <- [u v; q rr] (d, e) / 2^30 mod p   (d, e stay in (-2p, p))
    {
      const int32_t sd = d[L - 1] >> 31, se = e[L - 1] >> 31;
      int32_t md = (u & sd) + (v & se), me = (q & sd) + (rr & se);
      int64_t cd = (int64_t)u * d[0] + (int64_t)v * e[0];
      int64_t ce = (int64_t)q * d[0] + (int64_t)rr * e[0];
      md -= (int32_t)((F::PINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
      me -= (int32_t)((F::PINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
      cd += (int64_t)F::P30(0) * md;
      ce += (int64_t)F::P30(0) * me;
      cd >>= 30; ce >>= 30;
#pragma unroll
      for (int i = 1; i < L; i++) {
        cd += (int64_t)u * d[i] + (int64_t)v * e[i] + (int64_t)F::P30(i) * md;
        ce += (int64_t)q * d[i] + (int64_t)rr * e[i] + (int64_t)F::P30(i) * me;
        d[i - 1] = (int32_t)cd & M30; e[i - 1] = (int32_t)ce & M30;
        cd >>= 30; ce >>= 30;
      }
      d[L - 1] = (int32_t)cd; e[L - 1] = (int32_t)ce;
    }
    // ---- (f, g) <- [u v; q rr] (f, g) / 2^30   (exact)
    {
      int64_t cf = (int64_t)u * f[0] + (int64_t)v * g[0];
      int64_t cg = (int64_t)q * f[0] + (int64_t)rr * g[0];
      cf >>= 30; cg >>= 30;
#pragma unroll
      for (int i = 1; i < L; i++) {
        cf += (int64_t)u * f[i] + (int64_t)v * g[i];
        cg += (int64_t)q * f[i] + (int64_t)rr * g[i];
        f[i - 1] = (int32_t)cf & M30; g[i - 1] = (int32_t)cg & M30;
        cf >>= 30; cg >>= 30;
      }
      f[L - 1] = (int32_t)cf; g[L - 1] = (int32_t)cg;
    }
    int32_t gz = 0;
#pragma unroll
    for (int i = 0; i < L; i++) gz |= g[i];
    if (gz == 0) break;
  }
  // f = +-1 now; the inverse is sign(f) * d, brought into [0, p)
  {
    int32_t add = d[L - 1] >> 31;
    const int32_t neg = f[L - 1] >> 31;
#pragma unroll
    for (int i = 0; i < L; i++) {
      int32_t x = d[i] + (F::P30(i) & add);
      d[i] = (x ^ neg) - neg;
    }
#pragma unroll
    for (int i = 0; i < L - 1; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
    add = d[L - 1] >> 31;
#pragma unroll
    for (int i = 0; i < L; i++) d[i] += F::P30(i) & add;
#pragma unroll
    for (int i = 0; i < L - 1; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
  }
  // 30-bit limbs -> 32-bit limbs
#pragma unroll
  for (int w = 0; w < N; w++) {
    const int bit = 32 * w, j = bit / 30, s = bit % 30;
    uint32_t x = (uint32_t)d[j] >> s;
    if (j + 1 < L) x |= (uint32_t)d[j + 1] << (30 - s);
    if (j + 2 < L && 60 - s < 32) x |= (uint32_t)d[j + 2] << (60 - s);
    r[w] = x;
  }
}

// Montgomery-form inverses: a = x R  ->  x^-1 R.  Zero maps to zero.
template <class F>
B200_DEV Fp<F> fe_inverse(const Fp<F>& a) {
  Fp<F> r;
  fe_inv_safegcd<F>(r.l, a.l);
  return r;
}
// 1 / (a0 + a1 i) = (a0 - a1 i) / (a0^2 + a1^2)     (i^2 = -1; reference extension_fields/towers.nim inv on QuadraticExt)
template <class F>
B200_DEV Fp2<F> fe_inverse(const Fp2<F>& a) {
  Fp<F> n = fe_inverse(a.c0.sqr() + a.c1.sqr());
  Fp2<F> r;
  r.c0 = a.c0 * n;
  r.c1 = (a.c1 * n).neg();
  return r;
}

}  // namespace b200
