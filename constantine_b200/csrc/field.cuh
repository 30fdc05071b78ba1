// Prime-field and quadratic-extension arithmetic for sm_100a, register-resident, 32-bit limbs.
//
// What this replaces on the reference's hot path:
//   Fp.prod / square -> mulMont -> mulMont_CIOS_sparebit   reference constantine/math/arithmetic/limbs_montgomery.nim:180-217, 484-522
//   Fp.sum / diff / double / neg                            reference constantine/math/arithmetic/finite_fields.nim:172-266
//   Fp2 complex mul / sqr (i^2 = -1)                        reference constantine/math/extension_fields/towers.nim:798-885
//
// Same values as the reference (Montgomery residues a*R mod p with R = 2^(64*limbs64), canonical i.e. always
// fully reduced to [0, p)), different machine mapping: the B200 integer datapath is 32 bits wide
// (IMAD / IMAD.WIDE.U32 on the FMA pipe), so a 64-bit-limb value is processed as 2x as many 32-bit limbs --
// the little-endian byte image is identical, so ABI structs are loaded/stored as-is.
//
// Montgomery multiplication: operand-scanning CIOS where the 32x32->64 partial products are kept in two
// accumulator rows, one holding products that start on even limb positions and one holding products that
// start on odd positions. Inside a row every (lo,hi) product pair lands on an aligned limb pair, so a whole
// row update is ONE carry chain of mad.lo.cc / madc.hi.cc (ptxas pairs them into IMAD.WIDE.U32[.X]).
// Dividing by 2^32 at the end of each outer iteration swaps the roles of the two rows.
#pragma once
#include <cstdint>
#include "field_constants.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------------
// PTX carry-chain primitives. `volatile` keeps program order, which is what keeps the CC flag meaningful
// between consecutive statements (NVVM never emits .cc instructions on its own).
// ---------------------------------------------------------------------------------------------------------
#define B200_DEV __device__ __forceinline__

B200_DEV uint32_t p_add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_DEV uint32_t p_addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_DEV uint32_t p_addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_DEV uint32_t p_sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_DEV uint32_t p_subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_DEV uint32_t p_subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_DEV uint32_t p_mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_DEV uint32_t p_mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_DEV uint32_t p_mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
B200_DEV uint32_t p_madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
B200_DEV uint32_t p_madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
B200_DEV uint32_t p_madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

// ---------------------------------------------------------------------------------------------------------
// Raw limb-array kernels, F = one of the generated field-constant structs (F::N even).
// ---------------------------------------------------------------------------------------------------------

// r = a + b, returns carry-out
template <int N>
B200_DEV uint32_t limbs_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = p_add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = p_addc_cc(a[i], b[i]);
  return p_addc(0, 0);
}

// r = a - b, returns borrow mask (0xFFFFFFFF if a < b else 0)
template <int N>
B200_DEV uint32_t limbs_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = p_sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r[i] = p_subc_cc(a[i], b[i]);
  return p_subc(0, 0);
}

// if (r >= p) r -= p      (r < 2p on entry)
template <class F>
B200_DEV void final_sub(uint32_t* r) {
  constexpr int N = F::N;
  uint32_t t[N];
  t[0] = p_sub_cc(r[0], F::P(0));
#pragma unroll
  for (int i = 1; i < N; i++) t[i] = p_subc_cc(r[i], F::P(i));
  uint32_t borrow = p_subc(0, 0);  // all-ones if r < p
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = borrow ? r[i] : t[i];
}

template <class F>
B200_DEV void fe_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  // a, b < p < 2^(32N - 1): the sum cannot carry out of N limbs (every field here has >= 1 spare bit).
  limbs_add<F::N>(r, a, b);
  final_sub<F>(r);
}

template <class F>
B200_DEV void fe_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = F::N;
  uint32_t borrow = limbs_sub<N>(r, a, b);
  // add back p under the borrow mask
  r[0] = p_add_cc(r[0], F::P(0) & borrow);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r[i] = p_addc_cc(r[i], F::P(i) & borrow);
  r[N - 1] = p_addc(r[N - 1], F::P(N - 1) & borrow);
}

template <class F>
B200_DEV void fe_neg(uint32_t* r, const uint32_t* a) {
  constexpr int N = F::N;
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < N; i++) nz |= a[i];
  uint32_t t[N];
  t[0] = p_sub_cc(F::P(0), a[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) t[i] = p_subc_cc(F::P(i), a[i]);
  t[N - 1] = p_subc(F::P(N - 1), a[N - 1]);
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = nz ? t[i] : 0u;  // -0 = 0 stays canonical
}

// One Montgomery reduction round shared by every outer iteration:
//   m = E[0] * (-p^-1) ; (E,O) += m * p   =>  E[0] becomes 0.
// E holds limbs at positions 0..N-1, O holds limbs at positions 1..N.
template <class F>
B200_DEV void mont_round(uint32_t* E, uint32_t* O) {
  constexpr int N = F::N;
  const uint32_t m = E[0] * F::INV;
  // odd-position products m*p[1], m*p[3], ... : one carry chain over O; the value is bounded so no carry-out
  O[0] = p_mad_lo_cc(m, F::P(1), O[0]);
  O[1] = p_madc_hi_cc(m, F::P(1), O[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    O[j] = p_madc_lo_cc(m, F::P(j + 1), O[j]);
    O[j + 1] = p_madc_hi_cc(m, F::P(j + 1), O[j + 1]);
  }
  // even-position products: one carry chain over E; its carry-out lands on position N = O[N-1]
  E[0] = p_mad_lo_cc(m, F::P(0), E[0]);
  E[1] = p_madc_hi_cc(m, F::P(0), E[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    E[j] = p_madc_lo_cc(m, F::P(j), E[j]);
    E[j + 1] = p_madc_hi_cc(m, F::P(j), E[j + 1]);
  }
  O[N - 1] = p_addc(O[N - 1], 0);
}

// Outer iteration i >= 1. On entry the running value is T = Eold + Oold*2^32 with Eold[0] == 0; the implied
// division by 2^32 makes Oold the new position-0 row (`E`) and Eold >> 64 the new position-1 row (`O`, shifted
// in place by two limbs), with the stray limb Eold[1] folded into E[0].
template <class F>
B200_DEV void mont_step(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
  constexpr int N = F::N;
  E[0] = p_add_cc(E[0], O[1]);  // carry goes to position 1 = head of the O chain below
#pragma unroll
  for (int j = 0; j < N - 2; j += 2) {
    O[j] = p_madc_lo_cc(a[j + 1], bi, O[j + 2]);
    O[j + 1] = p_madc_hi_cc(a[j + 1], bi, O[j + 3]);
  }
  O[N - 2] = p_madc_lo_cc(a[N - 1], bi, 0);
  O[N - 1] = p_madc_hi(a[N - 1], bi, 0);
  E[0] = p_mad_lo_cc(a[0], bi, E[0]);
  E[1] = p_madc_hi_cc(a[0], bi, E[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    E[j] = p_madc_lo_cc(a[j], bi, E[j]);
    E[j + 1] = p_madc_hi_cc(a[j], bi, E[j + 1]);
  }
  O[N - 1] = p_addc(O[N - 1], 0);
  mont_round<F>(E, O);
}

// r = a * b * R^-1 mod p, canonical.  a, b canonical (< p).   r may alias a or b.
template <class F>
B200_DEV void fe_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = F::N;
  static_assert(N % 2 == 0, "even limb count");
  uint32_t A[N], B[N];
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    A[j] = p_mul_lo(a[j], b[0]);
    A[j + 1] = p_mul_hi(a[j], b[0]);
    B[j] = p_mul_lo(a[j + 1], b[0]);
    B[j + 1] = p_mul_hi(a[j + 1], b[0]);
  }
  mont_round<F>(A, B);
#pragma unroll
  for (int i = 1; i < N; i++) {
    if (i & 1)
      mont_step<F>(B, A, a, b[i]);
    else
      mont_step<F>(A, B, a, b[i]);
  }
  // N is even: after the last step (i = N-1, odd) the position-0 row is B, the position-1 row is A... and the
  // final /2^32 gives  result = (row0 >> 32) + row1.
  uint32_t* E = ((N - 1) & 1) ? B : A;
  uint32_t* O = ((N - 1) & 1) ? A : B;
  uint32_t t[N];
  t[0] = p_add_cc(O[0], E[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) t[k] = p_addc_cc(O[k], E[k + 1]);
  t[N - 1] = p_addc(O[N - 1], 0);
  final_sub<F>(t);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = t[k];
}

// (E,O) += c * di  -- a second product row on top of a step's first one (same even/odd carry chains as mont_step).
// No chain can carry out of its top limb: the running value stays below 2^(32(N+1)) (see fe_dot2).
template <class F>
B200_DEV void add_product_row(uint32_t* E, uint32_t* O, const uint32_t* c, uint32_t di) {
  constexpr int N = F::N;
  O[0] = p_mad_lo_cc(c[1], di, O[0]);
  O[1] = p_madc_hi_cc(c[1], di, O[1]);
#pragma unroll
  for (int j = 2; j < N - 2; j += 2) {
    O[j] = p_madc_lo_cc(c[j + 1], di, O[j]);
    O[j + 1] = p_madc_hi_cc(c[j + 1], di, O[j + 1]);
  }
  O[N - 2] = p_madc_lo_cc(c[N - 1], di, O[N - 2]);
  O[N - 1] = p_madc_hi(c[N - 1], di, O[N - 1]);
  E[0] = p_mad_lo_cc(c[0], di, E[0]);
  E[1] = p_madc_hi_cc(c[0], di, E[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    E[j] = p_madc_lo_cc(c[j], di, E[j]);
    E[j + 1] = p_madc_hi_cc(c[j], di, E[j + 1]);
  }
  O[N - 1] = p_addc(O[N - 1], 0);
}

// Outer iteration of the two-product form: same as mont_step up to the reduction round, with the second row in between.
template <class F>
B200_DEV void mont_step2(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, const uint32_t* c, uint32_t di) {
  constexpr int N = F::N;
  E[0] = p_add_cc(E[0], O[1]);
#pragma unroll
  for (int j = 0; j < N - 2; j += 2) {
    O[j] = p_madc_lo_cc(a[j + 1], bi, O[j + 2]);
    O[j + 1] = p_madc_hi_cc(a[j + 1], bi, O[j + 3]);
  }
  O[N - 2] = p_madc_lo_cc(a[N - 1], bi, 0);
  O[N - 1] = p_madc_hi(a[N - 1], bi, 0);
  E[0] = p_mad_lo_cc(a[0], bi, E[0]);
  E[1] = p_madc_hi_cc(a[0], bi, E[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    E[j] = p_madc_lo_cc(a[j], bi, E[j]);
    E[j + 1] = p_madc_hi_cc(a[j], bi, E[j + 1]);
  }
  O[N - 1] = p_addc(O[N - 1], 0);
  add_product_row<F>(E, O, c, di);
  mont_round<F>(E, O);
}

// 3p < 2^(32N)  (sufficient test on the top limb)
template <class F>
__host__ __device__ constexpr bool dot2_fits() { return 3ull * ((unsigned long long)F::P(F::N - 1) + 1ull) <= (1ull << 32); }

// r = (a*b + c*d) * R^-1 mod p, canonical: ONE interleaved Montgomery reduction for two products -- 3N^2 + N
// multiply-accumulates instead of 2(2N^2 + N). Per outer step T <- (T + a*b_i + c*d_i + m*p) / 2^32, so T < 3p
// throughout (inputs < p), which fits N limbs for every field here (3p < 2^(32N): BLS12-381 0.31, BN254 0.57,
// Pasta 0.75 of 2^(32N)) and the value before each division stays below 2^(32(N+1)). Two conditional subtractions finish.
template <class F>
B200_DEV void fe_dot2(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d) {
  constexpr int N = F::N;
  static_assert(N % 2 == 0 && N >= 4, "even limb count");
  static_assert(dot2_fits<F>(), "fe_dot2 needs 3p < 2^(32N)");
  uint32_t A[N], B[N];
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    A[j] = p_mul_lo(a[j], b[0]);
    A[j + 1] = p_mul_hi(a[j], b[0]);
    B[j] = p_mul_lo(a[j + 1], b[0]);
    B[j + 1] = p_mul_hi(a[j + 1], b[0]);
  }
  add_product_row<F>(A, B, c, d[0]);
  mont_round<F>(A, B);
#pragma unroll
  for (int i = 1; i < N; i++) {
    if (i & 1)
      mont_step2<F>(B, A, a, b[i], c, d[i]);
    else
      mont_step2<F>(A, B, a, b[i], c, d[i]);
  }
  uint32_t* E = ((N - 1) & 1) ? B : A;
  uint32_t* O = ((N - 1) & 1) ? A : B;
  uint32_t t[N];
  t[0] = p_add_cc(O[0], E[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) t[k] = p_addc_cc(O[k], E[k + 1]);
  t[N - 1] = p_addc(O[N - 1], 0);
  final_sub<F>(t);
  final_sub<F>(t);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = t[k];
}

// r = a^2 * R^-1 mod p, canonical. Separate-operand-scanning squaring: the N(N-1)/2 off-diagonal products are computed
// once and doubled, the N diagonal squares ride on one carry chain, then N reduction rounds: N(N-1)/2 + N + N^2
// 32x32->64 multiply-accumulates (222 for N = 12) against 2N^2 (288) for the general product.
// Carry discipline (every chain is a run of consecutive limb positions, exactly like the rows of fe_mul):
//   * rows a_i * a_j with j - i odd accumulate in X, with j - i even in Y; a row's carry-out lands on the limb just above
//     its end, which earlier rows (ending no higher) have only ever loaded with carries;
//   * reduction round i adds m*p at positions i..i+N in two chains (even / odd limbs of p); their carry-outs belong to
//     positions >= N and are parked in the side words C (each <= 2) instead of rippling through T.
template <class F>
B200_DEV void fe_sqr(uint32_t* r, const uint32_t* a) {
  constexpr int N = F::N;
  static_assert(N % 2 == 0 && N >= 4, "even limb count");
  uint32_t X[2 * N], Y[2 * N];
#pragma unroll
  for (int k = 0; k < 2 * N; k++) { X[k] = 0; Y[k] = 0; }
#pragma unroll
  for (int i = 0; i < N - 1; i++) {
    // j = i+1, i+3, ...  -> X
    {
#pragma unroll
      for (int j = i + 1; j < N; j += 2) {
        X[i + j] = (j == i + 1) ? p_mad_lo_cc(a[i], a[j], X[i + j]) : p_madc_lo_cc(a[i], a[j], X[i + j]);
        X[i + j + 1] = p_madc_hi_cc(a[i], a[j], X[i + j + 1]);
      }
      const int jl = i + 1 + 2 * ((N - 1 - (i + 1)) / 2);   // last j of this class
      X[i + jl + 2] = p_addc(X[i + jl + 2], 0);
    }
    // j = i+2, i+4, ...  -> Y
    if (i + 2 < N) {
#pragma unroll
      for (int j = i + 2; j < N; j += 2) {
        Y[i + j] = (j == i + 2) ? p_mad_lo_cc(a[i], a[j], Y[i + j]) : p_madc_lo_cc(a[i], a[j], Y[i + j]);
        Y[i + j + 1] = p_madc_hi_cc(a[i], a[j], Y[i + j + 1]);
      }
      const int jl = i + 2 + 2 * ((N - 1 - (i + 2)) / 2);
      Y[i + jl + 2] = p_addc(Y[i + jl + 2], 0);
    }
  }
  // T = 2 (X + Y) + sum_i a_i^2 2^(64 i)
  uint32_t T[2 * N];
  T[0] = p_add_cc(X[0], Y[0]);
#pragma unroll
  for (int k = 1; k < 2 * N - 1; k++) T[k] = p_addc_cc(X[k], Y[k]);
  T[2 * N - 1] = p_addc(X[2 * N - 1], Y[2 * N - 1]);
  T[0] = p_add_cc(T[0], T[0]);
#pragma unroll
  for (int k = 1; k < 2 * N - 1; k++) T[k] = p_addc_cc(T[k], T[k]);
  T[2 * N - 1] = p_addc(T[2 * N - 1], T[2 * N - 1]);
  T[0] = p_mad_lo_cc(a[0], a[0], T[0]);
  T[1] = p_madc_hi_cc(a[0], a[0], T[1]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) {
    T[2 * i] = p_madc_lo_cc(a[i], a[i], T[2 * i]);
    T[2 * i + 1] = p_madc_hi_cc(a[i], a[i], T[2 * i + 1]);
  }
  T[2 * N - 2] = p_madc_lo_cc(a[N - 1], a[N - 1], T[2 * N - 2]);
  T[2 * N - 1] = p_madc_hi(a[N - 1], a[N - 1], T[2 * N - 1]);
  // N reduction rounds
  uint32_t C[N + 1];
#pragma unroll
  for (int k = 0; k <= N; k++) C[k] = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t m = T[i] * F::INV;
    T[i] = p_mad_lo_cc(m, F::P(0), T[i]);
    T[i + 1] = p_madc_hi_cc(m, F::P(0), T[i + 1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      T[i + j] = p_madc_lo_cc(m, F::P(j), T[i + j]);
      T[i + j + 1] = p_madc_hi_cc(m, F::P(j), T[i + j + 1]);
    }
    C[i] = p_addc(C[i], 0);
    T[i + 1] = p_mad_lo_cc(m, F::P(1), T[i + 1]);
    T[i + 2] = p_madc_hi_cc(m, F::P(1), T[i + 2]);
#pragma unroll
    for (int j = 3; j < N; j += 2) {
      T[i + j] = p_madc_lo_cc(m, F::P(j), T[i + j]);
      T[i + j + 1] = p_madc_hi_cc(m, F::P(j), T[i + j + 1]);
    }
    C[i + 1] = p_addc(C[i + 1], 0);
  }
  uint32_t t[N];
  t[0] = p_add_cc(T[N], C[0]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) t[k] = p_addc_cc(T[N + k], C[k]);
  t[N - 1] = p_addc(T[2 * N - 1], C[N - 1]);
  final_sub<F>(t);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = t[k];
}

// Same multiplication with the outer loop kept ROLLED (two CIOS steps per iteration, the multiplier limbs rotated through
// registers so every index stays static): one third of the code of the fully unrolled form. Experiment for the
// instruction-fetch stalls ncu shows in k_accumulate (profiles/ncu_k_accumulate_r1.txt); selected with -DB200_ROLLED_MUL.
template <class F>
B200_DEV void fe_mul_rolled(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = F::N;
  static_assert(N % 2 == 0 && N >= 4, "even limb count");
  uint32_t A[N], B[N];
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    A[j] = p_mul_lo(a[j], b[0]);
    A[j + 1] = p_mul_hi(a[j], b[0]);
    B[j] = p_mul_lo(a[j + 1], b[0]);
    B[j + 1] = p_mul_hi(a[j + 1], b[0]);
  }
  mont_round<F>(A, B);
  uint32_t bb[N];  // bb[0], bb[1] are the next two multiplier limbs
#pragma unroll
  for (int j = 0; j < N - 1; j++) bb[j] = b[j + 1];
  bb[N - 1] = 0;
#pragma unroll 1
  for (int it = 0; it < (N - 2) / 2; it++) {
    mont_step<F>(B, A, a, bb[0]);   // odd step
    mont_step<F>(A, B, a, bb[1]);   // even step
#pragma unroll
    for (int j = 0; j < N - 2; j++) bb[j] = bb[j + 2];
  }
  mont_step<F>(B, A, a, bb[0]);     // last (odd) step, i = N-1
  uint32_t t[N];
  t[0] = p_add_cc(A[0], B[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) t[k] = p_addc_cc(A[k], B[k + 1]);
  t[N - 1] = p_addc(A[N - 1], 0);
  final_sub<F>(t);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = t[k];
}

// Fully uniform rolled form: N/2 iterations of [even step, odd step] starting from zero rows (the first step then
// computes 0 + a*b[0] at the price of a few additions with zero); smallest code.
template <class F>
B200_DEV void fe_mul_rolled2(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = F::N;
  uint32_t A[N], B[N], bb[N];
#pragma unroll
  for (int j = 0; j < N; j++) { A[j] = 0; B[j] = 0; bb[j] = b[j]; }
#pragma unroll 1
  for (int it = 0; it < N / 2; it++) {
    mont_step<F>(A, B, a, bb[0]);   // even step (i = 2 it)
    mont_step<F>(B, A, a, bb[1]);   // odd step  (i = 2 it + 1)
#pragma unroll
    for (int j = 0; j < N - 2; j++) bb[j] = bb[j + 2];
  }
  uint32_t t[N];
  t[0] = p_add_cc(A[0], B[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) t[k] = p_addc_cc(A[k], B[k + 1]);
  t[N - 1] = p_addc(A[N - 1], 0);
  final_sub<F>(t);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = t[k];
}

// ---------------------------------------------------------------------------------------------------------
// Value types with a common interface (zero/one/is_zero/==, +, -, *, sqr, neg, dbl) so that the
// elliptic-curve code is written once for G1 (Fp) and G2 (Fp2).
// ---------------------------------------------------------------------------------------------------------
template <class F>
struct Fp {
  using Params = F;
  static constexpr int N = F::N;
  static constexpr int WORDS = F::N;  // 32-bit words per element
  uint32_t l[N];

  B200_DEV static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = 0;
    return r;
  }
  B200_DEV static Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = F::ONE(i);
    return r;
  }
  B200_DEV bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= l[i];
    return o == 0;
  }
  B200_DEV bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i];
    return o == 0;
  }
  B200_DEV Fp operator+(const Fp& b) const { Fp r; fe_add<F>(r.l, l, b.l); return r; }
  B200_DEV Fp operator-(const Fp& b) const { Fp r; fe_sub<F>(r.l, l, b.l); return r; }
  // The general-purpose product (point additions / doublings of the reduce, fix-up and tail kernels). Inlining the fully
  // unrolled 381-bit multiplier 12-14 times per point operation overflows the instruction caches; a rolled outer loop (one third
  // of the code) avoids that but spends a quarter of its issue slots on IMAD.MOV register rotation (ncu opcode mix,
  // profiles/README.md). 12-limb fields therefore CALL one out-of-line copy of the unrolled multiplier: operands and result
  // travel by value in registers (no stack traffic; ~28 moves per call against ~170 in the rolled loop).
  // B200_MUL_VARIANT forces 0 = unrolled inline, 1 = rolled, 2 = uniform rolled, 3 = out-of-line call.
#ifndef B200_MUL_VARIANT
#define B200_MUL_VARIANT (-1)
#endif
  static __device__ __noinline__ Fp mul_call(Fp a, Fp b) { Fp r; fe_mul<F>(r.l, a.l, b.l); return r; }
  // Two INDEPENDENT products per call: the two unrolled carry-chain programs are interleaved by the scheduler, so a chain of
  // dependent point operations (reduce / fix-up / tail kernels at low occupancy) pays ~1.3 multiplication latencies for two
  // products instead of 2. Point formulas pair their independent products through this (ec.cuh).
  struct Pair { Fp a, b; };
#ifdef B200_NO_MUL2
  static constexpr bool HAS_MUL2 = false;
#else
  static constexpr bool HAS_MUL2 = (N >= 12) && (B200_MUL_VARIANT < 0 || B200_MUL_VARIANT == 3);
#endif
  static __device__ __noinline__ Pair mul2_call(Fp a, Fp b, Fp c, Fp d) {
    Pair r;
    fe_mul<F>(r.a.l, a.l, b.l);
    fe_mul<F>(r.b.l, c.l, d.l);
    return r;
  }
  B200_DEV Fp operator*(const Fp& b) const {
    constexpr int V = (B200_MUL_VARIANT >= 0) ? B200_MUL_VARIANT : (N >= 12 ? 3 : 0);
    if constexpr (V == 3) return mul_call(*this, b);
    Fp r;
    if constexpr (V == 2) fe_mul_rolled2<F>(r.l, l, b.l);
    else if constexpr (V == 1) fe_mul_rolled<F>(r.l, l, b.l);
    else fe_mul<F>(r.l, l, b.l);
    return r;
  }
  B200_DEV Fp sqr() const { return (*this) * (*this); }
  // fully unrolled multiplier regardless of the policy above (the hot mixed add of k_accumulate is faster with it)
  B200_DEV Fp mul_u(const Fp& b) const { Fp r; fe_mul<F>(r.l, l, b.l); return r; }
#ifdef B200_NO_SQR
  B200_DEV Fp sqr_u() const { Fp r; fe_mul<F>(r.l, l, l); return r; }
#else
  B200_DEV Fp sqr_u() const { Fp r; fe_sqr<F>(r.l, l); return r; }
#endif
  // a*b + c*d with one reduction (unrolled)
  static B200_DEV Fp dot2_u(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    if constexpr (dot2_fits<F>()) { Fp r; fe_dot2<F>(r.l, a.l, b.l, c.l, d.l); return r; }
    else return a.mul_u(b) + c.mul_u(d);   // fields without the headroom (BLS12-381 Fr): two reductions
  }
  B200_DEV Fp neg() const { Fp r; fe_neg<F>(r.l, l); return r; }
  B200_DEV Fp dbl() const { Fp r; fe_add<F>(r.l, l, l); return r; }
  // this = cond ? -this : this
  B200_DEV void cneg(bool cond) {
    Fp n = neg();
#pragma unroll
    for (int i = 0; i < N; i++) l[i] = cond ? n.l[i] : l[i];
  }
  // word-wise (de)serialisation used by the point loaders; k in [0, WORDS)
  B200_DEV uint32_t word(int k) const { return l[k]; }
  B200_DEV void set_word(int k, uint32_t v) { l[k] = v; }
  // a^(p-2) (Fermat). Off the hot path only (affine normalisation in the input generator / test hooks).
  __device__ __noinline__ Fp inv() const {
    Fp r = one(), b = *this;
    uint32_t borrow = 2u;  // exponent p - 2, computed word by word with borrow (p's low word may be 1, e.g. Pasta)
    uint32_t w = 0;
#pragma unroll 1
    for (int i = 0; i < 32 * N; i++) {
      if ((i & 31) == 0) {
        uint32_t pw = F::P(i >> 5);
        w = pw - borrow;
        borrow = (pw < borrow) ? 1u : 0u;
      }
      if ((w >> (i & 31)) & 1u) fe_mul_ni(r.l, r.l, b.l);
      fe_mul_ni(b.l, b.l, b.l);
    }
    return r;
  }
  __device__ __noinline__ static void fe_mul_ni(uint32_t* r, const uint32_t* a, const uint32_t* b) { fe_mul<F>(r, a, b); }
};

// Fp2 = Fp[i] / (i^2 + 1)   (reference extension_fields/towers.nim:39-50: coords[0] + coords[1]*i)
template <class F>
struct Fp2 {
  using Params = F;
  using Base = Fp<F>;
  static constexpr int N = F::N;
  static constexpr int WORDS = 2 * F::N;
  static constexpr bool HAS_MUL2 = false;
  Base c0, c1;

  B200_DEV static Fp2 zero() { Fp2 r; r.c0 = Base::zero(); r.c1 = Base::zero(); return r; }
  B200_DEV static Fp2 one() { Fp2 r; r.c0 = Base::one(); r.c1 = Base::zero(); return r; }
  B200_DEV bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  B200_DEV bool operator==(const Fp2& b) const { return (c0 == b.c0) && (c1 == b.c1); }
  B200_DEV Fp2 operator+(const Fp2& b) const { Fp2 r; r.c0 = c0 + b.c0; r.c1 = c1 + b.c1; return r; }
  B200_DEV Fp2 operator-(const Fp2& b) const { Fp2 r; r.c0 = c0 - b.c0; r.c1 = c1 - b.c1; return r; }
  B200_DEV Fp2 operator*(const Fp2& b) const {
    // Karatsuba over i^2 = -1: 3 base multiplications
    Base v0 = c0 * b.c0;
    Base v1 = c1 * b.c1;
    Base s = (c0 + c1) * (b.c0 + b.c1);
    Fp2 r;
    r.c0 = v0 - v1;
    r.c1 = (s - v0) - v1;
    return r;
  }
  B200_DEV Fp2 sqr() const {
    // (a0+a1 i)^2 = (a0+a1)(a0-a1) + 2 a0 a1 i : 2 base multiplications
    Base t = c0 * c1;
    Fp2 r;
    r.c0 = (c0 + c1) * (c0 - c1);
    r.c1 = t + t;
    return r;
  }
  // B200_FP2_CALL (default; measured N = 2^18 G2: 8.28 -> 7.75 ms): the three base-field products of an Fp2 product go through the out-of-line
  // multiplier (Fp::mul_call) even in the hot mixed add -- 27 inlined 381-bit multipliers per mixed add spill registers and
  // overflow the instruction caches.
#ifndef B200_FP2_CALL
#define B200_FP2_CALL 1
#endif
  B200_DEV Fp2 mul_u(const Fp2& b) const {
    if constexpr (B200_FP2_CALL && N >= 12) return (*this) * b;
    Base v0 = c0.mul_u(b.c0);
    Base v1 = c1.mul_u(b.c1);
    Base s = (c0 + c1).mul_u(b.c0 + b.c1);
    Fp2 r;
    r.c0 = v0 - v1;
    r.c1 = (s - v0) - v1;
    return r;
  }
  B200_DEV Fp2 sqr_u() const {
    if constexpr (B200_FP2_CALL && N >= 12) return sqr();
    Base t = c0.mul_u(c1);
    Fp2 r;
    r.c0 = (c0 + c1).mul_u(c0 - c1);
    r.c1 = t + t;
    return r;
  }
  // a*b + c*d: no fused form here (a Karatsuba product is already cheaper than two fused pairs)
  static B200_DEV Fp2 dot2_u(const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d) { return a.mul_u(b) + c.mul_u(d); }
  B200_DEV Fp2 neg() const { Fp2 r; r.c0 = c0.neg(); r.c1 = c1.neg(); return r; }
  B200_DEV Fp2 dbl() const { Fp2 r; r.c0 = c0.dbl(); r.c1 = c1.dbl(); return r; }
  B200_DEV void cneg(bool cond) { c0.cneg(cond); c1.cneg(cond); }
  __device__ __noinline__ Fp2 inv() const {
    Base n = (c0.sqr() + c1.sqr()).inv();
    Fp2 r;
    r.c0 = c0 * n;
    r.c1 = (c1 * n).neg();
    return r;
  }
  B200_DEV uint32_t word(int k) const { return k < N ? c0.l[k] : c1.l[k - N]; }
  B200_DEV void set_word(int k, uint32_t v) { if (k < N) c0.l[k] = v; else c1.l[k - N] = v; }
};

}  // namespace b200
