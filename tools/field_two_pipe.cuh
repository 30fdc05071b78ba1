// Two-pipe variants of the Montgomery multiplier of constantine_b200/csrc/field.cuh -- a recorded NEGATIVE result (round 2), kept
// beside the micro-benchmark that measures it (tools/ubench.cu, profiles/ubench_two_pipe_r2.jsonl, profiles/sass_integer_pipe_r2.txt).
// Not part of the product library.
#pragma once
#include "../constantine_b200/csrc/field.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------------
// Two-pipe form of the same CIOS step. Measured on B200 (profiles/README.md, tools/ubench.cu): IMAD.WIDE.U32 without
// addend issues at the full IMAD rate (63.7 /clk/SM), IMAD.WIDE.U32.X (64-bit addend + carry in/out, what a
// mad.lo.cc/madc.hi.cc pair becomes) takes two passes (31.65 /clk/SM), and the ALU pipe (IADD3[.X], 64 /clk/SM) runs
// beside the multiplier. A product row can therefore be formed either
//   (F) on the multiplier alone: N/2 IMAD.WIDE.U32.X, 2 multiplier passes per product, or
//   (A) as N/2 independent IMAD.WIDE.U32 products (1 pass each) whose N halves are added to the row by ONE
//       add-with-carry chain on the ALU pipe (2 ALU operations per product).
// MASK selects (A) per row: bit 0 a*b_i odd positions, bit 1 a*b_i even positions, bit 2 m*p odd, bit 3 m*p even.
// With two of the four rows on each form the multiplier pipe carries 1.5 passes per product instead of 2.
// ---------------------------------------------------------------------------------------------------------
B200_DEV void p_mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}

template <class F, int MASK>
B200_DEV void mont_round_v(uint32_t* E, uint32_t* O) {
  constexpr int N = F::N;
  const uint32_t m = E[0] * F::INV;
  if constexpr (MASK & 4) {
    uint32_t lo[N / 2], hi[N / 2];
#pragma unroll
    for (int j = 0; j < N; j += 2) p_mul_wide(lo[j / 2], hi[j / 2], m, F::P(j + 1));
    O[0] = p_add_cc(O[0], lo[0]);
    O[1] = p_addc_cc(O[1], hi[0]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      O[j] = p_addc_cc(O[j], lo[j / 2]);
      O[j + 1] = p_addc_cc(O[j + 1], hi[j / 2]);
    }
  } else {
    O[0] = p_mad_lo_cc(m, F::P(1), O[0]);
    O[1] = p_madc_hi_cc(m, F::P(1), O[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      O[j] = p_madc_lo_cc(m, F::P(j + 1), O[j]);
      O[j + 1] = p_madc_hi_cc(m, F::P(j + 1), O[j + 1]);
    }
  }
  if constexpr (MASK & 8) {
    uint32_t lo[N / 2], hi[N / 2];
#pragma unroll
    for (int j = 0; j < N; j += 2) p_mul_wide(lo[j / 2], hi[j / 2], m, F::P(j));
    E[0] = p_add_cc(E[0], lo[0]);
    E[1] = p_addc_cc(E[1], hi[0]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      E[j] = p_addc_cc(E[j], lo[j / 2]);
      E[j + 1] = p_addc_cc(E[j + 1], hi[j / 2]);
    }
  } else {
    E[0] = p_mad_lo_cc(m, F::P(0), E[0]);
    E[1] = p_madc_hi_cc(m, F::P(0), E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      E[j] = p_madc_lo_cc(m, F::P(j), E[j]);
      E[j + 1] = p_madc_hi_cc(m, F::P(j), E[j + 1]);
    }
  }
  O[N - 1] = p_addc(O[N - 1], 0);
}

template <class F, int MASK>
B200_DEV void mont_step_v(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
  constexpr int N = F::N;
  if constexpr (MASK & 1) {
    uint32_t lo[N / 2], hi[N / 2];
#pragma unroll
    for (int j = 0; j < N; j += 2) p_mul_wide(lo[j / 2], hi[j / 2], a[j + 1], bi);
    E[0] = p_add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
      O[j] = p_addc_cc(O[j + 2], lo[j / 2]);
      O[j + 1] = p_addc_cc(O[j + 3], hi[j / 2]);
    }
    O[N - 2] = p_addc_cc(lo[N / 2 - 1], 0);
    O[N - 1] = p_addc(hi[N / 2 - 1], 0);
  } else {
    E[0] = p_add_cc(E[0], O[1]);  // carry goes to position 1 = head of the O chain below
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
      O[j] = p_madc_lo_cc(a[j + 1], bi, O[j + 2]);
      O[j + 1] = p_madc_hi_cc(a[j + 1], bi, O[j + 3]);
    }
    O[N - 2] = p_madc_lo_cc(a[N - 1], bi, 0);
    O[N - 1] = p_madc_hi(a[N - 1], bi, 0);
  }
  if constexpr (MASK & 2) {
    uint32_t lo[N / 2], hi[N / 2];
#pragma unroll
    for (int j = 0; j < N; j += 2) p_mul_wide(lo[j / 2], hi[j / 2], a[j], bi);
    E[0] = p_add_cc(E[0], lo[0]);
    E[1] = p_addc_cc(E[1], hi[0]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      E[j] = p_addc_cc(E[j], lo[j / 2]);
      E[j + 1] = p_addc_cc(E[j + 1], hi[j / 2]);
    }
  } else {
    E[0] = p_mad_lo_cc(a[0], bi, E[0]);
    E[1] = p_madc_hi_cc(a[0], bi, E[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      E[j] = p_madc_lo_cc(a[j], bi, E[j]);
      E[j + 1] = p_madc_hi_cc(a[j], bi, E[j + 1]);
    }
  }
  O[N - 1] = p_addc(O[N - 1], 0);
  mont_round_v<F, MASK>(E, O);
}

// r = a * b * R^-1 mod p, canonical; same value as fe_mul for every MASK.
template <class F, int MASK>
B200_DEV void fe_mul_v(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = F::N;
  static_assert(N % 2 == 0, "even limb count");
  uint32_t A[N], B[N];
#pragma unroll
  for (int j = 0; j < N; j += 2) {
    p_mul_wide(A[j], A[j + 1], a[j], b[0]);
    p_mul_wide(B[j], B[j + 1], a[j + 1], b[0]);
  }
  mont_round_v<F, MASK>(A, B);
#pragma unroll
  for (int i = 1; i < N; i++) {
    if (i & 1)
      mont_step_v<F, MASK>(B, A, a, b[i]);
    else
      mont_step_v<F, MASK>(A, B, a, b[i]);
  }
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

}  // namespace b200
