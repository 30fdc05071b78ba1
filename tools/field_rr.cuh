// Reduced-radix prime-field arithmetic for sm_100a: limbs of W < 32 bits in 32-bit registers, column sums in 64-bit
// accumulators, NO carry chains in the multiplier.
//
// Why: on B200 IMAD.WIDE.U32 without carry issues at 64 lanes/clk/SM, while the carry-chained form
// (IMAD.WIDE.U32.X with predicate carry in/out, what mad.lo.cc/madc.hi.cc pairs compile to) issues at half that
// (profiles/ubench_r1.jsonl). With W-bit limbs (W = 28 for 381-bit, 29 for 254/255-bit fields) the 2n column sums of
// n^2 + n^2 products fit 64 bits, so every product is a plain `mad.wide.u32` and carries are resolved once per
// column with shifts on the ALU pipe, which runs in parallel with the FMA pipe.
//
// Same field, same values as field.cuh / the reference (constantine/math/arithmetic/limbs_montgomery.nim:180-217 is
// the CIOS this restructures); only the machine representation differs:
//   element x is held as the lazy residue  X' = x * R' mod p  (+ k*p),  R' = 2^(W*n),  limbs < 2^W after normalise.
// The ABI's Montgomery residue x*2^(64*N64) becomes X' by a left shift of (W*n - 64*N64) bits, i.e. for free while
// re-limbing. Results of mul are < 2p; add/sub are lazy (no modular reduction), equality tests canonicalise.
#pragma once
#include <cstdint>
#include "field_constants.cuh"

namespace b200 {

#ifndef B200_DEV
#define B200_DEV __device__ __forceinline__
#endif

// T += a * b  (32 x 32 -> 64, plain IMAD.WIDE.U32: no carry flag involved)
B200_DEV void mad_wide(uint64_t& t, uint32_t a, uint32_t b) { asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a), "r"(b)); }

template <class F, int W_, int NL_>
struct RRParams {
  static constexpr int W = W_;            // bits per limb
  static constexpr int NL = NL_;          // limbs
  static constexpr uint32_t MASK = (1u << W_) - 1u;
  static constexpr int SHIFT = W_ * NL_ - 64 * F::N64;   // R' / R = 2^SHIFT
  static_assert(SHIFT >= 0 && SHIFT < 32, "re-limbing shift");
  // limb i of the modulus in radix 2^W
  __host__ __device__ static constexpr uint32_t P(int i) {
    // bits [W*i, W*i + W) of p, from its 32-bit limbs
    int bit = W_ * i;
    int w = bit >> 5, s = bit & 31;
    uint64_t lo = (w < F::N) ? F::P(w) : 0u;
    uint64_t hi = (w + 1 < F::N) ? F::P(w + 1) : 0u;
    return (uint32_t)(((lo | (hi << 32)) >> s) & ((1u << W_) - 1u));
  }
  // -p^-1 mod 2^W  (low W bits of the 32-bit constant)
  static constexpr uint32_t INV = F::INV & ((1u << W_) - 1u);
};

template <class F, int W, int NL>
struct FpRR {
  using PR = RRParams<F, W, NL>;
  static constexpr uint32_t MASK = PR::MASK;
  uint32_t l[NL];

  B200_DEV static FpRR zero() {
    FpRR r;
#pragma unroll
    for (int i = 0; i < NL; i++) r.l[i] = 0;
    return r;
  }

  // carry-propagate so that limbs 0..NL-2 are < 2^W (top limb keeps the excess)
  B200_DEV void normalise() {
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < NL - 1; i++) {
      uint32_t v = l[i] + c;
      l[i] = v & MASK;
      c = v >> W;
    }
    l[NL - 1] += c;
  }

  // from the ABI's 32-bit-limb Montgomery residue (canonical, < p): X' = x << SHIFT, re-limbed
  B200_DEV static FpRR from_abi(const uint32_t* a) {
    FpRR r;
#pragma unroll
    for (int i = 0; i < NL; i++) {
      int bit = W * i - PR::SHIFT;  // bit position in the ABI integer that lands at bit 0 of limb i
      uint32_t v;
      if (bit < 0) {
        v = (a[0] << (-bit)) & MASK;  // only limb 0 when SHIFT > 0 (SHIFT < W)
      } else {
        int w = bit >> 5, s = bit & 31;
        uint64_t lo = (w < F::N) ? a[w] : 0u;
        uint64_t hi = (w + 1 < F::N) ? a[w + 1] : 0u;
        v = (uint32_t)(((lo | (hi << 32)) >> s)) & MASK;
      }
      r.l[i] = v;
    }
    return r;
  }

  // r = a*b/R' (mod p), 0 <= r < 2p, limbs normalised.  Inputs: limbs < 2^(W+1) (one lazy add allowed), values < 2^5 p.
  B200_DEV FpRR operator*(const FpRR& b) const {
    uint64_t T[2 * NL];
#pragma unroll
    for (int k = 0; k < 2 * NL; k++) T[k] = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) {
      // row i of a*b
#pragma unroll
      for (int j = 0; j < NL; j++) mad_wide(T[i + j], l[j], b.l[i]);
      // reduction step i: make column i divisible by 2^W, push its carry into column i+1
      uint32_t m = ((uint32_t)T[i] * PR::INV) & MASK;
#pragma unroll
      for (int j = 0; j < NL; j++) mad_wide(T[i + j], m, PR::P(j));
      T[i + 1] += T[i] >> W;
    }
    FpRR r;
#pragma unroll
    for (int k = 0; k < NL - 1; k++) {
      r.l[k] = (uint32_t)T[NL + k] & MASK;
      T[NL + k + 1] += T[NL + k] >> W;
    }
    r.l[NL - 1] = (uint32_t)T[2 * NL - 1];
    return r;
  }
  // product-scanning variant: one 64-bit column accumulator (two interleaved for ILP)
  B200_DEV FpRR mul_ps(const FpRR& b) const {
    uint32_t m[NL];
    FpRR r;
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 2 * NL - 1; k++) {
      uint64_t acc0 = carry, acc1 = 0;
#pragma unroll
      for (int i = 0; i < NL; i++) {
        int j = k - i;
        if (j < 0 || j >= NL) continue;
        mad_wide(acc0, l[j], b.l[i]);
        if (i < k && i < NL && !(k < NL && i == k)) { if (k - i >= 0 && k - i < NL && i < NL && (k >= NL || i < k)) mad_wide(acc1, m[i], PR::P(j)); }
      }
      acc0 += acc1;
      if (k < NL) {
        m[k] = ((uint32_t)acc0 * PR::INV) & MASK;
        mad_wide(acc0, m[k], PR::P(0));
        carry = acc0 >> W;
      } else {
        r.l[k - NL] = (uint32_t)acc0 & MASK;
        carry = acc0 >> W;
      }
    }
    r.l[NL - 1] = (uint32_t)carry;
    return r;
  }
  B200_DEV FpRR sqr() const { return (*this) * (*this); }

  // lazy add: limb-wise, no reduction
  B200_DEV FpRR operator+(const FpRR& b) const {
    FpRR r;
#pragma unroll
    for (int i = 0; i < NL; i++) r.l[i] = l[i] + b.l[i];
    return r;
  }
};

}  // namespace b200
