// EIP-2537 wire format around the GPU MSM: ctt_eth_evm_bls12381_g1msm / ctt_eth_evm_bls12381_g2msm
// (SURVEY.md section 8f item 3 -- a caller of the hot path, one format either side of it).
//
// Replaces reference constantine/ethereum_evm_precompiles.nim:894-975 (eth_evm_bls12381_g1msm) and :977-1060 (g2msm),
// exported to C as include/constantine/protocols/ethereum_evm_precompiles.h:386-389, 419-422:
//   - input  = k x (point || 32-byte scalar), big-endian; an Fp coordinate is 64 bytes whose top 16 bytes must be zero
//     (parseEip2537, :258-288), an Fp2 coordinate is c0 || c1; (0,0) encodes infinity;
//   - every point must be on the curve and in the prime-order subgroup (fromRawCoords, :316-389);
//   - scalars may exceed the group order and are reduced mod r first (:948-961);
//   - output = the affine sum, same encoding (128 / 256 bytes); status codes = reference CttEVMStatus (:49-57).
// Parsing, the checks and the final affine conversion are host code (tiny, per-pair work); the sum itself is the same
// engine call as every other entry point (msm_host). There is no CPU path for the MSM.
#include "msm_hooks.cuh"

namespace b200 {
B200_DECLARE_CURVE(Bls12381G1)
B200_DECLARE_CURVE(Bls12381G2)

namespace evm {

enum Status : int { Success = 0, InvalidInputSize = 1, InvalidOutputSize = 2, IntLargerThanModulus = 3, PointNotOnCurve = 4,
                    PointNotInSubgroup = 5 };

using Fp = host::HFp<Bls12381Fp>;
using Fp2 = host::HFp2<Bls12381Fp>;

static Fp fp_r2() { Fp r; for (int i = 0; i < 6; i++) r.l[i] = Bls12381Fp::R264(i); return r; }

// 64-byte big-endian integer -> Montgomery residue (reference parseEip2537): top 16 bytes zero, value < p
static bool parse_fp(Fp& out, const uint8_t* src) {
  bool ok = true;
  for (int i = 0; i < 16; i++) ok = ok && (src[i] == 0);
  Fp raw;
  for (int limb = 0; limb < 6; limb++) {
    uint64_t v = 0;
    const uint8_t* p = src + 16 + (5 - limb) * 8;   // most significant limb first in the byte string
    for (int b = 0; b < 8; b++) v = (v << 8) | p[b];
    raw.l[limb] = v;
  }
  if (!ok || Fp::geq_p(raw.l)) return false;
  out = raw * fp_r2();   // to Montgomery form
  return true;
}
static bool parse_coord(Fp& out, const uint8_t* src) { return parse_fp(out, src); }
static bool parse_coord(Fp2& out, const uint8_t* src) { return parse_fp(out.c0, src) && parse_fp(out.c1, src + 64); }

static void write_fp(uint8_t* dst, const Fp& mont) {
  Fp one_raw = Fp::zero();
  one_raw.l[0] = 1;
  Fp v = mont * one_raw;   // from Montgomery form: multiply by the integer 1
  memset(dst, 0, 16);
  for (int limb = 0; limb < 6; limb++) {
    uint64_t x = v.l[limb];
    uint8_t* p = dst + 16 + (5 - limb) * 8;
    for (int b = 7; b >= 0; b--) { p[b] = (uint8_t)x; x >>= 8; }
  }
}
static void write_coord(uint8_t* dst, const Fp& v) { write_fp(dst, v); }
static void write_coord(uint8_t* dst, const Fp2& v) { write_fp(dst, v.c0); write_fp(dst + 64, v.c1); }

template <class T> T curve_b();
template <> Fp curve_b<Fp>() {   // b = 4 (reference config_fields_and_curves.nim:281-282)
  Fp four = Fp::one(); four = four.dbl().dbl(); return four;
}
template <> Fp2 curve_b<Fp2>() { // b = 4 (1 + i) on the M-twist (:276,287)
  Fp2 r; r.c0 = curve_b<Fp>(); r.c1 = curve_b<Fp>(); return r;
}

// group order r, 255 bits (reference config_fields_and_curves.nim:277)
static const uint64_t ORDER[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};

// [r]P == infinity, by plain double-and-add on the host (the reference uses an endomorphism-based test; any complete
// test accepts the same set of points)
template <class T>
static bool in_subgroup(const T& x, const T& y) {
  host::HXyzz<T> base; base.x = x; base.y = y; base.zz = T::one(); base.zzz = T::one();
  host::HXyzz<T> acc = host::HXyzz<T>::inf();
  for (int bit = 254; bit >= 0; bit--) {
    acc = host::xyzz_dbl(acc);
    if ((ORDER[bit >> 6] >> (bit & 63)) & 1) acc = host::xyzz_add(acc, base);
  }
  return acc.is_inf();
}

// 32-byte big-endian scalar -> canonical little-endian limbs reduced mod r (s < 2^256 < 3r: at most two subtractions)
static void parse_scalar(uint64_t out[4], const uint8_t* src) {
  for (int limb = 0; limb < 4; limb++) {
    uint64_t v = 0;
    const uint8_t* p = src + (3 - limb) * 8;
    for (int b = 0; b < 8; b++) v = (v << 8) | p[b];
    out[limb] = v;
  }
  for (int rounds = 0; rounds < 3; rounds++) {
    bool geq = true;
    for (int i = 3; i >= 0; i--) {
      if (out[i] > ORDER[i]) { geq = true; break; }
      if (out[i] < ORDER[i]) { geq = false; break; }
    }
    if (!geq) break;
    unsigned __int128 borrow = 0;
    for (int i = 0; i < 4; i++) {
      unsigned __int128 d = (unsigned __int128)out[i] - ORDER[i] - borrow;
      out[i] = (uint64_t)d;
      borrow = (d >> 64) & 1;
    }
  }
}

template <class C, class T>
static int msm_precompile(uint8_t* r, size_t r_len, const uint8_t* inputs, size_t inputs_len) {
  constexpr size_t CB = sizeof(T) / 6 / 8 * 64;   // encoded bytes per coordinate: 64 (Fp) or 128 (Fp2)
  constexpr size_t PAIR = 2 * CB + 32;
  if (inputs_len == 0 || inputs_len % PAIR != 0) return InvalidInputSize;
  if (r_len != 2 * CB) return InvalidOutputSize;
  const size_t n = inputs_len / PAIR;
  std::vector<uint64_t> coefs(4 * n);
  std::vector<T> pts(2 * n);
  const T b = curve_b<T>();
  for (size_t i = 0; i < n; i++) {
    const uint8_t* p = inputs + i * PAIR;
    T x, y;
    if (!parse_coord(x, p) || !parse_coord(y, p + CB)) return IntLargerThanModulus;
    if (!(x.is_zero() && y.is_zero())) {
      if (!(y.sqr() == x.sqr() * x + b)) return PointNotOnCurve;
      if (!in_subgroup(x, y)) return PointNotInSubgroup;
    }
    pts[2 * i] = x; pts[2 * i + 1] = y;
    parse_scalar(&coefs[4 * i], p + 2 * CB);
  }
  struct { T X, Y, Z; } jac;
  msm_host<C>(&jac, coefs.data(), pts.data(), n, /*fr_mont=*/false, OUT_JAC);
  memset(r, 0, r_len);
  if (!jac.Z.is_zero()) {        // affine: x = X / Z^2, y = Y / Z^3
    T zi = jac.Z.inv();
    T zi2 = zi.sqr();
    write_coord(r, jac.X * zi2);
    write_coord(r + CB, jac.Y * zi2 * zi);
  }
  return Success;
}

}  // namespace evm
}  // namespace b200

extern "C" {
// reference include/constantine/protocols/ethereum_evm_precompiles.h:386-389
unsigned char ctt_eth_evm_bls12381_g1msm(unsigned char* r, size_t r_len, const unsigned char* inputs, size_t inputs_len) {
  return (unsigned char)b200::evm::msm_precompile<b200::Bls12381G1, b200::evm::Fp>(r, r_len, inputs, inputs_len);
}
// reference include/constantine/protocols/ethereum_evm_precompiles.h:419-422
unsigned char ctt_eth_evm_bls12381_g2msm(unsigned char* r, size_t r_len, const unsigned char* inputs, size_t inputs_len) {
  return (unsigned char)b200::evm::msm_precompile<b200::Bls12381G2, b200::evm::Fp2>(r, r_len, inputs, inputs_len);
}
}
