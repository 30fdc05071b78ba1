// EIP-4844 blob_to_kzg_commitment on the GPU MSM (SURVEY.md section 8f item 2 -- a caller of the hot path with byte-pinned
// known answers).
//
// Replaces reference constantine/ethereum_eip4844_kzg_parallel.nim:125-159 (blob_to_kzg_commitment_parallel) and its serial
// twin constantine/ethereum_eip4844_kzg.nim (blob_to_kzg_commitment), C declarations
// include/constantine/protocols/ethereum_eip4844_kzg_parallel.h:40-45 and ethereum_eip4844_kzg.h:106-110:
//   1. blob = 4096 x 32 bytes, big-endian field elements; each must be < r, else cttEthKzg_ScalarLargerThanCurveOrder
//      (blob_to_bigint_polynomial_parallel, :47-85 -> bytes_to_bls_bigint, constantine/serialization/codecs_status_codes.nim);
//   2. commitment = sum_i blob_i * SRS_i over the 4096 G1 points of the trusted setup in Lagrange form, bit-reversal
//      permuted (kzg_commit_parallel, constantine/commitments/kzg_parallel.nim:33-47 = ONE 4096-term MSM);
//   3. the affine result serialised in the 48-byte compressed ZCash format (serialize_g1_compressed,
//      constantine/serialization/codecs_bls12_381.nim).
// The reference keeps the SRS inside an opaque EthereumKZGContext loaded from a trusted-setup file; here the context is the
// 4096 points resident in HBM (ctt_b200_bases_upload, optionally with the precomputed window table), built either from the
// affine Montgomery structs a Constantine caller already holds (ctx.srs_lagrange_brp_g1) or from the 48-byte compressed
// encodings every trusted-setup file carries. Parsing, the range check and the final inversion are host code (4096 items);
// the MSM is the same engine call as every other entry point. There is no CPU path for it.
#define CTT_B200_BUILDING_LIBRARY
#include "../../include/ctt_b200_msm.h"
#include "host_field.hpp"
#include <cstdlib>
#include <vector>

namespace b200 {
namespace kzg {

using Fp = host::HFp<Bls12381Fp>;
constexpr size_t FIELD_ELEMENTS_PER_BLOB = 4096;

// reference include/constantine/protocols/ethereum_eip4844_kzg.h:28-39 (cttEthKzg_* status codes)
enum Status : int { Success = 0, VerificationFailure = 1, InputsLengthsMismatch = 2, ScalarZero = 3, ScalarLargerThanCurveOrder = 4,
                    EccInvalidEncoding = 5, EccCoordinateGreaterThanOrEqualModulus = 6, EccPointNotOnCurve = 7, EccPointNotInSubgroup = 8 };

// group order r (reference config_fields_and_curves.nim:277)
static const uint64_t ORDER[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};

static Fp fp_r2() { Fp r; for (int i = 0; i < 6; i++) r.l[i] = Bls12381Fp::R264(i); return r; }
static Fp from_mont(const Fp& m) { Fp one_raw = Fp::zero(); one_raw.l[0] = 1; return m * one_raw; }

// a^e for a little-endian 6-limb exponent
static Fp fp_pow(const Fp& a, const uint64_t e[6]) {
  Fp r = Fp::one(), b = a;
  for (int i = 0; i < 384; i++) {
    if ((e[i >> 6] >> (i & 63)) & 1) r = r * b;
    b = b.sqr();
  }
  return r;
}

// y > (p - 1) / 2 as integers ("lexicographically largest", the sign bit of the compressed format)
static bool is_lexicographically_largest(const Fp& y_mont) {
  const Fp y = from_mont(y_mont);
  uint64_t half[6];   // (p - 1) / 2
  {
    uint64_t t[6];
    for (int i = 0; i < 6; i++) t[i] = Bls12381Fp::P64(i);
    t[0] -= 1;
    for (int i = 0; i < 6; i++) half[i] = (t[i] >> 1) | (i + 1 < 6 ? t[i + 1] << 63 : 0);
  }
  for (int i = 5; i >= 0; i--) {
    if (y.l[i] > half[i]) return true;
    if (y.l[i] < half[i]) return false;
  }
  return false;
}

// 48-byte compressed G1 (ZCash flags: 0x80 compressed, 0x40 infinity, 0x20 y is the larger root) -> affine Montgomery (x, y);
// infinity -> (0, 0). No subgroup check (trusted-setup points; the reference checks them when it loads the file).
static int decompress_g1(Fp& x, Fp& y, const uint8_t src[48]) {
  const uint8_t flags = src[0];
  if (!(flags & 0x80)) return EccInvalidEncoding;
  if (flags & 0x40) {
    if (flags & 0x3F) return EccInvalidEncoding;
    for (int i = 1; i < 48; i++) if (src[i]) return EccInvalidEncoding;
    x = Fp::zero(); y = Fp::zero();
    return Success;
  }
  Fp raw;
  for (int limb = 0; limb < 6; limb++) {
    uint64_t v = 0;
    const uint8_t* p = src + (5 - limb) * 8;
    for (int b = 0; b < 8; b++) v = (v << 8) | (uint8_t)((limb == 5 && b == 0) ? (p[b] & 0x1F) : p[b]);
    raw.l[limb] = v;
  }
  if (Fp::geq_p(raw.l)) return EccCoordinateGreaterThanOrEqualModulus;
  x = raw * fp_r2();
  Fp four = Fp::one(); four = four.dbl().dbl();
  const Fp rhs = x.sqr() * x + four;                       // y^2 = x^3 + 4
  uint64_t e[6];                                           // (p + 1) / 4: p = 3 mod 4
  {
    uint64_t t[6];
    unsigned __int128 c = 1;
    for (int i = 0; i < 6; i++) { c += Bls12381Fp::P64(i); t[i] = (uint64_t)c; c >>= 64; }
    for (int i = 0; i < 6; i++) e[i] = (t[i] >> 2) | (i + 1 < 6 ? t[i + 1] << 62 : 0);
  }
  Fp root = fp_pow(rhs, e);
  if (!(root.sqr() == rhs)) return EccPointNotOnCurve;
  if (is_lexicographically_largest(root) != ((flags & 0x20) != 0)) root = root.neg();
  y = root;
  return Success;
}

static void compress_g1(uint8_t dst[48], const Fp& x_mont, const Fp& y_mont, bool inf) {
  memset(dst, 0, 48);
  if (inf) { dst[0] = 0xC0; return; }
  const Fp x = from_mont(x_mont);
  for (int limb = 0; limb < 6; limb++) {
    uint64_t v = x.l[limb];
    uint8_t* p = dst + (5 - limb) * 8;
    for (int b = 7; b >= 0; b--) { p[b] = (uint8_t)v; v >>= 8; }
  }
  dst[0] |= 0x80;
  if (is_lexicographically_largest(y_mont)) dst[0] |= 0x20;
}

struct Context {
  ctt_b200_bases* bases = nullptr;
};

}  // namespace kzg
}  // namespace b200

using namespace b200::kzg;

extern "C" {

struct ctt_b200_eth_kzg_context;

// SRS as the reference holds it (ctx.srs_lagrange_brp_g1: 4096 EC_ShortW_Aff[Fp[BLS12_381], G1], Montgomery residues)
ctt_b200_eth_kzg_context* ctt_b200_eth_kzg_context_new(const void* srs_lagrange_brp_g1_aff) {
  Context* c = new Context;
  c->bases = ctt_b200_bases_upload(CTT_B200_BLS12_381_G1, srs_lagrange_brp_g1_aff, FIELD_ELEMENTS_PER_BLOB);
  if (!c->bases) { delete c; return nullptr; }
  return reinterpret_cast<ctt_b200_eth_kzg_context*>(c);
}

// SRS as trusted-setup files carry it: 4096 x 48-byte compressed G1 points, already in the bit-reversal-permuted Lagrange order.
// Returns null and writes the cttEthKzg_* reason to *status (if not null) when a point does not decode.
ctt_b200_eth_kzg_context* ctt_b200_eth_kzg_context_new_compressed(const unsigned char* srs_compressed, int* status) {
  std::vector<Fp> pts(2 * FIELD_ELEMENTS_PER_BLOB);
  for (size_t i = 0; i < FIELD_ELEMENTS_PER_BLOB; i++) {
    const int rc = decompress_g1(pts[2 * i], pts[2 * i + 1], srs_compressed + 48 * i);
    if (rc != Success) { if (status) *status = rc; return nullptr; }
  }
  if (status) *status = Success;
  return ctt_b200_eth_kzg_context_new(pts.data());
}

// one-time: window table next to the resident SRS (all windows share one bucket set, no doublings at run time)
int ctt_b200_eth_kzg_context_precompute(ctt_b200_eth_kzg_context* ctx, int c) {
  Context* k = reinterpret_cast<Context*>(ctx);
  if (!k) return -1;
  return ctt_b200_bases_precompute(k->bases, c);
}

void ctt_b200_eth_kzg_context_delete(ctt_b200_eth_kzg_context* ctx) {
  Context* k = reinterpret_cast<Context*>(ctx);
  if (!k) return;
  ctt_b200_bases_free(k->bases);
  delete k;
}

// reference ctt_eth_kzg_blob_to_kzg_commitment_parallel(tp, ctx, dst, blob) / ctt_eth_kzg_blob_to_kzg_commitment(ctx, dst, blob)
// with the resident-SRS context above. Returns the reference's ctt_eth_kzg_status values.
unsigned char ctt_b200_eth_kzg_blob_to_kzg_commitment(const ctt_b200_eth_kzg_context* ctx, unsigned char dst[48], const unsigned char* blob) {
  const Context* k = reinterpret_cast<const Context*>(ctx);
  if (!k) return (unsigned char)VerificationFailure;
  std::vector<uint64_t> coefs(4 * FIELD_ELEMENTS_PER_BLOB);
  for (size_t i = 0; i < FIELD_ELEMENTS_PER_BLOB; i++) {
    uint64_t* o = &coefs[4 * i];
    const uint8_t* src = blob + 32 * i;
    for (int limb = 0; limb < 4; limb++) {
      uint64_t v = 0;
      const uint8_t* p = src + (3 - limb) * 8;
      for (int b = 0; b < 8; b++) v = (v << 8) | p[b];
      o[limb] = v;
    }
    bool geq = true;                       // canonical: < r (zero is a valid evaluation)
    for (int j = 3; j >= 0; j--) {
      if (o[j] > ORDER[j]) { geq = true; break; }
      if (o[j] < ORDER[j]) { geq = false; break; }
    }
    if (geq) return (unsigned char)ScalarLargerThanCurveOrder;
  }
  struct { Fp X, Y, Z; } jac;
  if (ctt_b200_msm_cached_bases(k->bases, CTT_B200_OUT_JAC, &jac, coefs.data(), FIELD_ELEMENTS_PER_BLOB, /*fr_mont=*/0) != 0)
    return (unsigned char)VerificationFailure;
  if (jac.Z.is_zero()) { compress_g1(dst, Fp::zero(), Fp::zero(), true); return (unsigned char)Success; }
  const Fp zi = jac.Z.inv();
  const Fp zi2 = zi.sqr();
  compress_g1(dst, jac.X * zi2, jac.Y * zi2 * zi, false);
  return (unsigned char)Success;
}

}  // extern "C"
