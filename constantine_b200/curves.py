"""Curve / field parameters for the MSM hot path and the Montgomery constants derived from them.

Moduli, orders and curve coefficients restate
  reference constantine/named/config_fields_and_curves.nim:116-133 (BN254_Snarks),
  :214-229 (Pallas, Vesta), :269-287 (BLS12_381);
generators: constantine/named/constants/bls12_381_generators.nim:20-34,
  bn254_snarks_generators.nim:22-40.
Derived constants follow constantine/named/deriv/precompute.nim:248-261 (spare bits),
  :293-311 (m0ninv = -p^-1 mod 2^w), :349-372 (R^2 mod p, Montgomery one = R mod p).

Pure Python integers: this module is host-side metadata (no arithmetic on the product path).
"""
from dataclasses import dataclass, field
from typing import Optional, Tuple


@dataclass(frozen=True)
class FieldParams:
    name: str          # C identifier fragment, e.g. "bls12_381_fp"
    modulus: int
    bits: int          # declared bit width (reference `bitwidth` / `orderBitwidth`)

    @property
    def limbs64(self) -> int:
        return (self.bits + 63) // 64

    @property
    def limbs32(self) -> int:
        return self.limbs64 * 2

    @property
    def nbytes(self) -> int:
        return self.limbs64 * 8

    @property
    def R(self) -> int:               # Montgomery radix 2^(64*limbs)
        return 1 << (64 * self.limbs64)

    @property
    def one_mont(self) -> int:        # R mod p
        return self.R % self.modulus

    @property
    def r2(self) -> int:              # R^2 mod p
        return (self.R * self.R) % self.modulus

    @property
    def m0ninv64(self) -> int:        # -p^-1 mod 2^64
        return (-pow(self.modulus, -1, 1 << 64)) % (1 << 64)

    @property
    def m0ninv32(self) -> int:
        return self.m0ninv64 & 0xFFFFFFFF

    @property
    def spare_bits(self) -> int:
        return 64 * self.limbs64 - self.modulus.bit_length()

    def to_mont(self, a: int) -> int:
        return (a * self.R) % self.modulus

    def from_mont(self, a: int) -> int:
        return (a * pow(self.R, -1, self.modulus)) % self.modulus


@dataclass(frozen=True)
class CurveParams:
    """One short-Weierstrass group y^2 = x^3 + b (a = 0 for every curve on this path)."""
    name: str                  # "bls12_381_g1", "bn254_snarks_g1", "pallas_ec", "vesta_ec", "bls12_381_g2", "bn254_snarks_g2"
    cprefix: str               # ctt_<cprefix>_{jac,prj}_multi_scalar_mul_... (C symbol family)
    fp: FieldParams
    fr: FieldParams
    ext_degree: int            # 1 = G1 over Fp, 2 = G2 over Fp2 = Fp[i]/(i^2+1)
    b: Tuple[int, ...]         # curve coefficient b as ext_degree Fp coordinates
    gen: Tuple[Tuple[int, ...], Tuple[int, ...]]  # affine generator (x, y), each ext_degree coordinates
    curve_id: int              # index used across the C ABI / kernels
    cofactor: int = 1

    @property
    def coord_bytes(self) -> int:
        return self.fp.nbytes * self.ext_degree

    @property
    def aff_bytes(self) -> int:
        return 2 * self.coord_bytes

    @property
    def jac_bytes(self) -> int:
        return 3 * self.coord_bytes

    @property
    def scalar_bits(self) -> int:
        return self.fr.bits


BN254_FP = FieldParams("bn254_snarks_fp", 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47, 254)
BN254_FR = FieldParams("bn254_snarks_fr", 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001, 254)
BLS12_381_FP = FieldParams(
    "bls12_381_fp",
    0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab, 381)
BLS12_381_FR = FieldParams("bls12_381_fr", 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 255)
PALLAS_FP = FieldParams("pallas_fp", 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001, 255)
PALLAS_FR = FieldParams("pallas_fr", 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001, 255)
VESTA_FP = FieldParams("vesta_fp", PALLAS_FR.modulus, 255)
VESTA_FR = FieldParams("vesta_fr", PALLAS_FP.modulus, 255)

BLS12_381_G1 = CurveParams(
    "bls12_381_g1", "bls12_381_g1", BLS12_381_FP, BLS12_381_FR, 1, (4,),
    ((0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,),
     (0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1,)),
    curve_id=0, cofactor=0x396c8c005555e1568c00aaab0000aaab)
BN254_G1 = CurveParams("bn254_snarks_g1", "bn254_snarks_g1", BN254_FP, BN254_FR, 1, (3,), ((1,), (2,)), curve_id=1)
# Pasta generators: (-1, 2) on both curves (y^2 = x^3 + 5  =>  (-1)^3 + 5 = 4 = 2^2).
PALLAS = CurveParams("pallas_ec", "pallas_ec", PALLAS_FP, PALLAS_FR, 1, (5,),
                     ((PALLAS_FP.modulus - 1,), (2,)), curve_id=2)
VESTA = CurveParams("vesta_ec", "vesta_ec", VESTA_FP, VESTA_FR, 1, (5,),
                    ((VESTA_FP.modulus - 1,), (2,)), curve_id=3)
BLS12_381_G2 = CurveParams(
    "bls12_381_g2", "bls12_381_g2", BLS12_381_FP, BLS12_381_FR, 2, (4, 4),
    ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
      0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
     (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
      0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)),
    curve_id=4,
    cofactor=0x5d543a95414e7f1091d50792876a202cd91de4547085abaa68a205b2e5a7ddfa628f1cb4d9e82ef21537e293a6691ae1616ec6e786f0c70cf1c38e31c7238e5)


def _bn254_g2_b():
    # b' = 3 / (9 + i) in Fp2 = Fp[i]/(i^2+1)   (reference config_fields_and_curves.nim:126, D-twist)
    p = BN254_FP.modulus
    inv_norm = pow(9 * 9 + 1, -1, p)
    return ((3 * 9 * inv_norm) % p, (-3 * inv_norm) % p)


BN254_G2 = CurveParams(
    "bn254_snarks_g2", "bn254_snarks_g2", BN254_FP, BN254_FR, 2, _bn254_g2_b(),
    ((0x1800DEEF121F1E76426A00665E5C4479674322D4F75EDADD46DEBD5CD992F6ED,
      0x198E9393920D483A7260BFB731FB5D25F1AA493335A9E71297E485B7AEF312C2),
     (0x12C85EA5DB8C6DEB4AAB71808DCB408FE3D1E7690C43D37B4CE6CC0166FA7DAA,
      0x090689D0585FF075EC9E99AD690C3395BC4B313370B38EF355ACDADCD122975B)),
    curve_id=5,
    cofactor=0x30644e72e131a029b85045b68181585e06ceecda572a2489345f2299c0f9fa8d)

CURVES = {c.name: c for c in (BLS12_381_G1, BN254_G1, PALLAS, VESTA, BLS12_381_G2, BN254_G2)}
CURVES_BY_ID = {c.curve_id: c for c in CURVES.values()}
FIELDS = {f.name: f for f in (BN254_FP, BN254_FR, BLS12_381_FP, BLS12_381_FR, PALLAS_FP, PALLAS_FR, VESTA_FP, VESTA_FR)}
