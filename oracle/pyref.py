"""ORACLE (test infrastructure, not product code) -- exact big-integer tier.

Affine short-Weierstrass arithmetic (a = 0) over Fp or Fp2 = Fp[i]/(i^2+1) on Python ints, a naive
double-and-add scalar multiplication, a naive MSM  sum_i [k_i]P_i, and the byte-layout helpers that turn
Python values into the reference's C structs (little-endian 64-bit limbs, Montgomery residues;
reference constantine/platforms/abstractions.nim:132-143, include/constantine/curves/bls12_381.h:19-27).

This is the *definition* the C restatement (oracle/msm_oracle.c) and the CUDA path are pinned to; it follows
the naive side of the reference's own differential tests
(reference tests/math_elliptic_curves/t_ec_template.nim:1466-1480: naive sum of scalarMul vs MSM).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constantine_b200.curves import CurveParams, FieldParams  # noqa: E402


# ---------------------------------------------------------------- field elements as tuples of ints
def f_add(a, b, p):
    return tuple((x + y) % p for x, y in zip(a, b))


def f_sub(a, b, p):
    return tuple((x - y) % p for x, y in zip(a, b))


def f_neg(a, p):
    return tuple((-x) % p for x in a)


def f_mul(a, b, p):
    if len(a) == 1:
        return ((a[0] * b[0]) % p,)
    # complex multiplication, i^2 = -1 (reference extension_fields/towers.nim:798-885)
    return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)


def f_inv(a, p):
    if len(a) == 1:
        return (pow(a[0], -1, p),)
    n = pow(a[0] * a[0] + a[1] * a[1], -1, p)
    return ((a[0] * n) % p, (-a[1] * n) % p)


def f_zero(d):
    return (0,) * d


def f_is_zero(a):
    return all(x == 0 for x in a)


# ---------------------------------------------------------------- affine points: None = infinity
def on_curve(P, curve: CurveParams):
    if P is None:
        return True
    p = curve.fp.modulus
    x, y = P
    lhs = f_mul(y, y, p)
    rhs = f_add(f_mul(f_mul(x, x, p), x, p), tuple(curve.b), p)
    return lhs == rhs


def ec_neg(P, curve):
    if P is None:
        return None
    return (P[0], f_neg(P[1], curve.fp.modulus))


def ec_add(P, Q, curve):
    p = curve.fp.modulus
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if y1 == y2 and not f_is_zero(y1):
            three = (3,) + (0,) * (curve.ext_degree - 1)
            two = (2,) + (0,) * (curve.ext_degree - 1)
            lam = f_mul(f_mul(three, f_mul(x1, x1, p), p), f_inv(f_mul(two, y1, p), p), p)
        else:
            return None
    else:
        lam = f_mul(f_sub(y2, y1, p), f_inv(f_sub(x2, x1, p), p), p)
    x3 = f_sub(f_sub(f_mul(lam, lam, p), x1, p), x2, p)
    y3 = f_sub(f_mul(lam, f_sub(x1, x3, p), p), y1, p)
    return (x3, y3)


def ec_mul(k, P, curve):
    """[k]P for any integer k >= 0 (not reduced mod the group order: the reference tests feed
    scalars >= r, tests/parallel/t_ec_template_parallel.nim:176)."""
    R = None
    Q = P
    while k:
        if k & 1:
            R = ec_add(R, Q, curve)
        Q = ec_add(Q, Q, curve)
        k >>= 1
    return R


def msm_naive(scalars, points, curve):
    acc = None
    for k, P in zip(scalars, points):
        acc = ec_add(acc, ec_mul(k, P, curve), curve)
    return acc


# ---------------------------------------------------------------- Jacobian arithmetic on ints (fast exact tier)
def _jac_dbl(P, p, d):
    X, Y, Z = P
    if f_is_zero(Z):
        return P
    A = f_mul(X, X, p)
    B = f_mul(Y, Y, p)
    C = f_mul(B, B, p)
    t = f_add(X, B, p)
    D = f_sub(f_sub(f_mul(t, t, p), A, p), C, p)
    D = f_add(D, D, p)
    E = f_add(f_add(A, A, p), A, p)
    F = f_mul(E, E, p)
    X3 = f_sub(F, f_add(D, D, p), p)
    C8 = f_add(C, C, p)
    C8 = f_add(C8, C8, p)
    C8 = f_add(C8, C8, p)
    Y3 = f_sub(f_mul(E, f_sub(D, X3, p), p), C8, p)
    Z3 = f_mul(Y, Z, p)
    Z3 = f_add(Z3, Z3, p)
    return (X3, Y3, Z3)


def _jac_add(P, Q, p, d):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if f_is_zero(Z1):
        return Q
    if f_is_zero(Z2):
        return P
    Z1Z1 = f_mul(Z1, Z1, p)
    Z2Z2 = f_mul(Z2, Z2, p)
    U1 = f_mul(X1, Z2Z2, p)
    U2 = f_mul(X2, Z1Z1, p)
    S1 = f_mul(f_mul(Y1, Z2, p), Z2Z2, p)
    S2 = f_mul(f_mul(Y2, Z1, p), Z1Z1, p)
    H = f_sub(U2, U1, p)
    Rr = f_sub(S2, S1, p)
    if f_is_zero(H):
        if f_is_zero(Rr):
            return _jac_dbl(P, p, d)
        one = (1,) + (0,) * (d - 1)
        return (one, one, f_zero(d))
    HH = f_mul(H, H, p)
    HHH = f_mul(H, HH, p)
    V = f_mul(U1, HH, p)
    X3 = f_sub(f_sub(f_mul(Rr, Rr, p), HHH, p), f_add(V, V, p), p)
    Y3 = f_sub(f_mul(Rr, f_sub(V, X3, p), p), f_mul(S1, HHH, p), p)
    Z3 = f_mul(f_mul(Z1, Z2, p), H, p)
    return (X3, Y3, Z3)


def jac_from_affine(P, d):
    one = (1,) + (0,) * (d - 1)
    if P is None:
        return (one, one, f_zero(d))
    return (P[0], P[1], one)


def jac_to_affine(P, p):
    X, Y, Z = P
    if f_is_zero(Z):
        return None
    zi = f_inv(Z, p)
    zi2 = f_mul(zi, zi, p)
    return (f_mul(X, zi2, p), f_mul(Y, f_mul(zi2, zi, p), p))


def ec_mul_fast(k, P, curve):
    p, d = curve.fp.modulus, curve.ext_degree
    R = jac_from_affine(None, d)
    Q = jac_from_affine(P, d)
    while k:
        if k & 1:
            R = _jac_add(R, Q, p, d)
        Q = _jac_dbl(Q, p, d)
        k >>= 1
    return jac_to_affine(R, p)


def msm_naive_fast(scalars, points, curve):
    p, d = curve.fp.modulus, curve.ext_degree
    acc = jac_from_affine(None, d)
    for k, P in zip(scalars, points):
        if P is None or k == 0:
            continue
        R = jac_from_affine(None, d)
        Q = jac_from_affine(P, d)
        while k:
            if k & 1:
                R = _jac_add(R, Q, p, d)
            Q = _jac_dbl(Q, p, d)
            k >>= 1
        acc = _jac_add(acc, R, p, d)
    return jac_to_affine(acc, p)


# ---------------------------------------------------------------- reference C-struct byte layouts
def fe_to_bytes(a: int, f: FieldParams, mont=True) -> bytes:
    v = f.to_mont(a) if mont else a
    return v.to_bytes(f.nbytes, "little")


def fe_from_bytes(b: bytes, f: FieldParams, mont=True) -> int:
    v = int.from_bytes(b, "little")
    return f.from_mont(v) if mont else v


def coord_to_bytes(c, f: FieldParams) -> bytes:
    return b"".join(fe_to_bytes(x, f) for x in c)


def aff_to_bytes(P, curve: CurveParams) -> bytes:
    """EC_ShortW_Aff {x, y}; infinity is (0, 0) (reference ec_shortweierstrass_affine.nim:52-62)."""
    if P is None:
        return bytes(curve.aff_bytes)
    return coord_to_bytes(P[0], curve.fp) + coord_to_bytes(P[1], curve.fp)


def coord_from_bytes(b: bytes, curve: CurveParams):
    n = curve.fp.nbytes
    return tuple(fe_from_bytes(b[i * n:(i + 1) * n], curve.fp) for i in range(curve.ext_degree))


def aff_from_bytes(b: bytes, curve: CurveParams):
    cb = curve.coord_bytes
    x = coord_from_bytes(b[:cb], curve)
    y = coord_from_bytes(b[cb:2 * cb], curve)
    if f_is_zero(x) and f_is_zero(y):
        return None
    return (x, y)


def jac_bytes_to_affine(b: bytes, curve: CurveParams):
    """EC_ShortW_Jac {x,y,z}: x = X/Z^2, y = Y/Z^3, infinity iff Z = 0 (reference ec_shortweierstrass_jacobian.nim:28-63)."""
    cb = curve.coord_bytes
    X, Y, Z = (coord_from_bytes(b[i * cb:(i + 1) * cb], curve) for i in range(3))
    return jac_to_affine((X, Y, Z), curve.fp.modulus)


def prj_bytes_to_affine(b: bytes, curve: CurveParams):
    """EC_ShortW_Prj {x,y,z}: x = X/Z, y = Y/Z, infinity iff Z = 0 (reference ec_shortweierstrass_projective.nim:28-62)."""
    p = curve.fp.modulus
    cb = curve.coord_bytes
    X, Y, Z = (coord_from_bytes(b[i * cb:(i + 1) * cb], curve) for i in range(3))
    if f_is_zero(Z):
        return None
    zi = f_inv(Z, p)
    return (f_mul(X, zi, p), f_mul(Y, zi, p))


def scalar_to_bytes(k: int, curve: CurveParams, fr_mont=False) -> bytes:
    """big_coefs: canonical BigInt[bits]; fr_coefs: Fr Montgomery residue (value must be < r)."""
    if fr_mont:
        return fe_to_bytes(k % curve.fr.modulus, curve.fr, mont=True)
    return k.to_bytes(curve.fr.nbytes, "little")


# ---------------------------------------------------------------- BLS12-381 G1 compressed encoding (ZCash / IETF format)
# used by the reference's KZG known-answer vectors (reference constantine/serialization/codecs_bls12_381.nim,
# tests/protocol_ethereum_eip4844_deneb_kzg/): 48 bytes big-endian x, flag bits in the top byte:
#   0x80 compressed, 0x40 infinity, 0x20 set iff y is the lexicographically larger root (y > (p-1)/2).
def bls12_381_g1_decompress(b: bytes, curve: CurveParams):
    assert len(b) == 48 and (b[0] & 0x80)
    if b[0] & 0x40:
        return None
    p = curve.fp.modulus
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    y2 = (x * x * x + curve.b[0]) % p
    y = pow(y2, (p + 1) // 4, p)          # p = 3 mod 4
    assert (y * y) % p == y2, "not on curve"
    if bool(b[0] & 0x20) != (y > (p - 1) // 2):
        y = p - y
    return ((x,), (y,))


def bls12_381_g1_compress(P, curve: CurveParams) -> bytes:
    if P is None:
        return bytes([0xC0]) + bytes(47)
    p = curve.fp.modulus
    x, y = P[0][0], P[1][0]
    raw = bytearray(x.to_bytes(48, "big"))
    raw[0] |= 0x80 | (0x20 if y > (p - 1) // 2 else 0)
    return bytes(raw)
