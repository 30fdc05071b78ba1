"""ORACLE (test infrastructure) -- ctypes binding of oracle/_build/libmsm_oracle.so (oracle/msm_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from constantine_b200.curves import CurveParams, FieldParams  # noqa: E402

LIB_PATH = os.path.join(_HERE, "_build", "libmsm_oracle.so")
_lib = None


class FieldT(ctypes.Structure):
    _fields_ = [("nl", ctypes.c_int), ("bits", ctypes.c_int), ("p", ctypes.c_uint64 * 6), ("one", ctypes.c_uint64 * 6),
                ("r2", ctypes.c_uint64 * 6), ("m0ninv", ctypes.c_uint64)]


class CurveT(ctypes.Structure):
    _fields_ = [("fp", FieldT), ("fr", FieldT), ("ext", ctypes.c_int), ("scalar_bits", ctypes.c_int)]


def _limbs(v, n=6):
    return (ctypes.c_uint64 * 6)(*[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)])


def field_t(f: FieldParams) -> FieldT:
    return FieldT(f.limbs64, f.bits, _limbs(f.modulus), _limbs(f.one_mont), _limbs(f.r2), f.m0ninv64)


def curve_t(c: CurveParams) -> CurveT:
    return CurveT(field_t(c.fp), field_t(c.fr), c.ext_degree, c.scalar_bits)


def build(force=False):
    src = [os.path.join(_HERE, "msm_oracle.c"), os.path.join(_HERE, "msm_oracle_impl.h")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src if os.path.exists(s))):
        return LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        lib.oracle_msm.argtypes = [ctypes.POINTER(CurveT), vp, vp, vp, sz, ci, ci, ci, ci]
        lib.oracle_msm.restype = ci
        lib.oracle_best_bucket_bit_size.argtypes = [ctypes.c_long, ci, ci, ci]
        lib.oracle_best_bucket_bit_size.restype = ci
        lib.oracle_parallel_dispatch_c.argtypes = [ctypes.c_long, ci]
        lib.oracle_parallel_dispatch_c.restype = ci
        lib.oracle_fp_op.argtypes = [ctypes.POINTER(FieldT), ci, vp, vp, vp, sz]
        lib.oracle_fp_op.restype = ci
        lib.oracle_ec_op.argtypes = [ctypes.POINTER(CurveT), ci, vp, vp, vp]
        lib.oracle_ec_op.restype = ci
        lib.oracle_signed_digit.argtypes = [vp, ci, ci, ci, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ci)]
        lib.oracle_signed_digit.restype = ci
        _lib = lib
    return _lib


IMPL_NAIVE, IMPL_REFERENCE, IMPL_SIGNED, IMPL_SIGNED_AFFINE = 0, 1, 2, 3


def msm(curve: CurveParams, coefs: bytes, points: bytes, n: int, fr_mont=False, impl=IMPL_SIGNED, c=0, nthreads=0) -> bytes:
    """Returns the Jacobian result struct bytes (reference EC_ShortW_Jac layout)."""
    lib = load()
    cv = curve_t(curve)
    out = ctypes.create_string_buffer(curve.jac_bytes)
    cb = ctypes.create_string_buffer(bytes(coefs), len(coefs)) if not isinstance(coefs, ctypes.Array) else coefs
    pb = ctypes.create_string_buffer(bytes(points), len(points)) if not isinstance(points, ctypes.Array) else points
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    rc = lib.oracle_msm(ctypes.byref(cv), out, cb, pb, n, int(fr_mont), impl, c, nthreads)
    if rc < 0:
        raise RuntimeError("oracle_msm failed")
    return out.raw


def fp_op(f: FieldParams, op: int, a: bytes, b: bytes, count: int) -> bytes:
    lib = load()
    ft = field_t(f)
    out = ctypes.create_string_buffer(len(a))
    lib.oracle_fp_op(ctypes.byref(ft), op, out, a, b, count)
    return out.raw


def ec_op(curve: CurveParams, op: int, p: bytes, q: bytes) -> bytes:
    lib = load()
    cv = curve_t(curve)
    out = ctypes.create_string_buffer(curve.jac_bytes)
    lib.oracle_ec_op(ctypes.byref(cv), op, out, p, q)
    return out.raw


def signed_digit(k: int, bits: int, c: int, w: int):
    lib = load()
    kb = k.to_bytes(32, "little")
    val, neg = ctypes.c_uint64(0), ctypes.c_int(0)
    rc = lib.oracle_signed_digit(kb, bits, c, w, ctypes.byref(val), ctypes.byref(neg))
    if rc != 0:
        raise ValueError("bad window")
    return val.value, neg.value
