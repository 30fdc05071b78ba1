"""Host-side mirror of the reference's MSM interface on top of the C ABI.

Names follow the reference:
  multi_scalar_mul_vartime_parallel  <->  reference constantine/math/elliptic/ec_multi_scalar_mul_parallel.nim:588-628
      (`tp.multiScalarMul_vartime_parallel(r, coefs, points, len)`, BigInt and Fr overloads), exported to C as
      ctt_<curve>_<jac|prj>_multi_scalar_mul_<big|fr>_coefs_vartime_parallel (bindings/c_curve_decls_parallel.nim:31-45)
  multi_scalar_mul_vartime           <->  reference constantine/math/elliptic/ec_multi_scalar_mul.nim:525-568 (serial twin)
  Threadpool                          <->  reference constantine/threadpool/threadpool.nim:943-1041 (new / shutdown)
  PrecomputedMSM (init / msm_vartime) <->  reference constantine/math/elliptic/ec_multi_scalar_mul_precomp.nim:28-33, 109-161, 192-240
  PrecomputedMSMBank                  <->  the `openArray[PrecomputedMSM]` banks of reference constantine/math/matrix/toeplitz.nim:347-360
  sum_reduce_vartime(_parallel)       <->  reference constantine/math/elliptic/ec_shortweierstrass_batch_ops.nim:649-664,
                                           ec_shortweierstrass_batch_ops_parallel.nim:110-123

Buffers are `bytes`/`bytearray`/numpy arrays/anything exposing the buffer protocol, laid out exactly like the
reference's C structs (see include/ctt_b200_msm.h). Results come back as `bytes` of the _jac / _prj struct.
Every call goes through the named extern "C" symbol -- this module adds no arithmetic.
"""
import ctypes

from . import _lib
from .curves import CURVES, CurveParams

OUT_JAC, OUT_PRJ, OUT_XYZZ = 0, 1, 2


def _curve(curve) -> CurveParams:
    return curve if isinstance(curve, CurveParams) else CURVES[curve]


def _buf(b):
    """ctypes view of a read-only buffer without copying when possible."""
    if isinstance(b, (bytes, bytearray)):
        return (ctypes.c_char * len(b)).from_buffer_copy(b) if isinstance(b, bytes) else (ctypes.c_char * len(b)).from_buffer(b)
    mv = memoryview(b).cast("B")
    if mv.readonly:
        return (ctypes.c_char * len(mv)).from_buffer_copy(mv)
    return (ctypes.c_char * len(mv)).from_buffer(mv)


class Threadpool:
    """Opaque handle kept for source compatibility with reference callers (`Threadpool.new(n)` / `shutdown`)."""

    def __init__(self, num_threads: int = 0):
        lib = _lib.load()
        self._h = lib.ctt_threadpool_new(num_threads or lib.ctt_cpu_get_num_threads_os())

    @classmethod
    def new(cls, num_threads: int = 0):
        return cls(num_threads)

    def shutdown(self):
        if self._h:
            _lib.load().ctt_threadpool_shutdown(self._h)
            self._h = None


def set_devices(device_ids) -> None:
    """Spread every host-pointer MSM of this process over these CUDA devices (ctt_b200_set_devices; [] = caller's device only)."""
    ids = list(device_ids)
    arr = (ctypes.c_int * max(1, len(ids)))(*ids)
    if _lib.load().ctt_b200_set_devices(arr, len(ids)) != 0:
        raise ValueError(f"unknown CUDA device in {ids}")


def device_count() -> int:
    return _lib.load().ctt_b200_device_count()


def _symbol(curve: CurveParams, out: str, coef_kind: str, parallel: bool) -> str:
    return f"ctt_{curve.cprefix}_{out}_multi_scalar_mul_{coef_kind}_coefs_vartime" + ("_parallel" if parallel else "")


def multi_scalar_mul_vartime_parallel(tp, curve, coefs, points, length=None, out="jac", coef_kind="big") -> bytes:
    """r <- [a0]P0 + ... + [a_{n-1}]P_{n-1}  through ctt_<curve>_<out>_multi_scalar_mul_<coef_kind>_coefs_vartime_parallel.

    coefs: n x 32 bytes (canonical BigInt for coef_kind="big", Fr Montgomery residues for "fr");
    points: n affine structs. Returns the _jac / _prj struct bytes.
    """
    cv = _curve(curve)
    n = length if length is not None else len(memoryview(points).cast("B")) // cv.aff_bytes
    fn = _lib.named_msm(_symbol(cv, out, coef_kind, True))
    r = ctypes.create_string_buffer(cv.jac_bytes)
    cb, pb = _buf(coefs), _buf(points)
    fn(getattr(tp, "_h", None), r, cb, pb, n)
    return r.raw


def multi_scalar_mul_vartime(curve, coefs, points, length=None, out="jac", coef_kind="big") -> bytes:
    cv = _curve(curve)
    n = length if length is not None else len(memoryview(points).cast("B")) // cv.aff_bytes
    fn = _lib.named_msm(_symbol(cv, out, coef_kind, False))
    r = ctypes.create_string_buffer(cv.jac_bytes)
    cb, pb = _buf(coefs), _buf(points)
    fn(r, cb, pb, n)
    return r.raw


def msm_device_ptrs(curve, d_coefs: int, d_points: int, n: int, out=OUT_JAC, fr_mont=False, force_c=0,
                    win_begin=0, win_end=-1) -> bytes:
    """MSM over device-resident inputs (raw device pointers, e.g. torch tensor .data_ptr())."""
    cv = _curve(curve)
    size = cv.coord_bytes * (4 if out == OUT_XYZZ else 3)
    r = ctypes.create_string_buffer(size)
    rc = _lib.load().ctt_b200_msm_device(cv.curve_id, out, r, d_coefs, d_points, n, int(fr_mont), force_c, win_begin, win_end)
    if rc != 0:
        raise ValueError("ctt_b200_msm_device: bad curve id")
    return r.raw


def digits_per_window(c: int) -> int:
    """radix-16 digits of a window sum (k_plane_combine): ceil((c - 1) / 4)."""
    return (c - 1 + 3) // 4


def msm_device_digits(curve, d_digits_out: int, d_coefs: int, d_points: int, n: int, fr_mont=False, force_c=0, win_begin=0, win_end=-1) -> int:
    """Window range [win_begin, win_end) of an MSM over device-resident inputs; the radix-16 digits of the window sums stay in the
    device buffer d_digits_out (raw XYZZ, window-major). NOT synchronised: order later work on the engine's stream."""
    cv = _curve(curve)
    g = _lib.load().ctt_b200_msm_device_digits(cv.curve_id, d_digits_out, d_coefs, d_points, n, int(fr_mont), force_c, win_begin, win_end)
    if g < 0:
        raise ValueError("ctt_b200_msm_device_digits failed")
    return g


def combine_window_digits(curve, h_digits, c: int, num_windows: int, out=OUT_JAC) -> bytes:
    """digits of all windows 0..num_windows-1 (host buffer, window-major, digits_per_window(c) XYZZ points each) -> result struct."""
    cv = _curve(curve)
    r = ctypes.create_string_buffer(cv.coord_bytes * (4 if out == OUT_XYZZ else 3))
    if _lib.load().ctt_b200_combine_window_digits(cv.curve_id, out, r, _buf(h_digits), c, num_windows) != 0:
        raise ValueError("ctt_b200_combine_window_digits failed")
    return r.raw


def sum_partials(curve, partials: bytes, count: int, out=OUT_JAC) -> bytes:
    cv = _curve(curve)
    size = cv.coord_bytes * (4 if out == OUT_XYZZ else 3)
    r = ctypes.create_string_buffer(size)
    rc = _lib.load().ctt_b200_sum_partials(cv.curve_id, out, r, _buf(partials), count)
    if rc != 0:
        raise ValueError("ctt_b200_sum_partials: bad curve id")
    return r.raw


def plan(curve, n: int, force_c: int = 0):
    cv = _curve(curve)
    c, w = ctypes.c_int(0), ctypes.c_int(0)
    _lib.load().ctt_b200_plan(cv.curve_id, n, force_c, ctypes.byref(c), ctypes.byref(w))
    return c.value, w.value


def last_stats() -> dict:
    s = _lib.Stats()
    _lib.load().ctt_b200_last_stats(ctypes.byref(s))
    return {k: getattr(s, k) for k, _ in s._fields_}


class CachedBases:
    """Device-resident bases (reference constantine-rust/constantine-halo2-zal/src/lib.rs:58-95 caching hooks)."""

    def __init__(self, curve, points, length=None):
        self.curve = _curve(curve)
        self.n = length if length is not None else len(memoryview(points).cast("B")) // self.curve.aff_bytes
        self._h = _lib.load().ctt_b200_bases_upload(self.curve.curve_id, _buf(points), self.n)

    def precompute(self, c: int = 0, msm_len: int = 0) -> int:
        """One-time table of window multiples 2^(c w) P_i (ctt_b200_bases_precompute[_for]); returns the window size
        used. msm_len: length of the MSMs the bases will serve when they hold a whole bank (default: all bases)."""
        lib = _lib.load()
        rc = lib.ctt_b200_bases_precompute_for(self._h, msm_len, c) if msm_len else lib.ctt_b200_bases_precompute(self._h, c)
        if rc < 0:
            raise ValueError("ctt_b200_bases_precompute failed")
        return rc

    def msm_batch(self, coefs, batch: int, length: int, out=OUT_JAC, coef_kind="big", shared_points=False) -> list:
        """`batch` MSMs of `length` terms in one engine pass over the cached bases (ctt_b200_msm_batch_cached_bases)."""
        size = self.curve.coord_bytes * (4 if out == OUT_XYZZ else 3)
        r = ctypes.create_string_buffer(max(1, size * batch))
        rc = _lib.load().ctt_b200_msm_batch_cached_bases(self._h, out, r, _buf(coefs), batch, length, int(coef_kind == "fr"),
                                                         int(shared_points))
        if rc != 0:
            raise ValueError("ctt_b200_msm_batch_cached_bases failed (batch*len exceeds the cached bases?)")
        return [r.raw[m * size:(m + 1) * size] for m in range(batch)]

    def msm(self, coefs, length=None, out=OUT_JAC, coef_kind="big") -> bytes:
        n = length if length is not None else len(memoryview(coefs).cast("B")) // 32
        if out not in (OUT_JAC, OUT_PRJ, OUT_XYZZ):
            raise ValueError("out must be OUT_JAC, OUT_PRJ or OUT_XYZZ")
        r = ctypes.create_string_buffer(self.curve.coord_bytes * (4 if out == OUT_XYZZ else 3))
        rc = _lib.load().ctt_b200_msm_cached_bases(self._h, out, r, _buf(coefs), n, int(coef_kind == "fr"))
        if rc != 0:
            raise ValueError("ctt_b200_msm_cached_bases failed (len exceeds the cached bases?)")
        return r.raw

    def free(self):
        if self._h:
            _lib.load().ctt_b200_bases_free(self._h)
            self._h = None


def msm_batch(curve, coefs, points, batch: int, length: int, out=OUT_JAC, coef_kind="big", shared_points=False) -> list:
    """r[m] = sum_i coefs[m*length + i] * points[(0 if shared_points else m*length) + i] for m < batch, host buffers, one
    engine pass (ctt_b200_msm_batch_host). Returns the list of result structs."""
    cv = _curve(curve)
    size = cv.coord_bytes * (4 if out == OUT_XYZZ else 3)
    r = ctypes.create_string_buffer(max(1, size * batch))
    rc = _lib.load().ctt_b200_msm_batch_host(cv.curve_id, out, r, _buf(coefs), _buf(points), batch, length, int(coef_kind == "fr"),
                                             int(shared_points))
    if rc != 0:
        raise ValueError("ctt_b200_msm_batch_host: bad curve id")
    return [r.raw[m * size:(m + 1) * size] for m in range(batch)]


class PrecomputedMSM:
    """Fixed-base MSM context, the reference's `PrecomputedMSM[EC, N]`.

    reference: `ctx.init(basis, t, b)` builds comb tables (stride t, window b) on the host and `ctx.msm_vartime(r, scalars)`
    walks them with mixed additions. Here `init` uploads the basis and builds the table 2^(c w) P_i in HBM and
    `msm_vartime` is one engine pass with a single bucket set. (t, b) are accepted for source compatibility; they
    parametrise the reference's table shape, not the result -- the window c plays their role and is chosen by the
    engine unless given.
    """

    def __init__(self):
        self._bases = None
        self.N = 0

    def init(self, curve, basis, t: int = 0, b: int = 0, c: int = 0):
        self._bases = CachedBases(curve, basis)
        self.N = self._bases.n
        self.c = self._bases.precompute(c)
        return self

    def msm_vartime(self, scalars, out="jac", coef_kind="big") -> bytes:
        n = len(memoryview(scalars).cast("B")) // 32
        if n != self.N:
            raise ValueError("PrecomputedMSM.msm_vartime: need exactly N scalars")     # reference: openArray of length N
        return self._bases.msm(scalars, n, OUT_JAC if out == "jac" else OUT_PRJ, coef_kind)

    def free(self):
        if self._bases:
            self._bases.free()
            self._bases = None


class PrecomputedMSMBank:
    """`count` PrecomputedMSM objects of N points each, evaluated together (the reference loops over the bank,
    matrix/toeplitz.nim:357-360: `polyphaseSpectrumBank[i].msm_vartime(output[i], scalars_i)`)."""

    def __init__(self, curve, bases, count: int, n: int, c: int = 0):
        self.count, self.N = count, n
        self._bases = CachedBases(curve, bases, count * n)
        self.c = self._bases.precompute(c, msm_len=n)

    def msm_vartime(self, scalars, out="jac", coef_kind="big") -> list:
        return self._bases.msm_batch(scalars, self.count, self.N, OUT_JAC if out == "jac" else OUT_PRJ, coef_kind)

    def free(self):
        self._bases.free()


def sum_reduce_vartime(curve, points, length=None, out="jac") -> bytes:
    """r = P_0 + ... + P_{n-1} (ctt_b200_sum_reduce_host)."""
    cv = _curve(curve)
    n = length if length is not None else len(memoryview(points).cast("B")) // cv.aff_bytes
    r = ctypes.create_string_buffer(cv.jac_bytes)
    rc = _lib.load().ctt_b200_sum_reduce_host(cv.curve_id, OUT_JAC if out == "jac" else OUT_PRJ, r, _buf(points), n)
    if rc != 0:
        raise ValueError("ctt_b200_sum_reduce_host: bad curve id")
    return r.raw


def sum_reduce_vartime_parallel(tp, curve, points, length=None, out="jac") -> bytes:
    return sum_reduce_vartime(curve, points, length, out)


EVM_STATUS = ("cttEVM_Success", "cttEVM_InvalidInputSize", "cttEVM_InvalidOutputSize", "cttEVM_IntLargerThanModulus",
              "cttEVM_PointNotOnCurve", "cttEVM_PointNotInSubgroup", "cttEVM_VerificationFailure")


def eth_evm_bls12381_g1msm(inputs: bytes, out_len: int = 128):
    """EIP-2537 BLS12_G1MSM through ctt_eth_evm_bls12381_g1msm (reference constantine/ethereum_evm_precompiles.nim:894-975).
    Returns (status name, output bytes)."""
    r = ctypes.create_string_buffer(out_len)
    st = _lib.load().ctt_eth_evm_bls12381_g1msm(r, out_len, bytes(inputs), len(inputs))
    return EVM_STATUS[st], r.raw


def eth_evm_bls12381_g2msm(inputs: bytes, out_len: int = 256):
    """EIP-2537 BLS12_G2MSM through ctt_eth_evm_bls12381_g2msm (reference constantine/ethereum_evm_precompiles.nim:977-1060)."""
    r = ctypes.create_string_buffer(out_len)
    st = _lib.load().ctt_eth_evm_bls12381_g2msm(r, out_len, bytes(inputs), len(inputs))
    return EVM_STATUS[st], r.raw


class EthKzgContext:
    """EIP-4844 commitment context on the resident SRS, the role of the reference's EthereumKZGContext for
    blob_to_kzg_commitment[_parallel] (reference constantine/ethereum_eip4844_kzg_parallel.nim:125-159)."""

    ScalarLargerThanCurveOrder = 4        # cttEthKzg_ScalarLargerThanCurveOrder

    def __init__(self, srs_lagrange_brp_g1, compressed=True):
        lib = _lib.load()
        buf = _buf(srs_lagrange_brp_g1)
        if compressed:
            st = ctypes.c_int(0)
            self._h = lib.ctt_b200_eth_kzg_context_new_compressed(buf, ctypes.byref(st))
            if not self._h:
                raise ValueError(f"trusted setup point does not decode (cttEthKzg status {st.value})")
        else:
            self._h = lib.ctt_b200_eth_kzg_context_new(buf)
            if not self._h:
                raise ValueError("ctt_b200_eth_kzg_context_new failed")

    def precompute(self, c: int = 0) -> int:
        return _lib.load().ctt_b200_eth_kzg_context_precompute(self._h, c)

    def blob_to_kzg_commitment(self, blob) -> bytes:
        """48-byte compressed commitment; raises ValueError carrying the reference's status code for an invalid blob."""
        dst = ctypes.create_string_buffer(48)
        rc = _lib.load().ctt_b200_eth_kzg_blob_to_kzg_commitment(self._h, dst, _buf(blob))
        if rc != 0:
            raise ValueError(rc)
        return dst.raw

    def delete(self):
        if self._h:
            _lib.load().ctt_b200_eth_kzg_context_delete(self._h)
            self._h = None
