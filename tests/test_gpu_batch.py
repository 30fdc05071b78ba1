"""GPU (-m gpu): batches of independent MSMs, fixed-base (precomputed) MSM and the sum of points
through the C ABI, against the oracle. Shapes follow the reference's own tests:
  tests/math_elliptic_curves/t_ec_multi_scalar_mul_precomp.nim:24-75   N = 4, 256, 128, 2, and N = 1 with the scalar 25
  tests/parallel/t_ec_template_parallel.nim:84-141                      sum reduction incl. P + P and P - P pairs
  constantine/math/matrix/toeplitz.nim:347-360                          a bank of PrecomputedMSM called one per output
Bit-exact bar: equality of the affine-normalised results."""
import random

import numpy as np
import pytest

from helpers import CURVES, pack, point_pool, pyref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from constantine_b200 import msm
    return msm


def _oracle_each(oracle_lib, cv, cb, pb, batch, n, shared):
    out = []
    for m in range(batch):
        c = cb[m * n * 32:(m + 1) * n * 32]
        p = pb[:n * cv.aff_bytes] if shared else pb[m * n * cv.aff_bytes:(m + 1) * n * cv.aff_bytes]
        out.append(pyref.jac_bytes_to_affine(oracle_lib.msm(cv, c, p, n), cv) if n else None)
    return out


@pytest.mark.parametrize("curve,batch,n", [("bls12_381_g1", 128, 64), ("bls12_381_g1", 5, 1), ("bls12_381_g1", 3, 1000),
                                           ("bn254_snarks_g1", 33, 17), ("pallas_ec", 2, 4097), ("bls12_381_g2", 16, 24),
                                           ("vesta_ec", 1, 50), ("bn254_snarks_g2", 7, 9)])
def test_batch_vs_oracle(M, oracle_lib, rng, curve, batch, n):
    """ctt_b200_msm_batch_host: every MSM of the batch equals the oracle's MSM of its slice (own bases per MSM)."""
    cv = CURVES[curve]
    _, pool = point_pool(cv)
    pts = [pool[rng.randrange(len(pool))] for _ in range(batch * n)]
    ks = [rng.getrandbits(cv.scalar_bits) for _ in range(batch * n)]
    if batch * n > 4:
        pts[3] = None            # infinity among the bases
        ks[2] = 0                # zero scalar
        ks[1] = (1 << cv.scalar_bits) - 1
    cb, pb = pack(cv, ks, pts)
    want = _oracle_each(oracle_lib, cv, cb, pb, batch, n, False)
    got = M.msm_batch(cv, cb, pb, batch, n)
    assert [pyref.jac_bytes_to_affine(g, cv) for g in got] == want
    got = M.msm_batch(cv, cb, pb, batch, n, out=M.OUT_PRJ)
    assert [pyref.prj_bytes_to_affine(g, cv) for g in got] == want


def test_batch_shared_bases_and_fr_coefs(M, oracle_lib, rng):
    """all MSMs of the batch over the same bases; fr_coefs (Montgomery) scalars; an MSM of all-zero scalars inside"""
    cv = CURVES["bls12_381_g1"]
    _, pool = point_pool(cv)
    batch, n = 9, 200
    pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
    ks = [rng.randrange(cv.fr.modulus) for _ in range(batch * n)]
    ks[4 * n:5 * n] = [0] * n
    cb, pb = pack(cv, ks, pts)
    want = _oracle_each(oracle_lib, cv, cb, pb, batch, n, True)
    assert want[4] is None
    got = M.msm_batch(cv, cb, pb, batch, n, shared_points=True)
    assert [pyref.jac_bytes_to_affine(g, cv) for g in got] == want
    cbm, _ = pack(cv, ks, [], fr_mont=True)
    got = M.msm_batch(cv, cbm, pb, batch, n, coef_kind="fr", shared_points=True)
    assert [pyref.jac_bytes_to_affine(g, cv) for g in got] == want


def test_batch_degenerate_shapes(M, rng):
    cv = CURVES["bn254_snarks_g1"]
    assert M.msm_batch(cv, b"", b"", 0, 5) == []
    got = M.msm_batch(cv, b"", b"", 3, 0)
    assert [pyref.jac_bytes_to_affine(g, cv) for g in got] == [None] * 3


@pytest.mark.parametrize("n,t,b,samples", [(4, 4, 3, 4), (256, 32, 12, 2), (128, 128, 12, 2), (2, 1, 2, 4)])
def test_precomputed_msm_reference_shapes(M, oracle_lib, rng, n, t, b, samples):
    """reference t_ec_multi_scalar_mul_precomp.nim:24-58 testConfig: PrecomputedMSM.init(basis, t, b); msm_vartime == MSM"""
    cv = CURVES["bls12_381_g1"]
    _, pool = point_pool(cv)
    pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
    _, pb = pack(cv, [], pts)
    ctx = M.PrecomputedMSM().init(cv, pb, t=t, b=b)
    for _ in range(samples):
        ks = [rng.getrandbits(255) for _ in range(n)]
        cb, _ = pack(cv, ks, [])
        want = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n), cv)
        assert pyref.jac_bytes_to_affine(ctx.msm_vartime(cb), cv) == want
    ctx.free()


def test_precomputed_msm_scalar_25(M):
    """reference t_ec_multi_scalar_mul_precomp.nim:60-72: N = 1, t = 1, b = 2, scalar 25 -> 25 G (anti-regression)"""
    cv = CURVES["bls12_381_g1"]
    cb, pb = pack(cv, [25], [cv.gen])
    ctx = M.PrecomputedMSM().init(cv, pb, t=1, b=2)
    assert pyref.jac_bytes_to_affine(ctx.msm_vartime(cb), cv) == pyref.ec_mul_fast(25, cv.gen, cv)
    ctx.free()


@pytest.mark.parametrize("curve,count,n,c", [("bls12_381_g1", 128, 64, 0), ("bls12_381_g1", 128, 64, 5), ("bls12_381_g1", 6, 300, 11),
                                             ("bn254_snarks_g1", 20, 33, 0), ("bls12_381_g2", 8, 16, 0)])
def test_precomputed_bank_vs_oracle(M, oracle_lib, rng, curve, count, n, c):
    """a bank of fixed-base MSMs in one pass over the window table (PeerDAS shape: 128 x 64, reference
    commitments_setups/ethereum_kzg_srs.nim:133) -- each output equals the oracle's MSM over that member's bases"""
    cv = CURVES[curve]
    _, pool = point_pool(cv)
    pts = [pool[rng.randrange(len(pool))] for _ in range(count * n)]
    pts[5] = None
    _, pb = pack(cv, [], pts)
    bank = M.PrecomputedMSMBank(cv, pb, count, n, c=c)
    assert 2 <= bank.c <= 20 and (c == 0 or bank.c == c)
    for trial in range(2):
        ks = [rng.getrandbits(cv.scalar_bits) for _ in range(count * n)]
        if trial:
            ks[:n] = [0] * n
        cb, _ = pack(cv, ks, [])
        want = _oracle_each(oracle_lib, cv, cb, pb, count, n, False)
        got = bank.msm_vartime(cb)
        assert [pyref.jac_bytes_to_affine(g, cv) for g in got] == want, (curve, count, n, c, trial)
    bank.free()


def test_batch_closed_form_large(M):
    """2048 MSMs x 512 points over shared bases P_i = [k_i]G: result m = [sum_i s_mi k_i mod r] G (exact, any size)"""
    from constantine_b200 import _lib
    lib = _lib.load()
    cv = CURVES["bls12_381_g1"]
    batch, n = 2048, 512
    r = np.random.default_rng(5)
    k = r.integers(1, 2**63, size=n, dtype=np.uint64)
    gen = b"".join(cv.fp.to_mont(c).to_bytes(cv.fp.nbytes, "little") for coord in cv.gen for c in coord)
    pts = np.empty((n, cv.aff_bytes), dtype=np.uint8)
    assert lib.ctt_b200_scalar_mul_u64(cv.curve_id, gen, k.ctypes.data, n, pts.ctypes.data) == 0
    s = r.integers(0, 256, size=(batch * n, 32), dtype=np.uint8)
    s[:, 31] &= 0x7F
    got = M.msm_batch(cv, s, pts, batch, n, shared_points=True)
    kk = [int(x) for x in k]
    for m in (0, 1, 777, batch - 1):
        sc = [int.from_bytes(s[m * n + i].tobytes(), "little") for i in range(n)]
        e = sum(a * b for a, b in zip(sc, kk)) % cv.fr.modulus
        assert pyref.jac_bytes_to_affine(got[m], cv) == pyref.ec_mul_fast(e, cv.gen, cv), m


@pytest.mark.parametrize("curve", list(CURVES))
def test_sum_reduce_vs_exact(M, rng, curve):
    """reference t_ec_template_parallel.nim:84-141: random points, then the special cases P + P and P - P"""
    cv = CURVES[curve]
    _, pool = point_pool(cv)
    for n in (1, 2, 10, 100, 257, 1500):
        pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
        half = n // 2
        for i in range(half, n):       # second half: fresh point / the same point again / its negation / infinity
            kind = rng.randrange(4)
            if kind == 1:
                pts[i] = pts[i - half]
            elif kind == 2:
                pts[i] = pyref.ec_neg(pts[i - half], cv)
            elif kind == 3 and n > 10:
                pts[i] = None
        want = None
        for P in pts:
            want = pyref.ec_add(want, P, cv)
        _, pb = pack(cv, [], pts)
        assert pyref.jac_bytes_to_affine(M.sum_reduce_vartime(cv, pb, n), cv) == want, (curve, n)
        assert pyref.prj_bytes_to_affine(M.sum_reduce_vartime_parallel(None, cv, pb, n, out="prj"), cv) == want, (curve, n)
    assert pyref.jac_bytes_to_affine(M.sum_reduce_vartime(cv, b"", 0), cv) is None


def test_sum_reduce_closed_form_large(M):
    """2^20 points [k_i]G: sum = [sum k_i mod r] G"""
    from constantine_b200 import _lib
    lib = _lib.load()
    cv = CURVES["bn254_snarks_g1"]
    n = 1 << 20
    k = np.random.default_rng(9).integers(1, 2**63, size=n, dtype=np.uint64)
    gen = b"".join(cv.fp.to_mont(c).to_bytes(cv.fp.nbytes, "little") for coord in cv.gen for c in coord)
    pts = np.empty((n, cv.aff_bytes), dtype=np.uint8)
    assert lib.ctt_b200_scalar_mul_u64(cv.curve_id, gen, k.ctypes.data, n, pts.ctypes.data) == 0
    e = sum(int(x) for x in k) % cv.fr.modulus
    assert pyref.jac_bytes_to_affine(M.sum_reduce_vartime(cv, pts, n), cv) == pyref.ec_mul_fast(e, cv.gen, cv)


# ------------------------------------------------------------------ experimental paths (default off in the library)
@pytest.mark.gpu
def test_safegcd_inversion_kernel(oracle_lib):
    """field_inv.cuh fe_inv_safegcd through the test hook (op 7: a * inv(a) must be one; zero for a = 0), every field"""
    import ctypes
    from constantine_b200 import _lib
    from constantine_b200.curves import FIELDS
    lib = _lib.load()
    for fid, name in enumerate(("bls12_381_fp", "bn254_snarks_fp", "pallas_fp", "vesta_fp", "bls12_381_fr", "bn254_snarks_fr",
                                "pallas_fr", "vesta_fr")):
        f = FIELDS[name]
        r = np.random.default_rng(fid)
        a = r.integers(0, 256, size=(8192, f.nbytes), dtype=np.uint8)
        a[:, -1] &= (1 << ((f.bits - 1) % 8)) - 1          # below 2^(bits-1) < p: canonical
        a[0, :] = 0
        a[1, :] = 0
        a[1, 0] = 1
        a[2, :] = np.frombuffer((f.modulus - 1).to_bytes(f.nbytes, "little"), dtype=np.uint8)
        a[3, :] = np.frombuffer(f.one_mont.to_bytes(f.nbytes, "little"), dtype=np.uint8)
        out = ctypes.create_string_buffer(a.nbytes)
        assert lib.ctt_b200_test_field_op(fid, 7, out, a.ctypes.data, a.ctypes.data, len(a)) == 0
        one = f.one_mont.to_bytes(f.nbytes, "little")
        got = out.raw
        assert got[:f.nbytes] == bytes(f.nbytes), name
        bad = [i for i in range(1, len(a)) if got[i * f.nbytes:(i + 1) * f.nbytes] != one]
        assert not bad, (name, bad[:5])


@pytest.mark.gpu
@pytest.mark.parametrize("levels", [1, 2, 3, 5])
def test_batched_affine_levels_same_results(M, oracle_lib, rng, levels):
    """ctt_b200_set_affine_levels: the leading levels of the bucket sums as batched-affine additions (msm_affine.cuh) give the
    same MSM as the oracle -- few distinct points, so P + P and P - P land inside the batches; an infinity input; N = 1"""
    from constantine_b200 import _lib
    lib = _lib.load()
    try:
        lib.ctt_b200_set_affine_levels(levels)
        for curve, n in (("bls12_381_g1", 5000), ("bn254_snarks_g1", 3001), ("bls12_381_g2", 700), ("pallas_ec", 1),
                         ("vesta_ec", 33), ("bn254_snarks_g2", 257)):
            cv = CURVES[curve]
            _, pool = point_pool(cv)
            pts = [pool[rng.randrange(12)] for _ in range(n)]
            if n > 10:
                pts[3] = None
            ks = [rng.getrandbits(cv.scalar_bits) for _ in range(n)]
            if n > 10:
                ks[5] = 0
                ks[6] = ks[7]
                pts[6] = pts[7]
            cb, pb = pack(cv, ks, pts)
            want = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n), cv)
            got = M.multi_scalar_mul_vartime(cv, cb, pb, n)
            assert pyref.jac_bytes_to_affine(got, cv) == want, (curve, levels)
            assert M.last_stats()["affine_levels"] == levels
    finally:
        lib.ctt_b200_set_affine_levels(-1)


@pytest.mark.gpu
def test_batched_affine_adversarial_runs(M):
    """all scalars equal (one run of N entries per window: log2 N levels would be needed, the XYZZ slices finish it) and all
    points equal (every level-0 pair is a doubling) -- closed forms"""
    from constantine_b200 import _lib
    lib = _lib.load()
    cv = CURVES["bls12_381_g1"]
    n = 1 << 14
    try:
        lib.ctt_b200_set_affine_levels(3)
        k = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF1234567890ABCDE
        g = pyref.aff_to_bytes(cv.gen, cv)
        got = M.multi_scalar_mul_vartime(cv, pyref.scalar_to_bytes(k, cv) * n, g * n, n)
        want = pyref.ec_mul_fast(k * n % cv.fr.modulus, cv.gen, cv)
        assert pyref.jac_bytes_to_affine(got, cv) == want
        r = random.Random(3)
        ks = [r.getrandbits(255) for _ in range(n)]
        got = M.multi_scalar_mul_vartime(cv, b"".join(pyref.scalar_to_bytes(x, cv) for x in ks), g * n, n)
        assert pyref.jac_bytes_to_affine(got, cv) == pyref.ec_mul_fast(sum(ks) % cv.fr.modulus, cv.gen, cv)
    finally:
        lib.ctt_b200_set_affine_levels(-1)
