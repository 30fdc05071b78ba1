"""CPU: the C restatement against the exact big-int tier on seeded random inputs and the reference's edge cases
(SURVEY.md Appendix E)."""
import pytest

from helpers import CURVES, pack, point_pool, pyref


@pytest.mark.parametrize("curve", list(CURVES))
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 16, 33, 128])
def test_small_sizes(oracle_lib, rng, curve, n):
    """N in {1..8,16,32,64,128,...} like reference tests/parallel/t_ec_shortw_jac_g1_msm_parallel.nim:17-29;
    scalars uniform over the full declared width, i.e. possibly >= the group order (t_ec_template_parallel.nim:176)."""
    cv = CURVES[curve]
    _, pool = point_pool(cv)
    pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
    ks = [rng.getrandbits(cv.scalar_bits) for _ in range(n)]
    cb, pb = pack(cv, ks, pts)
    want = pyref.msm_naive_fast(ks, pts, cv)
    for impl in (1, 2):
        assert pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n, impl=impl), cv) == want


@pytest.mark.parametrize("curve", list(CURVES))
def test_edge_cases(oracle_lib, rng, curve):
    cv = CURVES[curve]
    _, pool = point_pool(cv)
    P, Q = pool[0], pool[1]
    r = cv.fr.modulus
    top = (1 << cv.scalar_bits) - 1
    cases = {
        "zero scalar": ([0], [P]),
        "one": ([1], [P]),
        "infinity point": ([rng.getrandbits(200)], [None]),
        "inf + inf": ([3, 4], [None, None]),
        "P + P same bucket": ([7, 7], [P, P]),
        "P - P same bucket": ([9, 9], [P, pyref.ec_neg(P, cv)]),
        "all equal points (bug-366 shape)": ([rng.getrandbits(cv.scalar_bits) for _ in range(40)], [P] * 40),
        "scalars >= r": ([r, r + 1, top], [P, Q, P]),
        "all-ones scalar": ([top], [Q]),
        "long 0/1 runs": ([int("1" * 100 + "0" * 60 + "1" * 90, 2), int("10" * 120, 2)], [P, Q]),
        "mixed": ([0, 1, 2, r - 1], [P, None, Q, Q]),
    }
    for name, (ks, pts) in cases.items():
        cb, pb = pack(cv, ks, pts)
        want = pyref.msm_naive_fast(ks, pts, cv)
        for impl in (0, 1, 2):
            assert pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, len(ks), impl=impl), cv) == want, (name, impl)


@pytest.mark.parametrize("c", [2, 3, 5, 8, 11, 15, 16, 17])
def test_every_window_size_gives_the_same_point(oracle_lib, rng, c):
    """incl. c dividing the bit width exactly (15, 17, 3, 5 | 255): the extra top window matters
    (reference ec_multi_scalar_mul_parallel.nim:157-158,186-190; bug-366 regression t_ec_shortw_jac_g2_msm_bug_366.nim)."""
    cv = CURVES["bls12_381_g1"]
    _, pool = point_pool(cv)
    n = 24
    pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
    ks = [rng.getrandbits(255) | (1 << 254) for _ in range(n)]
    cb, pb = pack(cv, ks, pts)
    want = pyref.msm_naive_fast(ks, pts, cv)
    assert pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n, impl=2, c=c), cv) == want


def test_threads_do_not_change_the_result(oracle_lib, rng):
    cv = CURVES["bn254_snarks_g1"]
    _, pool = point_pool(cv)
    n = 200
    pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
    ks = [rng.getrandbits(254) for _ in range(n)]
    cb, pb = pack(cv, ks, pts)
    a = oracle_lib.msm(cv, cb, pb, n, nthreads=1)
    b = oracle_lib.msm(cv, cb, pb, n, nthreads=8)
    assert a == b


def test_batched_affine_cpu_path_equals_the_checker(oracle_lib, rng):
    """impl 3 (the TIMED CPU baseline: signed windows + sort-based batched-affine bucket sums, the reference's arithmetic for
    c >= 9) returns the same group element as impl 2 (Jacobian buckets, the checker) on every group -- few distinct points so that
    P + P and P - P land inside the batches, an infinity input, a zero scalar, N = 1 and 2, several window sizes."""
    for curve, n in (("bls12_381_g1", 3000), ("bn254_snarks_g1", 777), ("bls12_381_g2", 300), ("pallas_ec", 50), ("vesta_ec", 1),
                     ("bn254_snarks_g2", 2)):
        cv = CURVES[curve]
        _, pool = point_pool(cv)
        pts = [pool[rng.randrange(6)] for _ in range(n)]
        ks = [rng.getrandbits(cv.scalar_bits) for _ in range(n)]
        if n > 10:
            pts[3] = None
            ks[5] = 0
            ks[6] = ks[7]
            pts[6] = pts[7]
        cb, pb = pack(cv, ks, pts)
        want = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n), cv)
        for c in (0, 2, 4, 9, 13):
            got = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n, impl=oracle_lib.IMPL_SIGNED_AFFINE, c=c), cv)
            assert got == want, (curve, c)
