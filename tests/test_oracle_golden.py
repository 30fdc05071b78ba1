"""CPU: pin the oracle (C restatement + exact tier) against the reference's own golden vectors and tables."""
import random

import pytest

from helpers import CURVES, case_inputs, pack, pyref


@pytest.mark.parametrize("impl", [0, 1, 2, 3])
def test_oracle_eip2537_vectors(kat, oracle_lib, impl):
    """reference tests/protocol_ethereum_evm_precompiles/eip-2537/multiexp_G{1,2}_bls.json -- byte-pinned MSM results."""
    assert len(kat["eip2537"]) == 27
    for case in kat["eip2537"]:
        cv, ks, pts, want = case_inputs(case)
        cb, pb = pack(cv, ks, pts)
        got = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, len(ks), impl=impl), cv)
        assert got == want, case["name"]


@pytest.mark.parametrize("impl", [1, 2, 3])
def test_oracle_sage_vector_sums(kat, oracle_lib, impl):
    """reference tests/math_elliptic_curves/vectors/tv_*_scalar_mul_*.json: each file as one 40-term MSM."""
    assert len(kat["sage_scalar_mul"]) == 18
    for case in kat["sage_scalar_mul"]:
        cv, ks, pts, want = case_inputs(case)
        cb, pb = pack(cv, ks, pts)
        got = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, len(ks), impl=impl), cv)
        assert got == want, case["name"]


def test_oracle_sage_single_products(kat, oracle_lib):
    """every ([k]P = Q) vector individually, as a 1-term MSM (N = 1 edge case) through the signed-window path"""
    for case in kat["sage_scalar_mul"]:
        cv, ks, pts, _ = case_inputs(case)
        for k, P, Q in list(zip(ks, pts, case["products"]))[:10]:
            cb, pb = pack(cv, [k], [P])
            got = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, 1), cv)
            want = (tuple(int(c, 16) for c in Q[0]), tuple(int(c, 16) for c in Q[1]))
            assert got == want


def test_oracle_fr_coefs_entry(kat, oracle_lib):
    """fr_coefs entry: Montgomery-form scalars are converted with one Montgomery reduction
    (reference ec_multi_scalar_mul_parallel.nim:611-628)."""
    for case in kat["eip2537"][:6] + kat["sage_scalar_mul"][::4]:
        cv, ks, pts, want = case_inputs(case)
        ks = [k % cv.fr.modulus for k in ks]
        cb, pb = pack(cv, ks, pts, fr_mont=True)
        got = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, len(ks), fr_mont=True), cv)
        assert got == pyref.msm_naive_fast(ks, pts, cv)


@pytest.mark.parametrize("bits,c", [(255, 13), (255, 15), (255, 16), (255, 12), (255, 5), (254, 7), (254, 2), (255, 17), (255, 3), (254, 8)])
def test_signed_recoding_reconstructs_scalar(oracle_lib, bits, c):
    """sum_w d_w 2^(w c) == k, 0 <= |d_w| <= 2^(c-1)  (reference bigints.nim:806-861)."""
    r = random.Random(bits * 100 + c)
    specials = [0, 1, (1 << bits) - 1, 1 << (bits - 1), (1 << (bits - 1)) - 1]
    for k in specials + [r.getrandbits(bits) for _ in range(300)]:
        total = 0
        for w in range(bits // c + 1):
            val, neg = oracle_lib.signed_digit(k, bits, c, w)
            assert 0 <= val <= 1 << (c - 1)
            total += (-val if neg else val) << (w * c)
        assert total == k


def test_best_bucket_bit_size_table(oracle_lib):
    """SURVEY.md Appendix C, recomputed from reference ec_multi_scalar_mul_scheduler.nim:172-223 and the
    parallel dispatch remap ec_multi_scalar_mul_parallel.nim:519-553."""
    lib = oracle_lib.load()
    table = {(256, 254): (7, 7), (1 << 16, 255): (13, 12), (1 << 18, 255): (14, 13), (1 << 20, 255): (15, 14), (1 << 22, 255): (16, 15)}
    for (n, bits), (best, par) in table.items():
        assert lib.oracle_best_bucket_bit_size(n, bits, 1, 1) == best
        assert lib.oracle_parallel_dispatch_c(n, bits) == par


def test_oracle_field_ops_vs_python_ints(oracle_lib):
    from constantine_b200.curves import FIELDS
    r = random.Random(5)
    for f in FIELDS.values():
        n = 64
        a = [r.randrange(f.modulus) for _ in range(n)]
        b = [r.randrange(f.modulus) for _ in range(n)]
        a[0], b[0] = 0, 0
        a[1], b[1] = f.modulus - 1, f.modulus - 1
        ab = b"".join(pyref.fe_to_bytes(x, f) for x in a)
        bb = b"".join(pyref.fe_to_bytes(x, f) for x in b)
        for op, fn in ((0, lambda x, y: x * y), (1, lambda x, y: x + y), (2, lambda x, y: x - y), (3, lambda x, y: -x),
                       (4, lambda x, y: x * pow(2, -1, f.modulus))):
            out = oracle_lib.fp_op(f, op, ab, bb, n)
            for i in range(n):
                got = pyref.fe_from_bytes(out[i * f.nbytes:(i + 1) * f.nbytes], f)
                assert got == fn(a[i], b[i]) % f.modulus, (f.name, op, i)


def test_oracle_point_ops_special_cases(oracle_lib):
    """sum_vartime / double / mixedSum_vartime incl. P+P, P-P, infinity operands
    (reference tests t_ec_shortw_jac_g1_add_double.nim, ..._mixed_add.nim model)."""
    for cv in CURVES.values():
        P = pyref.ec_mul_fast(5, cv.gen, cv)
        Q = pyref.ec_mul_fast(11, cv.gen, cv)
        cases = [(P, Q), (P, P), (P, pyref.ec_neg(P, cv)), (None, Q), (P, None), (None, None)]
        d = cv.ext_degree

        def jac_bytes(A):
            X, Y, Z = pyref.jac_from_affine(A, d)
            return pyref.coord_to_bytes(X, cv.fp) + pyref.coord_to_bytes(Y, cv.fp) + pyref.coord_to_bytes(Z, cv.fp)

        for A, B in cases:
            assert pyref.jac_bytes_to_affine(oracle_lib.ec_op(cv, 0, jac_bytes(A), jac_bytes(B)), cv) == pyref.ec_add(A, B, cv)
            assert pyref.jac_bytes_to_affine(oracle_lib.ec_op(cv, 2, jac_bytes(A), pyref.aff_to_bytes(B, cv)), cv) == pyref.ec_add(A, B, cv)
            assert pyref.jac_bytes_to_affine(oracle_lib.ec_op(cv, 1, jac_bytes(A), jac_bytes(A)), cv) == pyref.ec_add(A, A, cv)
        # non-trivial Z on both sides
        twoP = oracle_lib.ec_op(cv, 1, jac_bytes(P), jac_bytes(P))
        twoQ = oracle_lib.ec_op(cv, 1, jac_bytes(Q), jac_bytes(Q))
        assert pyref.jac_bytes_to_affine(oracle_lib.ec_op(cv, 0, twoP, twoQ), cv) == pyref.ec_mul_fast(32, cv.gen, cv)
        assert pyref.jac_bytes_to_affine(oracle_lib.ec_op(cv, 0, twoP, twoP), cv) == pyref.ec_mul_fast(20, cv.gen, cv)
