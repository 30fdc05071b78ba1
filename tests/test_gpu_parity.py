"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle / golden vectors / closed forms.
Bit-exact bar: integer work -- equality of the affine-normalised result (the reference's own notion of equality,
tests/parallel/t_ec_template_parallel.nim:188)."""
import ctypes
import random

import numpy as np
import pytest

from helpers import CURVES, case_inputs, pack, point_pool, pyref, xyzz_bytes_to_affine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from constantine_b200 import msm
    return msm


@pytest.fixture(scope="module")
def lib():
    from constantine_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="module")
def tp(M):
    t = M.Threadpool.new(2)
    yield t
    t.shutdown()


# ------------------------------------------------------------------ kernels: field and point primitives
def test_field_kernels_vs_oracle(lib, oracle_lib):
    """fe_mul / add / sub / neg / dbl kernels on 2^16 random pairs per field + edge values, vs the oracle's C restatement
    (and through it the exact tier, tests/test_oracle_golden.py)."""
    from constantine_b200.curves import FIELDS
    ids = {"bls12_381_fp": 0, "bn254_snarks_fp": 1, "pallas_fp": 2, "vesta_fp": 3, "bls12_381_fr": 4, "bn254_snarks_fr": 5,
           "pallas_fr": 6, "vesta_fr": 7}
    r = random.Random(11)
    for name, fid in ids.items():
        f = FIELDS[name]
        n = 1 << 16
        edge = [0, 1, f.modulus - 1, f.modulus - 2, f.one_mont, (1 << (f.bits - 1)) % f.modulus]
        a = edge + [r.randrange(f.modulus) for _ in range(64)]
        b = list(reversed(edge)) + [r.randrange(f.modulus) for _ in range(64)]
        ab = b"".join(x.to_bytes(f.nbytes, "little") for x in a)
        bb = b"".join(x.to_bytes(f.nbytes, "little") for x in b)
        # bulk random residues straight from numpy (reduced by clearing the top bits: < 2^(bits-1) < p)
        rng = np.random.default_rng(fid)
        bulk_a = rng.integers(0, 256, size=(n, f.nbytes), dtype=np.uint8)
        bulk_b = rng.integers(0, 256, size=(n, f.nbytes), dtype=np.uint8)
        top = (f.bits - 1) % 8
        bulk_a[:, -1] &= (1 << top) - 1
        bulk_b[:, -1] &= (1 << top) - 1
        ab += bulk_a.tobytes()
        bb += bulk_b.tobytes()
        cnt = len(ab) // f.nbytes
        for op in range(5):
            out = ctypes.create_string_buffer(len(ab))
            assert lib.ctt_b200_test_field_op(fid, op, out, ab, bb, cnt) == 0
            want = oracle_lib.fp_op(f, op if op < 4 else 1, ab, bb if op < 4 else ab, cnt)
            assert out.raw == want, (name, op)
        # op 5: a b + (a + b)(a - b) through the two-product multiplier (one Montgomery reduction for both products)
        out = ctypes.create_string_buffer(len(ab))
        assert lib.ctt_b200_test_field_op(fid, 5, out, ab, bb, cnt) == 0
        t1 = oracle_lib.fp_op(f, 0, ab, bb, cnt)
        t2 = oracle_lib.fp_op(f, 0, oracle_lib.fp_op(f, 1, ab, bb, cnt), oracle_lib.fp_op(f, 2, ab, bb, cnt), cnt)
        assert out.raw == oracle_lib.fp_op(f, 1, t1, t2, cnt), (name, "dot2")
        # op 6: a^2 + b^2 through the dedicated squaring
        out = ctypes.create_string_buffer(len(ab))
        assert lib.ctt_b200_test_field_op(fid, 6, out, ab, bb, cnt) == 0
        assert out.raw == oracle_lib.fp_op(f, 1, oracle_lib.fp_op(f, 0, ab, ab, cnt), oracle_lib.fp_op(f, 0, bb, bb, cnt), cnt), (name, "sqr")


@pytest.mark.parametrize("curve", list(CURVES))
def test_point_kernels_vs_exact(lib, curve, rng):
    """XYZZ mixed add / add / double incl. P+P, P-P and infinity operands vs the exact tier."""
    cv = CURVES[curve]
    ks, pool = point_pool(cv)
    P = [pool[i] for i in range(24)] + [pool[0], pool[1], None, pool[2], None]
    Q = [pool[i + 24] for i in range(24)] + [pool[0], pyref.ec_neg(pool[1], cv), pool[3], None, None]
    n = len(P)
    pb = b"".join(pyref.aff_to_bytes(x, cv) for x in P)
    qb = b"".join(pyref.aff_to_bytes(x, cv) for x in Q)
    dbl = lambda A: pyref.ec_add(A, A, cv)
    expect = {0: lambda A, B: pyref.ec_add(A, B, cv), 1: lambda A, B: dbl(A), 2: lambda A, B: pyref.ec_add(dbl(A), dbl(B), cv),
              3: lambda A, B: pyref.ec_add(dbl(A), B, cv), 4: lambda A, B: dbl(dbl(A))}
    for op, fn in expect.items():
        out = ctypes.create_string_buffer(n * 4 * cv.coord_bytes)
        assert lib.ctt_b200_test_ec_op(cv.curve_id, op, out, pb, qb, n) == 0
        for i in range(n):
            got = xyzz_bytes_to_affine(out.raw[i * 4 * cv.coord_bytes:(i + 1) * 4 * cv.coord_bytes], cv)
            assert got == fn(P[i], Q[i]), (curve, op, i)


# ------------------------------------------------------------------ golden vectors through the reference's C symbols
def test_eip2537_vectors_through_c_abi(kat, M, tp):
    for case in kat["eip2537"]:
        cv, ks, pts, want = case_inputs(case)
        cb, pb = pack(cv, ks, pts)
        got = M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, len(ks), out="jac")
        assert pyref.jac_bytes_to_affine(got, cv) == want, case["name"]
        got = M.multi_scalar_mul_vartime(cv, cb, pb, len(ks), out="prj")
        assert pyref.prj_bytes_to_affine(got, cv) == want, case["name"]


def test_sage_vector_sums_through_c_abi(kat, M, tp):
    for case in kat["sage_scalar_mul"]:
        cv, ks, pts, want = case_inputs(case)
        cb, pb = pack(cv, ks, pts)
        assert pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, len(ks)), cv) == want, case["name"]
        cbm, _ = pack(cv, [k % cv.fr.modulus for k in ks], pts, fr_mont=True)
        got = M.multi_scalar_mul_vartime_parallel(tp, cv, cbm, pb, len(ks), out="prj", coef_kind="fr")
        assert pyref.prj_bytes_to_affine(got, cv) == want, case["name"]


# ------------------------------------------------------------------ seeded random inputs vs the oracle
@pytest.mark.parametrize("curve", list(CURVES))
def test_sizes_vs_oracle(M, tp, oracle_lib, curve):
    """N in {1..8, 16, 32, 64, 128, 1024, 2048, 16384} (reference tests/parallel/t_ec_shortw_jac_g1_msm_parallel.nim:17-29),
    scalars uniform over the declared width (>= r allowed)."""
    cv = CURVES[curve]
    r = random.Random(hash(curve) & 0xFFFF)
    _, pool = point_pool(cv)
    sizes = [1, 2, 3, 4, 5, 6, 7, 8, 16, 32, 64, 128, 1024, 2048] + ([16384] if cv.ext_degree == 1 else [4096])
    for n in sizes:
        pts = [pool[r.randrange(len(pool))] for _ in range(n)]
        ks = [r.getrandbits(cv.scalar_bits) for _ in range(n)]
        cb, pb = pack(cv, ks, pts)
        want = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n), cv)
        assert pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, n), cv) == want, (curve, n)


@pytest.mark.parametrize("curve", list(CURVES))
def test_edge_cases_vs_exact(M, tp, curve, rng):
    """SURVEY.md Appendix E."""
    cv = CURVES[curve]
    _, pool = point_pool(cv)
    P, Q = pool[0], pool[1]
    r = cv.fr.modulus
    top = (1 << cv.scalar_bits) - 1
    cases = {
        "len 0": ([], []),
        "zero scalar": ([0], [P]), "one": ([1], [P]), "infinity point": ([rng.getrandbits(200)], [None]),
        "inf + inf": ([3, 4], [None, None]), "P + P": ([7, 7], [P, P]), "P - P": ([9, 9], [P, pyref.ec_neg(P, cv)]),
        "all equal points": ([rng.getrandbits(cv.scalar_bits) for _ in range(300)], [P] * 300),
        "all equal points and scalars": ([12345] * 257, [Q] * 257),
        "scalars >= r": ([r, r + 1, top], [P, Q, P]), "all-ones": ([top], [Q]),
        "long 0/1 runs": ([int("1" * 100 + "0" * 60 + "1" * 90, 2), int("10" * 120, 2)] * 20, [P, Q] * 20),
        "cancels to infinity": ([5, 5, 11, r - 11], [P, pyref.ec_neg(P, cv), Q, Q]),
    }
    for name, (ks, pts) in cases.items():
        cb, pb = pack(cv, ks, pts)
        want = pyref.msm_naive_fast(ks, pts, cv)
        got = M.multi_scalar_mul_vartime_parallel(tp, cv, cb or b"\0" * 32, pb or bytes(cv.aff_bytes), len(ks))
        assert pyref.jac_bytes_to_affine(got, cv) == want, (curve, name)


def test_every_window_size_same_point(M, lib, rng):
    """forced c = 2..20 incl. c | 255 (extra top window) -- device-resident entry + window-range partial sums"""
    import torch
    cv = CURVES["bls12_381_g1"]
    _, pool = point_pool(cv)
    n = 1500
    pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
    ks = [rng.getrandbits(255) | (1 << 254) for _ in range(n)]
    cb, pb = pack(cv, ks, pts)
    want = pyref.msm_naive_fast(ks, pts, cv)
    d_c = torch.frombuffer(bytearray(cb), dtype=torch.uint8).cuda()
    d_p = torch.frombuffer(bytearray(pb), dtype=torch.uint8).cuda()
    torch.cuda.synchronize()
    for c in (2, 3, 5, 8, 11, 13, 15, 16, 17, 20):
        got = M.msm_device_ptrs(cv, d_c.data_ptr(), d_p.data_ptr(), n, force_c=c)
        assert pyref.jac_bytes_to_affine(got, cv) == want, c
        W = 255 // c + 1
        cut = W // 3
        parts = b"".join(M.msm_device_ptrs(cv, d_c.data_ptr(), d_p.data_ptr(), n, out=M.OUT_XYZZ, force_c=c, win_begin=a, win_end=b)
                         for a, b in ((0, cut), (cut, W - 1), (W - 1, W)))
        assert pyref.jac_bytes_to_affine(M.sum_partials(cv, parts, 3), cv) == want, ("window ranges", c)


def test_host_entry_raw_partial_output(M, lib, oracle_lib, rng):
    """ctt_b200_msm_host with CTT_B200_OUT_XYZZ (what a rank contributes to the multi-GPU all_gather): two half-MSMs
    combined by ctt_b200_sum_partials equal the whole MSM."""
    cv = CURVES["pallas_ec"]
    _, pool = point_pool(cv)
    n = 2000
    pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
    ks = [rng.getrandbits(255) for _ in range(n)]
    cb, pb = pack(cv, ks, pts)
    want = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n), cv)
    parts = b""
    for lo, hi in ((0, 900), (900, n)):
        r = ctypes.create_string_buffer(4 * cv.coord_bytes)
        assert lib.ctt_b200_msm_host(cv.curve_id, M.OUT_XYZZ, r, cb[32 * lo:32 * hi], pb[cv.aff_bytes * lo:cv.aff_bytes * hi], hi - lo, 0) == 0
        parts += r.raw
    assert pyref.jac_bytes_to_affine(M.sum_partials(cv, parts, 2), cv) == want


def test_cached_bases(M, oracle_lib, rng):
    """device-resident bases (the C face of the reference ZAL's base caching)"""
    cv = CURVES["bn254_snarks_g1"]
    _, pool = point_pool(cv)
    n = 3000
    pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
    _, pb = pack(cv, [], pts)
    bases = M.CachedBases(cv, pb, n)
    for trial in range(3):
        ks = [rng.getrandbits(254) for _ in range(n)]
        cb, _ = pack(cv, ks, [])
        want = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n), cv)
        assert pyref.jac_bytes_to_affine(bases.msm(cb, n), cv) == want
    bases.free()


# ------------------------------------------------------------------ BASELINE sizes: size-independent properties
def _gen_points(lib, cv, k):
    gen = b"".join(cv.fp.to_mont(c).to_bytes(cv.fp.nbytes, "little") for coord in cv.gen for c in coord)
    out = np.empty((len(k), cv.aff_bytes), dtype=np.uint8)
    assert lib.ctt_b200_scalar_mul_u64(cv.curve_id, gen, k.ctypes.data, len(k), out.ctypes.data) == 0
    return out


def test_generator_hook_vs_exact(lib):
    for cv in CURVES.values():
        k = np.array([1, 2, 3, 0xFFFFFFFFFFFFFFFF, 0x123456789ABCDEF, 0], dtype=np.uint64)
        pts = _gen_points(lib, cv, k)
        for i, kk in enumerate(k):
            assert pyref.aff_from_bytes(pts[i].tobytes(), cv) == pyref.ec_mul_fast(int(kk), cv.gen, cv)


@pytest.mark.parametrize("curve,logn", [("bn254_snarks_g1", 8), ("bls12_381_g1", 16), ("bls12_381_g1", 20), ("pallas_ec", 20),
                                        ("bls12_381_g2", 16), ("bls12_381_g2", 18)])
def test_closed_form_at_baseline_sizes(M, lib, tp, curve, logn):
    """points P_i = [k_i]G with known k_i  =>  MSM = [sum s_i k_i mod r] G  (the bug-366 construction generalised,
    SURVEY.md 8c item 4) -- exact at any N, here at BASELINE.json's sizes; plus linearity MSM(s+t) = MSM(s) + MSM(t)."""
    cv = CURVES[curve]
    n = 1 << logn
    rng = np.random.default_rng(logn * 7 + cv.curve_id)
    k = rng.integers(1, 2**63, size=n, dtype=np.uint64)
    pts = _gen_points(lib, cv, k)
    scal = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    scal[:, 31] &= (1 << (cv.scalar_bits - 248)) - 1
    s_int = [int.from_bytes(scal[i].tobytes(), "little") for i in range(n)]
    total = sum(s * int(kk) for s, kk in zip(s_int, k)) % cv.fr.modulus
    want = pyref.ec_mul_fast(total, cv.gen, cv)
    got = M.multi_scalar_mul_vartime_parallel(tp, cv, scal, pts, n)
    assert pyref.jac_bytes_to_affine(got, cv) == want
    # linearity on the same bases: MSM(s) + MSM(t) == MSM(s + t mod r)
    t_int = [int(x) for x in rng.integers(0, 2**62, size=n)]
    tb = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in t_int), dtype=np.uint8).reshape(n, 32)
    st = np.frombuffer(b"".join(((a + b) % cv.fr.modulus).to_bytes(32, "little") for a, b in zip(s_int, t_int)), dtype=np.uint8).reshape(n, 32)
    a = pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, tb, pts, n), cv)
    b = pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, st, pts, n), cv)
    assert pyref.ec_add(want, a, cv) == b


def test_adversarial_distribution_large(M, lib, tp):
    """N = 2^18 with every scalar equal (one bucket per window) and long-0/1-run scalars: the slice/fix-up scheme must
    stay exact (and finish) under maximal skew (reference helpers/prng_unsafe.nim:198-287 distributions)."""
    cv = CURVES["bls12_381_g1"]
    n = 1 << 18
    rng = np.random.default_rng(3)
    k = rng.integers(1, 2**63, size=n, dtype=np.uint64)
    pts = _gen_points(lib, cv, k)
    ksum = int(k.astype(object).sum())
    for s in (0x5555AAAA5555AAAA5555AAAA5555AAAA5555AAAA5555AAAA5555AAAA5555AAAA & ((1 << 255) - 1), int("1" * 120 + "0" * 70 + "1" * 65, 2), 1):
        scal = np.frombuffer(s.to_bytes(32, "little") * n, dtype=np.uint8).reshape(n, 32)
        want = pyref.ec_mul_fast((s * ksum) % cv.fr.modulus, cv.gen, cv)
        assert pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, scal, pts, n), cv) == want


def test_concurrent_callers(M, oracle_lib):
    """The reference allows nested / concurrent MSM calls (SURVEY.md 8b "Threading"); concurrent callers lease separate
    engine slots (own streams and scratch) or queue on a slot's mutex. Four Python threads hammer the C symbol with
    different inputs; every result must match its own oracle value."""
    import threading
    cv = CURVES["bn254_snarks_g1"]
    _, pool = point_pool(cv)
    jobs = []
    for seed in range(4):
        r = random.Random(seed)
        n = 500 + 137 * seed
        pts = [pool[r.randrange(len(pool))] for _ in range(n)]
        ks = [r.getrandbits(254) for _ in range(n)]
        cb, pb = pack(cv, ks, pts)
        jobs.append((cb, pb, n, pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n), cv)))
    errors = []

    def worker(job):
        cb, pb, n, want = job
        tp = M.Threadpool.new(1)
        for _ in range(5):
            got = pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, n), cv)
            if got != want:
                errors.append(n)
        tp.shutdown()

    threads = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors


def test_closed_form_2_22(M, lib, tp):
    """N = 2^22 (the largest BASELINE size), Pallas through the fr_coefs entry (config 4's symbol)."""
    cv = CURVES["pallas_ec"]
    n = 1 << 22
    rng = np.random.default_rng(22)
    k = rng.integers(1, 2**63, size=n, dtype=np.uint64)
    pts = _gen_points(lib, cv, k)
    r = cv.fr.modulus
    s_lo = rng.integers(0, 2**63, size=n, dtype=np.uint64)
    # scalars s_i = s_lo_i * 2^190 + 1 (mod r), passed as Fr Montgomery residues
    s_int = [((int(a) << 190) + 1) % r for a in s_lo]
    scal = np.frombuffer(b"".join(((s * cv.fr.R) % r).to_bytes(32, "little") for s in s_int), dtype=np.uint8).reshape(n, 32)
    total = sum(a * int(b) for a, b in zip(s_int, k)) % r
    want = pyref.ec_mul_fast(total, cv.gen, cv)
    got = M.multi_scalar_mul_vartime_parallel(tp, cv, scal, pts, n, out="prj", coef_kind="fr")
    assert pyref.prj_bytes_to_affine(got, cv) == want


def test_cached_bases_with_precomputed_table(M, oracle_lib, rng):
    """ctt_b200_bases_precompute: window multiples 2^(c w) P_i cached on the device, all windows share one bucket set.
    Same results as the plain path for every prefix length, both scalar encodings, several window sizes."""
    for curve, n in (("bls12_381_g1", 3000), ("bn254_snarks_g1", 4097), ("bls12_381_g2", 600)):
        cv = CURVES[curve]
        _, pool = point_pool(cv)
        pts = [pool[rng.randrange(len(pool))] for _ in range(n)]
        pts[7] = None
        _, pb = pack(cv, [], pts)
        bases = M.CachedBases(cv, pb, n)
        for c in (0, 8, 13):
            used = bases.precompute(c)
            assert 2 <= used <= 20
            for m in (n, n // 2, 1):
                ks = [rng.getrandbits(cv.scalar_bits) for _ in range(m)]
                cb, _ = pack(cv, ks, [])
                want = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb[:m * cv.aff_bytes], m), cv)
                assert pyref.jac_bytes_to_affine(bases.msm(cb, m), cv) == want, (curve, c, m)
                cbm, _ = pack(cv, [k % cv.fr.modulus for k in ks], [], fr_mont=True)
                got = pyref.prj_bytes_to_affine(bases.msm(cbm, m, out=M.OUT_PRJ, coef_kind="fr"), cv)
                assert got == want, (curve, c, m, "fr")
        bases.free()


# ------------------------------------------------------------------ skewed scalar distributions of the reference's own tests
def _scalars_high_hamming_weight(rnd, n, bits):
    """reference helpers/prng_unsafe.nim:198-216 random_highHammingWeight: every 64-bit limb starts all-ones and loses up to
    64/3 randomly chosen bits; extra bits over the MSB are cleared."""
    out = []
    for _ in range(n):
        v = 0
        for limb in range(4):
            w = (1 << 64) - 1
            for _ in range(rnd.randrange(64 // 3)):
                w &= ~(1 << rnd.randrange(64))
            v |= w << (64 * limb)
        out.append(v & ((1 << bits) - 1))
    return out


def _scalars_long01(rnd, n, bits):
    """reference helpers/prng_unsafe.nim:233-260 random_long01Seq: runs of equal bits of length 1 + (u6 * u5 + 16) / 31, read
    in either byte order."""
    out = []
    nbytes = (bits + 7) // 8
    for _ in range(n):
        buf = bytearray(nbytes)
        bit = 0
        while bit < nbytes * 8:
            now = 1 + (rnd.randrange(64) * rnd.randrange(32) + 16) // 31
            val = rnd.randrange(2)
            while now > 0 and bit < nbytes * 8:
                buf[bit >> 3] |= val << (bit & 7)
                now -= 1
                bit += 1
        v = int.from_bytes(bytes(buf), rnd.choice(("big", "little")))
        out.append(v & ((1 << bits) - 1))
    return out


@pytest.mark.parametrize("dist", ["highHammingWeight", "long01Seq"])
def test_skewed_scalar_distributions_2_18(M, lib, tp, dist):
    """The reference draws its MSM test scalars from uniform / high-Hamming-weight / long-0-1-run generators
    (tests/math_elliptic_curves/t_ec_template.nim:1411-1480 with helpers/prng_unsafe.nim:198-260). Same distributions at
    N = 2^18, BLS12-381 G1 and Pallas, against the closed form; scalars are NOT reduced mod r."""
    rnd = random.Random(0xC77 + len(dist))
    for curve in ("bls12_381_g1", "pallas_ec"):
        cv = CURVES[curve]
        n = 1 << 18
        rng = np.random.default_rng(18)
        k = rng.integers(1, 2**63, size=n, dtype=np.uint64)
        pts = _gen_points(lib, cv, k)
        gen = _scalars_high_hamming_weight if dist == "highHammingWeight" else _scalars_long01
        # 4096 distinct skewed scalars tiled over the N points (generation in Python is the slow part)
        base = gen(rnd, 4096, cv.scalar_bits)
        s_int = [base[i & 4095] for i in range(n)]
        scal = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in base) * (n // 4096), dtype=np.uint8).reshape(n, 32)
        total = sum(s * int(kk) for s, kk in zip(s_int, k)) % cv.fr.modulus
        got = M.multi_scalar_mul_vartime_parallel(tp, cv, scal, pts, n)
        assert pyref.jac_bytes_to_affine(got, cv) == pyref.ec_mul_fast(total, cv.gen, cv), (curve, dist)


# ------------------------------------------------------------------ several GPUs / threads inside one process
def test_in_process_device_list_closed_form(M, lib, tp):
    """ctt_b200_set_devices: the unchanged C symbol spreads one host-pointer MSM over a device list (point shards, one host
    worker thread per entry, host addition of the partial points). Every visible GPU is listed; on a one-GPU box the single
    device is listed twice and then four times, which drives the same code (two engine slots of that device side by side)."""
    cv = CURVES["bls12_381_g1"]
    n = (1 << 17) + 5
    rng = np.random.default_rng(99)
    k = rng.integers(1, 2**63, size=n, dtype=np.uint64)
    pts = _gen_points(lib, cv, k)
    scal = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    scal[:, 31] &= 0x7F
    s_int = [int.from_bytes(scal[i].tobytes(), "little") for i in range(n)]
    want = pyref.ec_mul_fast(sum(s * int(kk) for s, kk in zip(s_int, k)) % cv.fr.modulus, cv.gen, cv)
    count = M.device_count()
    lists = [list(range(count))] if count > 1 else []
    lists += [[0, 0], [0, 0, 0, 0]]
    try:
        for devs in lists:
            M.set_devices(devs)
            got = M.multi_scalar_mul_vartime_parallel(tp, cv, scal, pts, n)
            assert pyref.jac_bytes_to_affine(got, cv) == want, devs
            assert M.last_stats()["entries"] > 0
            # short MSMs stay on one device; N smaller than the list still works
            g1 = M.multi_scalar_mul_vartime_parallel(tp, cv, scal[:3], pts[:3], 3)
            w1 = pyref.ec_mul_fast(sum(s * int(kk) for s, kk in zip(s_int[:3], k[:3])) % cv.fr.modulus, cv.gen, cv)
            assert pyref.jac_bytes_to_affine(g1, cv) == w1
    finally:
        M.set_devices([])
    got = M.multi_scalar_mul_vartime_parallel(tp, cv, scal, pts, n)
    assert pyref.jac_bytes_to_affine(got, cv) == want
    # the same 16 MiB of pageable numpy memory (staged through the pinned double buffer) with the points in 3 and 5 pieces and two
    # forced affine levels: a piece is staged only when the engine asks for it
    try:
        lib.ctt_b200_set_affine_levels(2)
        for pieces in (3, 5):
            lib.ctt_b200_set_point_chunks(pieces)
            got = M.multi_scalar_mul_vartime_parallel(tp, cv, scal, pts, n)
            assert pyref.jac_bytes_to_affine(got, cv) == want, pieces
    finally:
        lib.ctt_b200_set_point_chunks(0)
        lib.ctt_b200_set_affine_levels(-1)


def test_caller_thread_with_another_current_device(M, lib, tp, oracle_lib):
    """The CUDA current device is per host thread and new threads start on device 0: an MSM issued from a worker thread must
    still run on the device the engine is bound to (every entry point switches to it and restores the caller's device).
    With two GPUs the worker thread makes the OTHER device current first; with one GPU the thread simply starts fresh."""
    import threading
    import torch
    cv = CURVES["bn254_snarks_g1"]
    _, pool = point_pool(cv)
    r = random.Random(5)
    n = 1500
    ptsl = [pool[r.randrange(len(pool))] for _ in range(n)]
    ks = [r.getrandbits(254) for _ in range(n)]
    cb, pb = pack(cv, ks, ptsl)
    want = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n), cv)
    assert pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, n), cv) == want   # binds the engine
    bound = torch.cuda.current_device()
    other = (bound + 1) % max(1, torch.cuda.device_count())
    out = {}

    def worker():
        torch.cuda.set_device(other)
        out["got"] = pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, n), cv)
        out["dev_after"] = torch.cuda.current_device()

    t = threading.Thread(target=worker)
    t.start()
    t.join()
    assert out["got"] == want
    assert out["dev_after"] == other          # the caller's current device is restored


# ------------------------------------------------------------------ engine modes that must not change the result
@pytest.mark.parametrize("curve,n", [("bls12_381_g1", 70001), ("bn254_snarks_g1", 4099), ("bls12_381_g2", 2500), ("pallas_ec", 33)])
def test_input_chunks_and_reduce_modes_same_result(M, lib, tp, oracle_lib, curve, n):
    """ctt_b200_set_input_chunks (the host input crosses PCIe in k chunks that accumulate into the same buckets) x
    ctt_b200_set_reduce_mode (bit-plane / running-sum bucket reduction) x forced batched-affine levels: every combination
    returns the oracle's group element. Few distinct points, so equal points meet inside and across chunks."""
    cv = CURVES[curve]
    rnd = random.Random(n)
    _, pool = point_pool(cv)
    pts = [pool[rnd.randrange(10)] for _ in range(n)]
    ks = [rnd.getrandbits(cv.scalar_bits) for _ in range(n)]
    if n > 40:
        pts[11] = None
        ks[12] = 0
        ks[13] = ks[14] = 1                      # P and P again: a doubling inside a bucket
        pts[13] = pts[14]
    cb, pb = pack(cv, ks, pts)
    want = pyref.jac_bytes_to_affine(oracle_lib.msm(cv, cb, pb, n), cv)
    try:
        for chunks in (1, 2, 3, 5, 8):
            for mode in (0, 1):
                for levels in (-1, 0, 2):
                    lib.ctt_b200_set_input_chunks(chunks)
                    lib.ctt_b200_set_reduce_mode(mode)
                    lib.ctt_b200_set_affine_levels(levels)
                    got = M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, n)
                    assert pyref.jac_bytes_to_affine(got, cv) == want, (curve, chunks, mode, levels)
        # point pieces: level 0 of the affine sums partitioned by the last piece a pair touches (ctt_b200_set_point_chunks)
        lib.ctt_b200_set_input_chunks(1)
        lib.ctt_b200_set_reduce_mode(0)
        for pieces in (2, 3, 4, 7, 8):
            for levels in (1, 3, 0):
                lib.ctt_b200_set_point_chunks(pieces)
                lib.ctt_b200_set_affine_levels(levels)
                got = M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, n)
                assert pyref.jac_bytes_to_affine(got, cv) == want, (curve, "pieces", pieces, levels)
    finally:
        lib.ctt_b200_set_point_chunks(0)
        lib.ctt_b200_set_input_chunks(0)
        lib.ctt_b200_set_reduce_mode(0)
        lib.ctt_b200_set_affine_levels(-1)


def test_forced_window_sizes_both_reduce_modes(M, lib):
    """every window size 2..20 through the device-pointer entry with both bucket reductions (the bit-plane form splits the
    2^(c-1) buckets into a 2^ceil((c-1)/2) x 2^floor((c-1)/2) matrix: odd / even c, c = 2 with a single column) -- closed form"""
    import torch
    cv = CURVES["bn254_snarks_g1"]
    n = 3000
    rng = np.random.default_rng(5)
    k = rng.integers(1, 2**63, size=n, dtype=np.uint64)
    pts = _gen_points(lib, cv, k)
    scal = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    scal[:, 31] &= 0x3F
    s_int = [int.from_bytes(scal[i].tobytes(), "little") for i in range(n)]
    want = pyref.ec_mul_fast(sum(s * int(kk) for s, kk in zip(s_int, k)) % cv.fr.modulus, cv.gen, cv)
    d_s, d_p = torch.from_numpy(scal).cuda(), torch.from_numpy(pts).cuda()
    try:
        for mode in (0, 1):
            lib.ctt_b200_set_reduce_mode(mode)
            for c in range(2, 21):
                got = M.msm_device_ptrs(cv, d_s.data_ptr(), d_p.data_ptr(), n, force_c=c)
                assert pyref.jac_bytes_to_affine(got, cv) == want, (mode, c)
            # a window range (multi-GPU window sharding): partial sums of two halves add up
            c, W = M.plan(cv, n)
            parts = [M.msm_device_ptrs(cv, d_s.data_ptr(), d_p.data_ptr(), n, out=M.OUT_XYZZ, force_c=c, win_begin=a, win_end=b)
                     for a, b in ((0, W // 2), (W // 2, W))]
            both = pyref.ec_add(xyzz_bytes_to_affine(parts[0], cv), xyzz_bytes_to_affine(parts[1], cv), cv)
            assert both == want, mode
    finally:
        lib.ctt_b200_set_reduce_mode(0)


def test_window_digits_path_closed_form(M, lib):
    """ctt_b200_msm_device_digits + ctt_b200_combine_window_digits (the multi-GPU window-sharded leg: digits of the window sums
    stay on the device, one host pass at the end), here with the window ranges of 1, 2, 3 and 5 "ranks" run one after the
    other on one GPU -- closed form; BLS12-381 G1 and Pallas."""
    import torch
    from constantine_b200 import sharded
    for curve, n in (("bls12_381_g1", 50000), ("pallas_ec", 7001)):
        cv = CURVES[curve]
        rng = np.random.default_rng(n)
        k = rng.integers(1, 2**63, size=n, dtype=np.uint64)
        pts = _gen_points(lib, cv, k)
        scal = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        scal[:, 31] &= (1 << (cv.scalar_bits - 248)) - 1
        s_int = [int.from_bytes(scal[i].tobytes(), "little") for i in range(n)]
        want = pyref.ec_mul_fast(sum(s * int(kk) for s, kk in zip(s_int, k)) % cv.fr.modulus, cv.gen, cv)
        d_s, d_p = torch.from_numpy(scal).cuda(), torch.from_numpy(pts).cuda()
        c, W = M.plan(cv, n)
        groups = M.digits_per_window(c)
        xyzz = 4 * cv.coord_bytes
        for world in (1, 2, 3, 5):
            parts = []
            for rank in range(world):
                wb, we = sharded.window_range(W, world, rank)
                buf = torch.zeros((we - wb) * groups * xyzz, dtype=torch.uint8, device="cuda")
                torch.cuda.synchronize()          # the engine writes the buffer from its own stream
                assert M.msm_device_digits(cv, buf.data_ptr(), d_s.data_ptr(), d_p.data_ptr(), n, force_c=c, win_begin=wb, win_end=we) == groups
                torch.cuda.synchronize()
                parts.append(buf.cpu().numpy().tobytes())
            got = M.combine_window_digits(cv, b"".join(parts), c, W)
            assert pyref.jac_bytes_to_affine(got, cv) == want, (curve, world)
