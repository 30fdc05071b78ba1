"""CPU: host-side logic of the product -- constants, byte layouts, partial-sum combination (host tail arithmetic),
sharding helpers."""
import ctypes

import pytest

from helpers import CURVES, affine_to_xyzz_bytes, point_pool, pyref, xyzz_bytes_to_affine


def test_montgomery_constants_match_survey_appendix_a():
    from constantine_b200.curves import FIELDS
    exp = {
        "bn254_snarks_fp": (2, 0x87d20782e4866389, 0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d),
        "bn254_snarks_fr": (2, 0xc2e1f593efffffff, 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb),
        "bls12_381_fp": (3, 0x89f3fffcfffcfffd, 0x15f65ec3fa80e4935c071a97a256ec6d77ce5853705257455f48985753c758baebf4000bc40c0002760900000002fffd),
        "bls12_381_fr": (1, 0xfffffffeffffffff, 0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe),
        "pallas_fp": (1, 0x992d30ecffffffff, 0x3fffffffffffffffffffffffffffffff992c350be41914ad34786d38fffffffd),
        "pallas_fr": (1, 0x8c46eb20ffffffff, 0x3fffffffffffffffffffffffffffffff992c350be34205675b2b3e9cfffffffd),
    }
    for name, (spare, inv, one) in exp.items():
        f = FIELDS[name]
        assert (f.spare_bits, f.m0ninv64, f.one_mont) == (spare, inv, one)
    assert FIELDS["vesta_fp"].modulus == FIELDS["pallas_fr"].modulus and FIELDS["vesta_fr"].modulus == FIELDS["pallas_fp"].modulus


def test_struct_sizes_match_reference_layout():
    sizes = {"bls12_381_g1": (96, 144), "bn254_snarks_g1": (64, 96), "pallas_ec": (64, 96), "vesta_ec": (64, 96),
             "bls12_381_g2": (192, 288), "bn254_snarks_g2": (128, 192)}
    for name, (aff, jac) in sizes.items():
        assert (CURVES[name].aff_bytes, CURVES[name].jac_bytes) == (aff, jac)


@pytest.mark.parametrize("curve", list(CURVES))
def test_sum_partials_is_the_group_sum(curve, rng):
    """ctt_b200_sum_partials runs the product's host arithmetic (host_field.hpp: Montgomery CIOS on 64-bit limbs, XYZZ
    add/double, XYZZ -> Jacobian / projective) -- checked against the exact tier, incl. P+P, P-P and infinity."""
    from constantine_b200 import msm as M
    cv = CURVES[curve]
    _, pool = point_pool(cv)
    P, Q = pool[3], pool[4]
    sets = [[P, Q], [P, P], [P, pyref.ec_neg(P, cv)], [None, Q, None], [None], [pool[i] for i in range(8)], [P, P, P, Q, Q]]
    for pts in sets:
        want = None
        for A in pts:
            want = pyref.ec_add(want, A, cv)
        raw = b"".join(affine_to_xyzz_bytes(A, cv) for A in pts)
        assert pyref.jac_bytes_to_affine(M.sum_partials(cv, raw, len(pts), out=M.OUT_JAC), cv) == want
        assert pyref.prj_bytes_to_affine(M.sum_partials(cv, raw, len(pts), out=M.OUT_PRJ), cv) == want
        assert xyzz_bytes_to_affine(M.sum_partials(cv, raw, len(pts), out=M.OUT_XYZZ), cv) == want
    # reference encodings of the neutral element: Jacobian (1,1,0), projective (0,1,0)
    inf_j = M.sum_partials(cv, affine_to_xyzz_bytes(None, cv), 1, out=M.OUT_JAC)
    inf_p = M.sum_partials(cv, affine_to_xyzz_bytes(None, cv), 1, out=M.OUT_PRJ)
    cb = cv.coord_bytes
    one = pyref.coord_to_bytes((1,) + (0,) * (cv.ext_degree - 1), cv.fp)
    assert inf_j == one + one + bytes(cb)
    assert inf_p == bytes(cb) + one + bytes(cb)


def test_balanced_chunks():
    from constantine_b200.sharded import balanced_chunk, window_range
    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [balanced_chunk(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [window_range(16, 8, r) for r in range(8)] == [(2 * r, 2 * r + 2) for r in range(8)]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from constantine_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_specialised_multipliers_carry_chains(rng):
    """fe_dot2 / fe_sqr of csrc/field.cuh, emulated instruction by instruction (tests/carry_chain_emulation.py): exact
    results on edge and random operands and no dropped carry, for every base field the mixed add runs over."""
    import carry_chain_emulation as emu
    from constantine_b200.curves import FIELDS
    for name in ("bls12_381_fp", "bn254_snarks_fp", "pallas_fp", "vesta_fp"):
        f = FIELDS[name]
        p, n = f.modulus, f.nbytes // 4
        R = 1 << (32 * n)
        assert 3 * p < R, name                      # the headroom fe_dot2's static_assert demands
        rinv = pow(R, -1, p)
        edge = [0, 1, p - 1, p - 2, (1 << (f.bits - 1)) % p, R % p, (R - 1) % p]
        quads = [(a, b, c, d) for a in edge for b in edge[:4] for c in (p - 1, 0) for d in (p - 1, 1)]
        quads += [tuple(rng.randrange(p) for _ in range(4)) for _ in range(300)]
        for a, b, c, d in quads:
            got, _ = emu.fe_dot2(p, n, a, b, c, d)
            assert got == (a * b + c * d) * rinv % p, (name, a, b, c, d)
        for a in edge + [rng.randrange(p) for _ in range(300)]:
            got, macs = emu.fe_sqr(p, n, a)
            assert got == a * a * rinv % p, (name, a)
            assert macs == n * (n - 1) // 2 + n + n * n        # 222 for 12 limbs, against 2 n^2 = 288 for fe_mul


def test_batched_affine_schedule_prototype():
    """tools/proto_affine_levels.py: the level schedule of csrc/msm_affine.cuh (run bounds, level offsets, plan, per-thread
    batches with one shared inversion each, survivors) modelled statement for statement with exact arithmetic: exact bucket
    sums with infinities, P + P, P - P, single entries, zero digits and one giant run."""
    import importlib.util
    import os
    from helpers import ROOT
    spec = importlib.util.spec_from_file_location("proto_affine_levels", os.path.join(ROOT, "tools", "proto_affine_levels.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for seed in (5, 6):
        st = mod.self_test(seed)
        assert st["adds"] > 400 and st["inversions"] > 0


def test_safegcd_model():
    """csrc/field_inv.cuh (Bernstein-Yang divsteps in signed 30-bit limbs, batches of 30) modelled limb for limb with
    register-width assertions: R^2 / x mod p for every field, within the batch bound the device loop uses."""
    import safegcd_emulation as emu
    from constantine_b200.curves import FIELDS
    worst = emu.self_test({k: (f.modulus, f.bits) for k, f in FIELDS.items()}, samples=60)
    for name, (used, bound) in worst.items():
        assert used <= bound, name


def test_bit_plane_bucket_reduction_identity():
    """msm_kernels.cuh k_rowcol_sums / k_plane_sums / k_plane_combine + the host Horner pass of msm_engine.cuh, modelled over the
    integers (any abelian group): for every window size the radix-16 digits reproduce  sum_j (j+1) * bucket[j]."""
    import random
    rnd = random.Random(7)
    for c in range(2, 15):
        B = 1 << (c - 1)
        a = (c - 1) // 2
        rbits = (c - 1) - a
        C, R = 1 << a, 1 << rbits
        buckets = [rnd.randrange(-1000, 1000) if rnd.random() < 0.7 else 0 for _ in range(B)]
        want = sum((j + 1) * b for j, b in enumerate(buckets))
        H = [sum(buckets[h * C + l] for l in range(C)) for h in range(R)]
        L = [sum(buckets[h * C + l] for h in range(R)) for l in range(C)]
        planes = []
        for q in range(c - 1):
            if q < a:
                planes.append(sum(L[l] for l in range(C) if ((l + 1) >> q) & 1))
            else:
                v = sum(H[h] for h in range(R) if (h >> (q - a)) & 1)
                if q == a:
                    v += L[C - 1]
                planes.append(v)
        assert sum(v << q for q, v in enumerate(planes)) == want, c
        groups = (len(planes) + 3) // 4
        digits = []
        for g in range(groups):
            r = 0
            for k in (3, 2, 1, 0):
                r = 2 * r + (planes[4 * g + k] if 4 * g + k < len(planes) else 0)
            digits.append(r)
        # the host pass: one doubling per bit position from the top, one addition per digit
        emax = 4 * (groups - 1)
        by_exp = [0] * (emax + 1)
        for g, v in enumerate(digits):
            by_exp[4 * g] += v
        r = 0
        for e in range(emax, -1, -1):
            r = 2 * r + by_exp[e]
        assert r == want


def test_point_piece_boundaries_match_the_pair_classification():
    """msm_engine.cuh msm_host_on cuts the points into P pieces at ceil(n q / P); msm_affine.cuh pair_chunk classifies a point index i
    as floor(i P / n). A pair may only be scheduled with the pieces of both operands present, so the two must agree for every
    index -- in particular never classify an index EARLIER than the piece that carries it."""
    import random
    rnd = random.Random(11)
    for _ in range(300):
        n = rnd.choice([1, 2, 3, 7, 64, 1000, 4097, 65536, 70001, (1 << 20) + 5, rnd.randrange(1, 1 << 22)])
        for P in range(1, 9):
            if P > n:
                continue
            lo = [(n * q + P - 1) // P for q in range(P + 1)]
            assert lo[0] == 0 and lo[P] == n and all(a <= b for a, b in zip(lo, lo[1:]))
            for i in {0, n - 1, n // 2, *(min(n - 1, x) for x in lo[:-1]), *(max(0, x - 1) for x in lo[1:]), rnd.randrange(n)}:
                q = min(P - 1, i * P // n)
                assert lo[q] <= i < lo[q + 1], (n, P, i, q)


def test_accumulate_slice_fit_covers_every_entry_count():
    """msm_kernels.cuh accumulate_waves / accumulate_slice_len: the kernel picks the slice length K from the ACTUAL entry count n, the host
    sizes the grid from its upper bound of n. Restated here to pin the two properties the launch relies on: (1) ceil(n / K) slices never
    exceed m(n) * threads, (2) m is monotone in n -- so a grid of m(n_upper) * threads covers any n <= n_upper; and K stays within
    [16, limit] once a wave is full."""
    MIN_SLICE = 16

    def waves(n, threads, k_max):
        per_wave = threads * k_max
        return max(1, -(-n // per_wave))

    def slice_len(n, threads, k_max):
        per = waves(n, threads, k_max) * threads
        return max(MIN_SLICE, -(-n // per))

    import random
    r = random.Random(20)
    for threads in (148 * 2 * 128, 148 * 3 * 128, 148 * 128, 4 * 128):
        for k_max in (32, 64, 128):
            ns = [0, 1, 15, 16, 17, threads * MIN_SLICE - 1, threads * MIN_SLICE, threads * MIN_SLICE + 1,
                  threads * k_max - 1, threads * k_max, threads * k_max + 1, 3 * threads * k_max + 7, 1 << 24, (1 << 31) - 1]
            ns += [r.randrange(1, 1 << 27) for _ in range(300)]
            prev_m = 0
            for n in sorted(ns):
                m, K = waves(n, threads, k_max), slice_len(n, threads, k_max)
                assert m >= prev_m
                prev_m = m
                assert -(-n // K) <= m * threads                       # the grid of m waves holds every slice
                assert MIN_SLICE <= K <= max(MIN_SLICE, k_max)
                if n >= threads * MIN_SLICE:                           # a full wave: no slot of the m waves idles more than one entry's worth
                    assert m * threads * (K - 1) < n + m * threads
    # the shapes of profiles/slice_fit_r2t.txt (BLS12-381 G1: 2 blocks of 128 threads on 148 SMs)
    t = 148 * 2 * 128
    assert slice_len(2 * (1 << 20), t, 64) == 56            # 2 of 16 windows of N = 2^20: one wave
    assert slice_len(20 * (1 << 16), t, 64) == 35           # N = 2^16, c = 13
    assert slice_len(19 * (1 << 18), t, 64) == 44 and waves(19 * (1 << 18), t, 64) == 3


@pytest.mark.parametrize("curve", list(CURVES))
def test_combine_window_digits_is_the_weighted_sum(curve):
    """ctt_b200_combine_window_digits = the host tail of every single MSM (msm_engine.cuh horner_window_digits: one doubling per bit
    position, one addition per radix-16 digit point, on host_field.hpp): sum_w 2^(c w) sum_g 16^g D_{w,g} against the exact tier, for
    every coordinate field (6- and 4-limb Fp, both Fp2), with empty digits and repeated points in the list."""
    from constantine_b200 import msm as M
    cv = CURVES[curve]
    ks, pool = point_pool(cv)
    c = 9
    W = 255 // c + 1
    groups = M.digits_per_window(c)
    assert groups == 2
    parts, scalar = [], 0
    for w in range(W):
        for g in range(groups):
            i = (5 * w + 3 * g) % len(pool)
            if (w + g) % 7 == 3:
                parts.append(affine_to_xyzz_bytes(None, cv))
                continue
            parts.append(affine_to_xyzz_bytes(pool[i], cv))
            scalar += ks[i] << (c * w + 4 * g)
    want = pyref.ec_mul_fast(scalar % cv.fr.modulus, cv.gen, cv)
    raw = b"".join(parts)
    assert pyref.jac_bytes_to_affine(M.combine_window_digits(cv, raw, c, W, out=M.OUT_JAC), cv) == want
    assert pyref.prj_bytes_to_affine(M.combine_window_digits(cv, raw, c, W, out=M.OUT_PRJ), cv) == want
    assert xyzz_bytes_to_affine(M.combine_window_digits(cv, raw, c, W, out=M.OUT_XYZZ), cv) == want


def test_host_multiplication_mulx_adx_matches_portable(tmp_path):
    """host_field.hpp picks a MULX / ADCX / ADOX multiplication at run time on x86-64 hosts that have BMI2 + ADX. tools/bench_host_field.cpp
    compares it with the portable CIOS form on ~10^6 operand pairs per field (edge values: 0, 1, p-1, p-2, all-ones limbs; random;
    long chains of dependent products) -- built and run here with the host compiler."""
    import shutil
    import subprocess
    from helpers import ROOT
    import os
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "bench_host_field")
    subprocess.check_call([cxx, "-O2", "-D__host__=", "-D__device__=", "-I", os.path.join(ROOT, "constantine_b200", "csrc"),
                           os.path.join(ROOT, "tools", "bench_host_field.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count(" ok") == 4 and "MISMATCH" not in out.stdout
