"""Many small MSMs: the PeerDAS bank shape of the reference -- 128 fixed-base MSMs of 64
BLS12-381 G1 points each (reference commitments_setups/ethereum_kzg_srs.nim:133, benchmarks/bench_eth_eip7594_peerdas.nim) --
timed four ways on one GPU, host scalars in every timed call (pinned memory, result read back):
  loop      128 calls of the single-MSM C symbol (what a caller gets by linking this library and changing nothing)
  batch     one ctt_b200_msm_batch_host call (bases travel with the call)
  cached    one ctt_b200_msm_batch_cached_bases call over device-resident bases
  table     the same over the precomputed window table (fixed-base mode)
plus the sum of 2^20 points (ctt_b200_sum_reduce_host) and, as the CPU figure, the oracle port looping over the bank.
Prints one JSON line per measurement.  Usage: python tools/bench_batch.py [--count 128 --n 64 --reps 20]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=128)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--curve", default="bls12_381_g1")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import torch
    from constantine_b200 import _lib, msm as M
    from constantine_b200.curves import CURVES
    from oracle import pyref
    lib = _lib.load()
    cv = CURVES[a.curve]
    count, n = a.count, a.n
    tot = count * n
    rng = np.random.default_rng(17)
    k = rng.integers(1, 2**63, size=tot, dtype=np.uint64)
    gen = b"".join(cv.fp.to_mont(c).to_bytes(cv.fp.nbytes, "little") for coord in cv.gen for c in coord)
    pts_np = np.empty((tot, cv.aff_bytes), dtype=np.uint8)
    assert lib.ctt_b200_scalar_mul_u64(cv.curve_id, gen, k.ctypes.data, tot, pts_np.ctypes.data) == 0
    s_np = rng.integers(0, 256, size=(tot, 32), dtype=np.uint8)
    s_np[:, 31] &= 0x3F
    pts = torch.from_numpy(pts_np).pin_memory()
    sc = torch.from_numpy(s_np).pin_memory()
    size = cv.jac_bytes
    out = ctypes.create_string_buffer(size * count)

    def expect(m):
        kk = [int(x) for x in k[m * n:(m + 1) * n]]
        ss = [int.from_bytes(s_np[m * n + i].tobytes(), "little") for i in range(n)]
        return pyref.ec_mul_fast(sum(x * y for x, y in zip(kk, ss)) % cv.fr.modulus, cv.gen, cv)

    def check(tag):
        for m in (0, count // 2, count - 1):
            got = pyref.jac_bytes_to_affine(out.raw[m * size:(m + 1) * size], cv)
            assert got == expect(m), (tag, m)

    def timed(tag, fn, extra=None):
        for _ in range(3):
            fn()
        check(tag)
        t0 = time.perf_counter()
        for _ in range(a.reps):
            fn()
        dt = (time.perf_counter() - t0) / a.reps
        st = M.last_stats()
        rec = {"what": tag, "curve": a.curve, "msms": count, "terms_each": n, "ms_per_bank": round(dt * 1e3, 4),
               "msm_per_s": round(count / dt, 1), "c": st["c"], "windows": st["num_windows"], "kernel_launches": st["kernel_launches"],
               "engine_ms": round(st["ms_total"], 4)}
        if extra:
            rec.update(extra)
        print(json.dumps(rec), flush=True)
        return dt

    named = _lib.named_msm(f"ctt_{cv.cprefix}_jac_multi_scalar_mul_big_coefs_vartime_parallel")
    sp, pp, op = sc.data_ptr(), pts.data_ptr(), ctypes.addressof(out)

    def loop():
        for m in range(count):
            named(None, op + m * size, sp + m * n * 32, pp + m * n * cv.aff_bytes, n)
    timed("loop_of_single_msm_calls", loop)

    def batch():
        assert lib.ctt_b200_msm_batch_host(cv.curve_id, 0, op, sp, pp, count, n, 0, 0) == 0
    timed("batch_host", batch)

    bases = lib.ctt_b200_bases_upload(cv.curve_id, pp, tot)

    def cached():
        assert lib.ctt_b200_msm_batch_cached_bases(bases, 0, op, sp, count, n, 0, 0) == 0
    timed("batch_cached_bases", cached)

    t0 = time.perf_counter()
    c = lib.ctt_b200_bases_precompute_for(bases, n, 0)
    tpre = time.perf_counter() - t0
    timed("batch_fixed_base_table", cached, {"table_c": c, "precompute_s": round(tpre, 4),
                                             "table_bytes": (cv.scalar_bits // c + 1) * tot * cv.aff_bytes})
    lib.ctt_b200_bases_free(bases)

    # sum of points
    nsum = 1 << 20
    k2 = rng.integers(1, 2**63, size=nsum, dtype=np.uint64)
    p2 = np.empty((nsum, cv.aff_bytes), dtype=np.uint8)
    assert lib.ctt_b200_scalar_mul_u64(cv.curve_id, gen, k2.ctypes.data, nsum, p2.ctypes.data) == 0
    p2t = torch.from_numpy(p2).pin_memory()
    r = ctypes.create_string_buffer(size)
    for _ in range(3):
        lib.ctt_b200_sum_reduce_host(cv.curve_id, 0, r, p2t.data_ptr(), nsum)
    want = pyref.ec_mul_fast(sum(int(x) for x in k2) % cv.fr.modulus, cv.gen, cv)
    assert pyref.jac_bytes_to_affine(r.raw, cv) == want
    t0 = time.perf_counter()
    for _ in range(a.reps):
        lib.ctt_b200_sum_reduce_host(cv.curve_id, 0, r, p2t.data_ptr(), nsum)
    dt = (time.perf_counter() - t0) / a.reps
    print(json.dumps({"what": "sum_reduce_host", "curve": a.curve, "points": nsum, "ms": round(dt * 1e3, 4),
                      "Mpoints_per_s": round(nsum / dt / 1e6, 1), "h2d_GBps": round(nsum * cv.aff_bytes / dt / 1e9, 2)}), flush=True)

    if not a.no_cpu:
        from oracle import oracle
        oracle.build()
        oracle.load()
        sb, pb = s_np.tobytes(), pts_np.tobytes()
        t0 = time.perf_counter()
        for m in range(count):
            oracle.msm(cv, sb[m * n * 32:(m + 1) * n * 32], pb[m * n * cv.aff_bytes:(m + 1) * n * cv.aff_bytes], n)
        dt = time.perf_counter() - t0
        print(json.dumps({"what": "cpu_oracle_port_loop", "curve": a.curve, "msms": count, "terms_each": n,
                          "ms_per_bank": round(dt * 1e3, 2), "msm_per_s": round(count / dt, 1),
                          "note": "oracle restatement, one MSM after the other (serial per MSM)"}), flush=True)


if __name__ == "__main__":
    main()
