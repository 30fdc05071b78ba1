#!/usr/bin/env python3
"""Timing table for the BASELINE.json configs that fit one GPU (device-resident inputs, per-phase CUDA-event times).
   python tools/bench_configs.py [quick]
Each line: curve, N, c, ms per MSM (mean of 5 after 2 warm-ups), MSM/s, Mop point-adds/s, closed-form check."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from constantine_b200 import _lib, msm as M
from constantine_b200.curves import CURVES
from oracle import pyref

lib = _lib.load()
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
CONFIGS = [("bn254_snarks_g1", 8, "big"), ("bls12_381_g1", 16, "big"), ("bls12_381_g1", 18, "big"), ("bls12_381_g1", 20, "big"),
           ("bls12_381_g1", 22, "big"), ("pallas_ec", 20, "fr"), ("pallas_ec", 22, "fr"), ("bn254_snarks_g1", 20, "fr"),
           ("bls12_381_g2", 16, "big"), ("bls12_381_g2", 18, "big")]
if quick:
    CONFIGS = [c for c in CONFIGS if c[1] <= 18]
for name, logn, kind in CONFIGS:
    cv = CURVES[name]
    n = 1 << logn
    rng = np.random.default_rng(logn + cv.curve_id)
    k = rng.integers(1, 2**63, size=n, dtype=np.uint64)
    gen = b"".join(cv.fp.to_mont(c).to_bytes(cv.fp.nbytes, "little") for coord in cv.gen for c in coord)
    pts = np.empty((n, cv.aff_bytes), dtype=np.uint8)
    assert lib.ctt_b200_scalar_mul_u64(cv.curve_id, gen, k.ctypes.data, n, pts.ctypes.data) == 0
    r = cv.fr.modulus
    s_int = None
    if kind == "fr":   # Montgomery-form scalars < r (the fr_coefs entry points; config 4 of BASELINE.json)
        lo = rng.integers(0, 2**62, size=n, dtype=np.uint64).astype(object)
        hi = rng.integers(0, 2**62, size=n, dtype=np.uint64).astype(object)
        s_int = [(int(a) * (1 << 190) + int(b) * 0x9E3779B97F4A7C15F39CC0605CEDC835) % r for a, b in zip(lo, hi)]
        scal = np.frombuffer(b"".join(((s * cv.fr.R) % r).to_bytes(32, "little") for s in s_int), dtype=np.uint8).reshape(n, 32).copy()
    else:
        scal = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        scal[:, 31] &= (1 << (cv.scalar_bits - 248)) - 1
    d_s = torch.from_numpy(scal).cuda(); d_p = torch.from_numpy(pts).cuda(); torch.cuda.synchronize()
    for _ in range(2):
        res = M.msm_device_ptrs(cv, d_s.data_ptr(), d_p.data_ptr(), n, fr_mont=(kind == "fr"))
    stats = []
    for _ in range(5):
        res = M.msm_device_ptrs(cv, d_s.data_ptr(), d_p.data_ptr(), n, fr_mont=(kind == "fr"))
        stats.append(M.last_stats())
    ms = sum(s["ms_total"] for s in stats) / len(stats)
    st = stats[-1]
    ok = None
    if n <= (1 << 20):   # closed form: sum s_i k_i G
        if s_int is None:
            s_int = [int.from_bytes(scal[i].tobytes(), "little") for i in range(n)]
        tot = sum(a * int(b) for a, b in zip(s_int, k)) % r
        ok = pyref.jac_bytes_to_affine(res, cv) == pyref.ec_mul_fast(tot, cv.gen, cv)
    W = cv.scalar_bits // st["c"] + 1
    padds = W * (n + 2 * (1 << (st["c"] - 1))) + W * (st["c"] + 1)
    print(json.dumps({"curve": name, "logn": logn, "coefs": kind, "c": st["c"], "windows": st["num_windows"], "slice_len": st["slice_len"],
                      "ms_per_msm": round(ms, 3), "msm_per_s": round(1e3 / ms, 2), "Mop_point_adds_per_s": round(padds / ms / 1e3, 1),
                      "closed_form_ok": ok, **{k: round(sum(s[k] for s in stats) / len(stats), 3) for k in
                                               ("ms_digits", "ms_sort", "ms_accumulate", "ms_fixup", "ms_reduce", "ms_d2h_tail")}}), flush=True)
    del d_s, d_p
    torch.cuda.empty_cache()
