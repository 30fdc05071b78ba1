"""Experimental batched-affine levels (ctt_b200_set_affine_levels): per-level timing of one device-resident
MSM with a closed-form result check.   python tools/bench_affine.py [--curve bls12_381_g1 --logn 20 --levels 0,1,2,3,4]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--curve", default="bls12_381_g1")
    ap.add_argument("--logn", type=int, default=20)
    ap.add_argument("--levels", default="0,1,2,3,4")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cs", default="0", help="forced window sizes to sweep (0 = the engine's choice)")
    ap.add_argument("--slice", type=int, default=0, help="forced slice length K of k_accumulate (0 = automatic, -k = automatic with upper limit k)")
    ap.add_argument("--win", default="", help="begin:end -- only this window range (the shard of one rank of a window-sharded multi-GPU run; no result check)")
    a = ap.parse_args()
    import torch
    from constantine_b200 import _lib, msm as M
    from constantine_b200.curves import CURVES
    from oracle import pyref
    lib = _lib.load()
    cv = CURVES[a.curve]
    n = 1 << a.logn
    rng = np.random.default_rng(23)
    k = rng.integers(1, 2**63, size=n, dtype=np.uint64)
    gen = b"".join(cv.fp.to_mont(c).to_bytes(cv.fp.nbytes, "little") for coord in cv.gen for c in coord)
    pts = np.empty((n, cv.aff_bytes), dtype=np.uint8)
    assert lib.ctt_b200_scalar_mul_u64(cv.curve_id, gen, k.ctypes.data, n, pts.ctypes.data) == 0
    s = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    s[:, 31] &= 0x3F
    e = sum(int(x) * int.from_bytes(s[i].tobytes(), "little") for i, x in enumerate(k)) % cv.fr.modulus
    want = pyref.ec_mul_fast(e, cv.gen, cv)
    d_pts = torch.from_numpy(pts).cuda()
    d_s = torch.from_numpy(s).cuda()
    lib.ctt_b200_set_tuning(0, 0, a.slice if a.slice != 0 else -1)
    for lv, fc in [(int(x), int(y)) for y in a.cs.split(",") for x in a.levels.split(",")]:
        lib.ctt_b200_set_affine_levels(lv)
        ok = True
        best = None
        for _ in range(a.reps):
            if a.win:
                wb, we = (int(x) for x in a.win.split(":"))
                M.msm_device_ptrs(cv, d_s.data_ptr(), d_pts.data_ptr(), n, out=M.OUT_XYZZ, force_c=fc or M.plan(cv, n)[0], win_begin=wb, win_end=we)
                ok = None
            else:
                got = M.msm_device_ptrs(cv, d_s.data_ptr(), d_pts.data_ptr(), n, force_c=fc)
                ok = ok and pyref.jac_bytes_to_affine(got, cv) == want
            st = M.last_stats()
            if best is None or st["ms_total"] < best["ms_total"]:
                best = st
        print(json.dumps({"curve": a.curve, "logn": a.logn, "affine_levels": lv, "ok": ok,
                          **{kk: round(v, 4) if isinstance(v, float) else v for kk, v in best.items()}}), flush=True)
    lib.ctt_b200_set_affine_levels(-1)


if __name__ == "__main__":
    main()
