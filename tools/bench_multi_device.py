#!/usr/bin/env python3
"""One process, several GPUs behind the unchanged C symbol (ctt_b200_set_devices): end-to-end time of
ctt_<curve>_jac_multi_scalar_mul_big_coefs_vartime_parallel with pinned and with pageable host buffers for 1, 2, 4, ... devices,
closed-form check of every result.   python tools/bench_multi_device.py [--curve bls12_381_g1 --logn 20 --reps 8]"""
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
    ap.add_argument("--curve", default="bls12_381_g1")
    ap.add_argument("--logn", type=int, default=20)
    ap.add_argument("--reps", type=int, default=8)
    a = ap.parse_args()
    import torch
    from constantine_b200 import _lib, msm as M
    from constantine_b200.curves import CURVES
    from oracle import pyref
    lib = _lib.load()
    cv = CURVES[a.curve]
    n = 1 << a.logn
    rng = np.random.default_rng(77)
    k = rng.integers(1, 2**63, size=n, dtype=np.uint64)
    gen = b"".join(cv.fp.to_mont(c).to_bytes(cv.fp.nbytes, "little") for coord in cv.gen for c in coord)
    pts = np.empty((n, cv.aff_bytes), dtype=np.uint8)
    assert lib.ctt_b200_scalar_mul_u64(cv.curve_id, gen, k.ctypes.data, n, pts.ctypes.data) == 0
    scal = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    scal[:, 31] &= (1 << (cv.scalar_bits - 248)) - 1
    s_int = [int.from_bytes(scal[i].tobytes(), "little") for i in range(n)]
    want = pyref.ec_mul_fast(sum(x * int(y) for x, y in zip(s_int, k)) % cv.fr.modulus, cv.gen, cv)
    h_s, h_p = torch.from_numpy(scal).pin_memory(), torch.from_numpy(pts).pin_memory()
    fn = _lib.named_msm(f"ctt_{cv.cprefix}_jac_multi_scalar_mul_big_coefs_vartime_parallel")
    tp = M.Threadpool.new(1)
    r = ctypes.create_string_buffer(cv.jac_bytes)
    count = M.device_count()
    ks = [x for x in (1, 2, 4, 8) if x <= count]
    for nd in ks:
        M.set_devices(list(range(nd)) if nd > 1 else [])
        for label, sp, pp in (("pinned", h_s.data_ptr(), h_p.data_ptr()), ("pageable", scal.ctypes.data, pts.ctypes.data)):
            for _ in range(3):
                fn(tp._h, r, sp, pp, n)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(a.reps):
                fn(tp._h, r, sp, pp, n)
            dt = (time.perf_counter() - t0) / a.reps
            ok = pyref.jac_bytes_to_affine(r.raw, cv) == want
            print(json.dumps({"curve": a.curve, "logn": a.logn, "devices_in_process": nd, "host_buffers": label, "ms_per_msm_wall": round(dt * 1e3, 3),
                              "msm_per_s": round(1 / dt, 2), "closed_form_ok": ok}), flush=True)
    M.set_devices([])


if __name__ == "__main__":
    main()
