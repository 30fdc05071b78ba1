#!/usr/bin/env python3
"""Tuning sweep on one GPU: window size c and reduce chunk L for BLS12-381 G1 at N = 2^logn (device-resident inputs)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from constantine_b200 import _lib, msm as M
from constantine_b200.curves import CURVES

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [14, 15, 16, 17]
Ls = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8, 16, 32]
cv = CURVES["bls12_381_g1"]
lib = _lib.load()
n = 1 << logn
scal, pts, _ = bench.make_inputs(n, 1234)
d_s = torch.from_numpy(scal).cuda(); d_p = torch.from_numpy(pts).cuda(); torch.cuda.synchronize()
ref = None
for c in cs:
    for L in Ls:
        lib.ctt_b200_set_tuning(c, L, 0)
        for _ in range(2):
            r = M.msm_device_ptrs(cv, d_s.data_ptr(), d_p.data_ptr(), n)
        acc = []
        for _ in range(5):
            r = M.msm_device_ptrs(cv, d_s.data_ptr(), d_p.data_ptr(), n)
            acc.append(M.last_stats())
        avg = {k: sum(a[k] for a in acc) / len(acc) for k in acc[0] if k.startswith("ms_")}
        from oracle import pyref
        aff = pyref.jac_bytes_to_affine(r, cv)
        if ref is None: ref = aff
        print(json.dumps({"logn": logn, "c": c, "L": L, "same_point": aff == ref, **{k: round(v, 3) for k, v in avg.items()}}), flush=True)
