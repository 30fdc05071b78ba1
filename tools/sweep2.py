#!/usr/bin/env python3
"""slice-length sweep (K) at fixed c"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from constantine_b200 import _lib, msm as M
from constantine_b200.curves import CURVES
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
curve = sys.argv[4] if len(sys.argv) > 4 else "bls12_381_g1"
bench.CURVE = curve
cv = CURVES[curve]; lib = _lib.load(); n = 1 << logn
scal, pts, _ = bench.make_inputs(n, 1234)
d_s = torch.from_numpy(scal).cuda(); d_p = torch.from_numpy(pts).cuda(); torch.cuda.synchronize()
for c in ([int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16]):
    for K in ([int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [32, 64, 128]):
        lib.ctt_b200_set_tuning(c, 16, K)
        for _ in range(2): M.msm_device_ptrs(cv, d_s.data_ptr(), d_p.data_ptr(), n)
        acc = []
        for _ in range(5):
            M.msm_device_ptrs(cv, d_s.data_ptr(), d_p.data_ptr(), n); acc.append(M.last_stats())
        avg = {k: round(sum(a[k] for a in acc) / len(acc), 3) for k in acc[0] if k.startswith("ms_")}
        print(json.dumps({"curve": curve, "logn": logn, "c": c, "K": K, **avg}), flush=True)
