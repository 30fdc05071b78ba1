#!/usr/bin/env python3
"""Cached bases with and without the precomputed window table, BLS12-381 G1 (what the ZAL msm_with_cached_base hook
would call).  python tools/bench_cached.py [logn]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from constantine_b200 import msm as M
from constantine_b200.curves import CURVES
from oracle import pyref
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cv = CURVES["bls12_381_g1"]; n = 1 << logn
scal, pts, k = bench.make_inputs(n, 77)
bases = M.CachedBases(cv, pts, n)
ref = None
for mode in ("plain", "table"):
    c_used = None
    if mode == "table":
        t0 = time.perf_counter(); c_used = bases.precompute(0); t_pre = time.perf_counter() - t0
    for _ in range(2): r = bases.msm(scal, n)
    t0 = time.perf_counter(); st = []
    for _ in range(5):
        r = bases.msm(scal, n); st.append(M.last_stats())
    wall = (time.perf_counter() - t0) / 5 * 1e3
    aff = pyref.jac_bytes_to_affine(r, cv)
    if ref is None: ref = aff
    avg = {kk: round(sum(s[kk] for s in st) / 5, 3) for kk in st[0] if kk.startswith("ms_")}
    print(json.dumps({"mode": mode, "logn": logn, "c": st[-1]["c"], "wall_ms_incl_scalar_h2d": round(wall, 3), "same_point": aff == ref,
                      "precompute_s": round(t_pre, 3) if mode == "table" else None, **avg}), flush=True)
