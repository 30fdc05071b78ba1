#!/usr/bin/env python3
"""Small MSMs on every curve, meant to run under compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import CURVES, pack, point_pool, pyref
from constantine_b200 import msm as M
from oracle import oracle
r = random.Random(3)
tp = M.Threadpool.new(1)
if len(sys.argv) > 1:      # forced batched-affine levels (msm_affine.cuh), e.g. `sanitize_small.py 2`
    from constantine_b200 import _lib
    _lib.load().ctt_b200_set_affine_levels(int(sys.argv[1]))
for cv in CURVES.values():
    _, pool = point_pool(cv, size=16)
    for n, same in ((700, False), (900, True)):
        pts = [pool[0] if same else pool[r.randrange(16)] for _ in range(n)]
        ks = [r.getrandbits(cv.scalar_bits) for _ in range(n)]
        cb, pb = pack(cv, ks, pts)
        got = pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, n), cv)
        want = pyref.jac_bytes_to_affine(oracle.msm(cv, cb, pb, n), cv)
        print(cv.name, n, "all-equal points" if same else "random", "OK" if got == want else "MISMATCH", flush=True)
