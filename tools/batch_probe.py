"""Small sum-reduce / batch / fixed-base calls on every group, checked against the exact tier -- meant to be run under
compute-sanitizer (memcheck / racecheck) when chasing a device fault:
    compute-sanitizer --tool memcheck python tools/batch_probe.py [curve ...]"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import CURVES, pack, point_pool, pyref  # noqa: E402
from constantine_b200 import msm as M  # noqa: E402

rng = random.Random(3)
bad = 0
for name in sys.argv[1:] or list(CURVES):
    cv = CURVES[name]
    ks_pool, pool = point_pool(cv)
    for n in (1, 10, 257, 1500):
        idx = [rng.randrange(len(pool)) for _ in range(n)]
        pts = [pool[i] for i in idx]
        want = pyref.ec_mul_fast(sum(ks_pool[i] for i in idx) % cv.fr.modulus, cv.gen, cv)
        _, pb = pack(cv, [], pts)
        ok = pyref.jac_bytes_to_affine(M.sum_reduce_vartime(cv, pb, n), cv) == want
        bad += not ok
        print(name, "sum", n, "ok" if ok else "MISMATCH", flush=True)
    for batch, n in ((5, 7), (3, 40)):
        idx = [rng.randrange(len(pool)) for _ in range(batch * n)]
        pts = [pool[i] for i in idx]
        ks = [rng.getrandbits(cv.scalar_bits) for _ in range(batch * n)]
        cb, pb = pack(cv, ks, pts)
        want = [pyref.ec_mul_fast(sum(ks[m * n + i] * ks_pool[idx[m * n + i]] for i in range(n)) % cv.fr.modulus, cv.gen, cv)
                for m in range(batch)]
        got = [pyref.jac_bytes_to_affine(g, cv) for g in M.msm_batch(cv, cb, pb, batch, n)]
        bank = M.PrecomputedMSMBank(cv, pb, batch, n)
        got2 = [pyref.jac_bytes_to_affine(g, cv) for g in bank.msm_vartime(cb)]
        bank.free()
        ok = got == want and got2 == want
        bad += not ok
        print(name, "batch", batch, n, "ok" if ok else "MISMATCH", flush=True)
sys.exit(1 if bad else 0)
