"""Developer script (not collected by pytest): first-light parity + timing of the CUDA path against the oracle."""
import os, sys, time, random, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constantine_b200.curves import CURVES
from constantine_b200 import msm as M, _lib
from oracle import pyref, oracle

random.seed(7)
ok = True
tp = M.Threadpool.new(4)
for cv in CURVES.values():
    base = [pyref.ec_mul_fast(random.getrandbits(64) | 1, cv.gen, cv) for _ in range(64)]
    for n in (1, 2, 3, 7, 33, 100, 1000, 5000):
        pts = [base[random.randrange(64)] if n > 64 else base[i] for i in range(n)]
        ks = [random.getrandbits(cv.scalar_bits) for _ in range(n)]
        if n >= 7:
            pts[2] = None; ks[3] = 0; ks[4] = 1; pts[5] = pts[6]; ks[5] = ks[6]
        cb = b"".join(pyref.scalar_to_bytes(k, cv) for k in ks)
        pb = b"".join(pyref.aff_to_bytes(P, cv) for P in pts)
        want = pyref.jac_bytes_to_affine(oracle.msm(cv, cb, pb, n), cv)
        t0 = time.time()
        got = pyref.jac_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, n, out="jac"), cv)
        dt = time.time() - t0
        gotp = pyref.prj_bytes_to_affine(M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, n, out="prj"), cv)
        st = M.last_stats()
        good = (got == want) and (gotp == want)
        ok &= good
        print(f"{cv.name:16s} n={n:5d} c={st['c']:2d} W={st['num_windows']:2d} {'OK ' if good else 'MISMATCH'} {dt*1e3:8.2f} ms  launches={st['kernel_launches']}", flush=True)
print("ALL OK" if ok else "FAILURES")

# timing: BLS12-381 G1, 2^16 .. 2^20 with a 4096-point pool
cv = CURVES["bls12_381_g1"]
pool = [pyref.ec_mul_fast(random.getrandbits(64) | 1, cv.gen, cv) for _ in range(512)]
poolb = [pyref.aff_to_bytes(P, cv) for P in pool]
for logn in (16, 18, 20):
    n = 1 << logn
    pb = b"".join(poolb[random.randrange(512)] for _ in range(n))
    cb = random.getrandbits(8 * 32 * n).to_bytes(32 * n, "little")
    cb = bytearray(cb)
    for i in range(n):
        cb[32 * i + 31] &= 0x7F
    cb = bytes(cb)
    for it in range(3):
        t0 = time.time()
        r = M.multi_scalar_mul_vartime_parallel(tp, cv, cb, pb, n)
        dt = time.time() - t0
        st = M.last_stats()
        print(f"bls12_381_g1 n=2^{logn} it={it} wall={dt*1e3:.2f} ms  " + " ".join(f"{k}={v:.3f}" if isinstance(v, float) else f"{k}={v}" for k, v in st.items()), flush=True)
    if logn <= 16:
        want = pyref.jac_bytes_to_affine(oracle.msm(cv, cb, pb, n), cv)
        print("   parity vs oracle:", pyref.jac_bytes_to_affine(r, cv) == want, flush=True)
