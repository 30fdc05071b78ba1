"""Schedule prototype for batched-affine bucket accumulation -- exact arithmetic, CPU only.

Checks the *logic* the CUDA kernels k_affine_level / k_survivors implement, on the exact big-int tier (oracle/pyref.py):
  * sorted (key, ref) list, a bucket = a run of equal keys; start[q] = first position of q's run;
  * level r (0, 1, 2): position q is a pair head iff (q - start[q]) % 2^(r+1) == 0; its partner is q + 2^r when that is
    still inside the run (same start), otherwise the item just carries over; results live in place at the head position;
  * within a thread block the pair denominators are inverted together (Montgomery's trick); infinity operands, P + P and
    P - P are classified first and never enter the shared product with a zero;
  * after LEVELS levels the survivors are the positions at run offsets that are multiples of 2^LEVELS: a 8x shorter sorted
    list that goes through the ordinary (XYZZ) accumulation.
Run: python tools/proto_batched_affine.py   (also imported by tests/test_host_logic.py)"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

LEVELS = 3


def run_starts(keys):
    start, s = [], 0
    for q, k in enumerate(keys):
        if q == 0 or k != keys[q - 1]:
            s = q
        start.append(s)
    return start


def classify(P1, P2, cv):
    """-> ('copy', point) | ('inf',) | ('add', num, den)   with lambda = num / den for the shared-inversion path"""
    p = cv.fp.modulus
    if P1 is None:
        return ("copy", P2)
    if P2 is None:
        return ("copy", P1)
    (x1, y1), (x2, y2) = P1, P2
    if x1 == x2:
        if y1 == y2 and not pyref.f_is_zero(y1):
            xx = pyref.f_mul(x1, x1, p)
            return ("add", pyref.f_add(pyref.f_add(xx, xx, p), xx, p), pyref.f_add(y1, y1, p))   # doubling: 3 x^2 / 2 y
        return ("inf",)
    return ("add", pyref.f_sub(y2, y1, p), pyref.f_sub(x2, x1, p))


def batch_level(items, start, total, r, cv, block_pairs=64):
    """One level, in place. items[q] = affine point or None. Pairs are inverted in blocks of `block_pairs` (the thread
    block's shared inversion)."""
    p = cv.fp.modulus
    step = 1 << r
    pend = []          # (q, P1, P2, num, den)
    for q in range(total):
        if (q - start[q]) % (2 * step) != 0:
            continue
        partner = q + step
        if partner >= total or start[partner] != start[q]:
            continue                                   # no partner inside the run: the item carries over
        kind = classify(items[q], items[partner], cv)
        if kind[0] == "copy":
            items[q] = kind[1]
        elif kind[0] == "inf":
            items[q] = None
        else:
            pend.append((q, items[q], items[partner], kind[1], kind[2]))
        items[partner] = "dead"
    for b in range(0, len(pend), block_pairs):
        blk = pend[b:b + block_pairs]
        prefix, acc = [], None
        for (_, _, _, _, den) in blk:                  # running products
            acc = den if acc is None else pyref.f_mul(acc, den, p)
            prefix.append(acc)
        inv = pyref.f_inv(acc, p)                      # THE one inversion of the block
        for i in range(len(blk) - 1, -1, -1):          # unwind
            q, P1, P2, num, den = blk[i]
            inv_den = inv if i == 0 else pyref.f_mul(inv, prefix[i - 1], p)
            inv = pyref.f_mul(inv, den, p)
            lam = pyref.f_mul(num, inv_den, p)
            x3 = pyref.f_sub(pyref.f_sub(pyref.f_mul(lam, lam, p), P1[0], p), P2[0], p)
            y3 = pyref.f_sub(pyref.f_mul(lam, pyref.f_sub(P1[0], x3, p), p), P1[1], p)
            items[q] = (x3, y3)
    return len(pend)


def accumulate(keys, refs, points, cv):
    """Returns {key: bucket sum} via LEVELS batched-affine levels + plain summation of the survivors, plus statistics."""
    total = len(keys)
    start = run_starts(keys)
    items = []
    for ref in refs:                                   # level-0 gather with the sign applied
        P = points[ref & 0x7FFFFFFF]
        items.append(pyref.ec_neg(P, cv) if (ref >> 31) and P is not None else P)
    adds = 0
    for r in range(LEVELS):
        adds += batch_level(items, start, total, r, cv)
    buckets, survivors = {}, 0
    for q in range(total):
        if (q - start[q]) % (1 << LEVELS) == 0:
            assert items[q] != "dead"
            survivors += 1
            buckets[keys[q]] = pyref.ec_add(buckets.get(keys[q]), items[q], cv)
        else:
            assert items[q] == "dead"
    return buckets, adds, survivors


def self_check(curve_name="bn254_snarks_g1", n=600, nbuckets=23, seed=1):
    from constantine_b200.curves import CURVES
    cv = CURVES[curve_name]
    rnd = random.Random(seed)
    base = [pyref.ec_mul_fast(rnd.getrandbits(64) | 1, cv.gen, cv) for _ in range(12)]
    points = [base[rnd.randrange(len(base))] for _ in range(n)]   # few distinct points: P + P and P - P occur often
    points[5] = None
    points[17] = None
    entries = sorted(((rnd.randrange(nbuckets), i | (rnd.randrange(2) << 31)) for i in range(n)), key=lambda e: e[0])
    keys, refs = [e[0] for e in entries], [e[1] for e in entries]
    got, adds, survivors = accumulate(keys, refs, points, cv)
    want = {}
    for k, ref in zip(keys, refs):
        P = points[ref & 0x7FFFFFFF]
        P = pyref.ec_neg(P, cv) if (ref >> 31) and P is not None else P
        want[k] = pyref.ec_add(want.get(k), P, cv)
    assert got == want
    return adds, survivors, n


if __name__ == "__main__":
    for name in ("bn254_snarks_g1", "bls12_381_g1", "bls12_381_g2"):
        adds, survivors, n = self_check(name)
        print(f"{name}: {n} entries -> {adds} batched additions, {survivors} survivors: bucket sums exact")


# ---- thread-accurate simulation of k_affine_level (msm_kernels.cuh): same chunking, same capacity bound, same block scan ----
def simulate_kernel_level(items, start, total, r, cv, cap=32, threads=128):
    """items: in-place work list (level 0: already gathered). Mirrors the CUDA kernel phase by phase: per-thread pass 1,
    block-wide inclusive prefix / suffix products (Hillis-Steele), one inversion per block, per-thread unwind."""
    p = cv.fp.modulus
    one = (1,) + (0,) * (cv.ext_degree - 1)
    step = 1 << r
    pc = (cap - 1) * (step + 1)
    nthreads = (total + pc - 1) // pc
    nblocks = (nthreads + threads - 1) // threads
    inversions = 0
    for blk in range(nblocks):
        state = []
        for tid in range(threads):                      # pass 1
            lo = (blk * threads + tid) * pc
            hi = min(lo + pc, total)
            qs, prefix, run = [], [], one
            for q in range(lo, hi):
                s = start[q]
                if (q - s) & (2 * step - 1):
                    continue
                partner = q + step
                if partner >= total or start[partner] != s:
                    continue                            # (level 0 would copy the single item into the work array here)
                kind = classify(items[q], items[partner], cv)
                if kind[0] == "copy":
                    items[q] = kind[1]
                    continue
                if kind[0] == "inf":
                    items[q] = None
                    continue
                run = pyref.f_mul(run, kind[2], p)
                assert len(qs) < cap, "pair capacity of a thread exceeded"
                qs.append(q)
                prefix.append(run)
            state.append((qs, prefix, run))
        if not any(st[0] for st in state):
            continue
        pre = [st[2] for st in state]                   # Hillis-Steele inclusive scans, as in the kernel
        suf = [st[2] for st in state]
        d = 1
        while d < threads:
            a = [pre[t - d] if t >= d else None for t in range(threads)]
            b = [suf[t + d] if t + d < threads else None for t in range(threads)]
            pre = [pyref.f_mul(pre[t], a[t], p) if a[t] is not None else pre[t] for t in range(threads)]
            suf = [pyref.f_mul(suf[t], b[t], p) if b[t] is not None else suf[t] for t in range(threads)]
            d <<= 1
        inv_total = pyref.f_inv(pre[threads - 1], p)
        inversions += 1
        for tid in range(threads):                      # pass 2
            qs, prefix, run = state[tid]
            if not qs:
                continue
            inv = inv_total
            if tid > 0:
                inv = pyref.f_mul(inv, pre[tid - 1], p)
            if tid + 1 < threads:
                inv = pyref.f_mul(inv, suf[tid + 1], p)
            assert pyref.f_mul(inv, run, p) == one
            for i in range(len(qs) - 1, -1, -1):
                q = qs[i]
                P1, P2 = items[q], items[q + step]
                _, num, den = classify(P1, P2, cv)
                inv_den = inv if i == 0 else pyref.f_mul(inv, prefix[i - 1], p)
                inv = pyref.f_mul(inv, den, p)
                lam = pyref.f_mul(num, inv_den, p)
                x3 = pyref.f_sub(pyref.f_sub(pyref.f_mul(lam, lam, p), P1[0], p), P2[0], p)
                y3 = pyref.f_sub(pyref.f_mul(lam, pyref.f_sub(P1[0], x3, p), p), P1[1], p)
                items[q] = (x3, y3)
    return inversions


def self_check_kernel_shape(curve_name="bn254_snarks_g1", n=900, nbuckets=40, seed=3, cap=4, threads=8):
    """Small capacities so that chunk boundaries, the capacity bound and the block scan are all exercised."""
    from constantine_b200.curves import CURVES
    cv = CURVES[curve_name]
    rnd = random.Random(seed)
    base = [pyref.ec_mul_fast(rnd.getrandbits(64) | 1, cv.gen, cv) for _ in range(9)]
    points = [base[rnd.randrange(len(base))] for _ in range(n)]
    points[7] = None
    entries = sorted(((min(rnd.randrange(nbuckets), rnd.randrange(nbuckets)), i | (rnd.randrange(2) << 31)) for i in range(n)),
                     key=lambda e: e[0])
    keys, refs = [e[0] for e in entries], [e[1] for e in entries]
    total = len(keys)
    start = run_starts(keys)
    items = []
    for ref in refs:
        P = points[ref & 0x7FFFFFFF]
        items.append(pyref.ec_neg(P, cv) if (ref >> 31) and P is not None else P)
    want = {}
    for k, it in zip(keys, items):
        want[k] = pyref.ec_add(want.get(k), it, cv)
    inv = 0
    for r in range(LEVELS):
        inv += simulate_kernel_level(items, start, total, r, cv, cap=cap, threads=threads)
    got = {}
    for q in range(total):
        if (q - start[q]) % (1 << LEVELS) == 0:
            got[keys[q]] = pyref.ec_add(got.get(keys[q]), items[q], cv)
    assert got == want
    return inv
