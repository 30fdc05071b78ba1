"""CPU model of the batched-affine level schedule of constantine_b200/csrc/msm_affine.cuh, statement for statement:
k_bucket_bounds -> level offsets -> k_affine_plan -> k_affine_pairs (per-thread batches with ONE shared inversion each,
prefix products, special cases) -> survivor list. Exact arithmetic (oracle/pyref.py), tiny sizes.

Checks, for random and adversarial runs (single entries, P + P, P - P, infinity operands, one giant run):
  * every slot of every level is written exactly once;
  * per bucket, the sum of its survivors equals the sum of its entries;
  * the shared inversion is used once per thread and level, never on a zero.
Run: python tools/proto_affine_levels.py      (also imported by tests/test_host_logic.py)"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from constantine_b200.curves import CURVES  # noqa: E402
from oracle import pyref  # noqa: E402

NONE = 0xFFFFFFFF


def level_count(n, r):
    return (n + (1 << r) - 1) >> r


def build_plan(keys, vals, no_key, nb, L):
    n = len(keys)
    head, tail = [0] * nb, [0] * nb
    for q in range(n):                                      # k_bucket_bounds
        k = keys[q]
        if k >= no_key:
            continue
        if q == 0 or keys[q - 1] != k:
            head[k] = q
        if q + 1 == n or keys[q + 1] != k:
            tail[k] = q + 1
    off = [[0] * (nb + 1) for _ in range(L + 1)]             # k_level_blocksums / scan / offsets
    for r in range(L + 1):
        acc = 0
        for b in range(nb):
            off[r][b] = acc
            acc += level_count(tail[b] - head[b], r)
        off[r][nb] = acc
    plan0 = [None] * off[1][nb] if L >= 1 else []
    plan = [None] + [[None] * off[r + 1][nb] for r in range(1, L)]
    surv_keys, surv_vals = [None] * off[L][nb], [None] * off[L][nb]
    for q in range(n):                                      # k_affine_plan
        b = keys[q]
        if b >= no_key:
            continue
        h = head[b]
        i, cnt = q - h, tail[b] - h
        for r in range(L):
            if i & ((2 << r) - 1):
                break
            p = off[r + 1][b] + (i >> (r + 1))
            s = i >> r
            has2 = s + 1 < level_count(cnt, r)
            if r == 0:
                assert plan0[p] is None
                plan0[p] = (vals[q], vals[q + 1] if has2 else NONE)
            else:
                assert plan[r][p] is None
                plan[r][p] = (off[r][b] + s) | (0x80000000 if has2 else 0)
        if i & ((1 << L) - 1) == 0:
            ps = off[L][b] + (i >> L)
            assert surv_keys[ps] is None
            surv_keys[ps], surv_vals[ps] = b, ps
    assert all(x is not None for x in plan0) and all(x is not None for lv in plan[1:] for x in lv)
    assert all(x is not None for x in surv_keys)
    return off, plan0, plan, surv_keys, surv_vals


def affine_pairs(cv, first, plan, total, src, threads, stats):
    """k_affine_pairs: `threads` lanes (multiple of 32), slots split evenly, warp-contiguous ranges."""
    p_mod = cv.fp.modulus
    F = pyref
    dst = [None] * total
    M = (total + threads - 1) // threads

    def operand(ref):
        if first:
            P = src[ref & 0x7FFFFFFF]
            return pyref.ec_neg(P, cv) if (ref >> 31) and P is not None else P
        return src[ref]

    def task(p):
        if first:
            return plan[p]
        v = plan[p]
        a = v & 0x7FFFFFFF
        return (a, a + 1 if v >> 31 else NONE)

    def classify(single, P1, P2):
        if single:
            return "copy1", None
        if P1 is None:
            return "copy2", None
        if P2 is None:
            return "copy1", None
        den = F.f_sub(P2[0], P1[0], p_mod)
        if not F.f_is_zero(den):
            return "add", den
        if P1[1] != P2[1] or F.f_is_zero(P1[1]):
            return "inf", None
        return "dbl", F.f_add(P1[1], P1[1], p_mod)

    one = (1,) + (0,) * (cv.ext_degree - 1)
    for tid in range(threads):
        lane = tid & 31
        warp_base = (tid - lane) * M
        if warp_base >= total:
            continue
        first_slot = warp_base + lane
        cnt = 0
        if first_slot < total:
            cnt = min(M, (total - first_slot + 31) // 32)
        run, prefix = one, []
        for j in range(cnt):
            a, b = task(first_slot + 32 * j)
            kind, den = classify(b == NONE, operand(a), operand(b) if b != NONE else None)
            if kind in ("add", "dbl"):
                run = F.f_mul(run, den, p_mod)
            prefix.append(run)
        assert not F.f_is_zero(run)
        inv = F.f_inv(run, p_mod)
        stats["inversions"] += 1
        for j in range(cnt - 1, -1, -1):
            p = first_slot + 32 * j
            a, b = task(p)
            P1, P2 = operand(a), (operand(b) if b != NONE else None)
            kind, den = classify(b == NONE, P1, P2)
            if kind == "copy1":
                R = P1
            elif kind == "copy2":
                R = P2
            elif kind == "inf":
                R = None
            else:
                inv_den = inv if j == 0 else F.f_mul(inv, prefix[j - 1], p_mod)
                inv = F.f_mul(inv, den, p_mod)
                if kind == "add":
                    num = F.f_sub(P2[1], P1[1], p_mod)
                else:
                    xx = F.f_mul(P1[0], P1[0], p_mod)
                    num = F.f_add(F.f_add(xx, xx, p_mod), xx, p_mod)
                lam = F.f_mul(num, inv_den, p_mod)
                x3 = F.f_sub(F.f_sub(F.f_mul(lam, lam, p_mod), P1[0], p_mod), P2[0], p_mod)
                y3 = F.f_sub(F.f_mul(lam, F.f_sub(P1[0], x3, p_mod), p_mod), P1[1], p_mod)
                R = (x3, y3)
                stats["adds"] += 1
            assert dst[p] is None
            dst[p] = R
            stats["slots"] += 1
    assert stats["slots"] >= total
    return dst


def run_case(cv, keys, vals, points, no_key, nb, L, threads=64):
    off, plan0, plan, skeys, svals = build_plan(keys, vals, no_key, nb, L)
    stats = {"inversions": 0, "adds": 0, "slots": 0}
    work = points
    for r in range(L):
        total = off[r + 1][nb]
        work = affine_pairs(cv, r == 0, plan0 if r == 0 else plan[r], total, work, threads, stats)
        assert all(True for _ in work)
    # per bucket: survivors sum == entries sum
    want = {}
    for k, v in zip(keys, vals):
        if k >= no_key:
            continue
        P = points[v & 0x7FFFFFFF]
        if v >> 31 and P is not None:
            P = pyref.ec_neg(P, cv)
        want[k] = pyref.ec_add(want.get(k), P, cv)
    got = {}
    for k, ps in zip(skeys, svals):
        got[k] = pyref.ec_add(got.get(k), work[ps] if L else None, cv)
    for k in want:
        assert got.get(k) == want[k], ("bucket", k)
    return stats


def self_test(seed=5):
    cv = CURVES["bn254_snarks_g1"]
    rnd = random.Random(seed)
    pool = [pyref.ec_mul_fast(rnd.getrandbits(40) | 1, cv.gen, cv) for _ in range(24)] + [None]
    total = {"inversions": 0, "adds": 0, "slots": 0}
    for trial, (nb, n, L) in enumerate([(7, 90, 1), (7, 90, 3), (16, 400, 4), (3, 200, 5), (40, 60, 2), (1, 129, 3)]):
        ents = []
        for _ in range(n):
            k = rnd.randrange(nb + 1) if trial != 5 else 0      # key nb = "no bucket" (zero digit)
            ents.append((k, rnd.randrange(len(pool)) | (rnd.getrandbits(1) << 31)))
        ents.sort(key=lambda e: e[0])
        keys, vals = [e[0] for e in ents], [e[1] for e in ents]
        st = run_case(cv, keys, vals, pool, nb, nb, L, threads=32 * (1 + trial % 3))
        for k in total:
            total[k] += st[k]
    return total


if __name__ == "__main__":
    print(self_test())
