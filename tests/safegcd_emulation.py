"""Limb-level model of constantine_b200/csrc/field_inv.cuh (Bernstein-Yang "safegcd" inversion, batches of 30 divsteps,
signed 30-bit limbs, 32/64-bit machine integers). Every intermediate is checked against the width of the register the
CUDA code keeps it in, so a passing run means the device code cannot overflow on these inputs either.

The reference computes its variable-time inverses with the same family of algorithms (Bernstein-Yang divsteps with batched
transition matrices, reference constantine/math/arithmetic/limbs_exgcd.nim:708-876); this is an independent restatement from the
paper ("Fast constant-time gcd computation and modular inversion", 2019) for 32-bit lanes.
Run by the CPU suite (tests/test_host_logic.py::test_safegcd_model)."""
import random

M30 = (1 << 30) - 1


def i32(x):
    assert -(1 << 31) <= x < (1 << 31), "int32 overflow: %d" % x
    return x


def i64(x):
    assert -(1 << 63) <= x < (1 << 63), "int64 overflow: %d" % x
    return x


def to_signed30(x, n):
    out = []
    for _ in range(n):
        out.append(x & M30)
        x >>= 30
    assert x == 0
    return out


def from_signed30(v):
    return sum(l << (30 * i) for i, l in enumerate(v))


def divsteps_30(zeta, f0, g0):
    """30 divsteps on the low 32 bits of f and g (two's complement). Returns (zeta, (u, v, q, r)), matrix scaled by 2^30."""
    u, v, q, r = 1, 0, 0, 1
    f, g = f0 & 0xFFFFFFFF, g0 & 0xFFFFFFFF

    def s32(x):  # wrap to int32 (the device lets these wrap: only the low bits of f, g matter)
        x &= 0xFFFFFFFF
        return x - (1 << 32) if x >> 31 else x

    f, g = s32(f), s32(g)
    for _ in range(30):
        c1 = -1 if zeta < 0 else 0
        x, y, z = s32((f ^ c1) - c1), i32((u ^ c1) - c1), i32((v ^ c1) - c1)
        c2 = -(g & 1)
        g, q, r = s32(g + (x & c2)), i32(q + (y & c2)), i32(r + (z & c2))
        c1 &= c2
        zeta = i32((zeta ^ c1) - 1)
        f, u, v = s32(f + (g & c1)), i32(u + (q & c1)), i32(v + (r & c1))
        g, u, v = g >> 1, i32(u << 1), i32(v << 1)
    return zeta, (u, v, q, r)


def update_fg(f, g, t):
    u, v, q, r = t
    n = len(f)
    cf = i64(u * f[0] + v * g[0])
    cg = i64(q * f[0] + r * g[0])
    assert cf & M30 == 0 and cg & M30 == 0
    cf >>= 30
    cg >>= 30
    for i in range(1, n):
        cf = i64(cf + u * f[i] + v * g[i])
        cg = i64(cg + q * f[i] + r * g[i])
        f[i - 1] = cf & M30
        g[i - 1] = cg & M30
        cf >>= 30
        cg >>= 30
    f[n - 1] = i32(cf)
    g[n - 1] = i32(cg)


def update_de(d, e, t, m, m_inv30):
    u, v, q, r = t
    n = len(d)
    sd = -1 if d[n - 1] < 0 else 0
    se = -1 if e[n - 1] < 0 else 0
    md = i32((u & sd) + (v & se))
    me = i32((q & sd) + (r & se))
    cd = i64(u * d[0] + v * e[0])
    ce = i64(q * d[0] + r * e[0])
    md = i32(md - ((m_inv30 * (cd & 0xFFFFFFFF) + md) & M30))
    me = i32(me - ((m_inv30 * (ce & 0xFFFFFFFF) + me) & M30))
    cd = i64(cd + m[0] * md)
    ce = i64(ce + m[0] * me)
    assert cd & M30 == 0 and ce & M30 == 0
    cd >>= 30
    ce >>= 30
    for i in range(1, n):
        cd = i64(cd + u * d[i] + v * e[i] + m[i] * md)
        ce = i64(ce + q * d[i] + r * e[i] + m[i] * me)
        d[i - 1] = cd & M30
        e[i - 1] = ce & M30
        cd >>= 30
        ce >>= 30
    d[n - 1] = i32(cd)
    e[n - 1] = i32(ce)


def normalize(v, neg, m):
    """v in (-2M, M) as signed30 with a signed top limb -> canonical [0, M); negated first when neg."""
    n = len(v)
    mask_add = -1 if v[n - 1] < 0 else 0
    mask_neg = -1 if neg else 0
    for i in range(n):
        x = v[i] + (m[i] & mask_add)
        v[i] = i32((x ^ mask_neg) - mask_neg)
    for i in range(n - 1):
        v[i + 1] = i32(v[i + 1] + (v[i] >> 30))
        v[i] &= M30
    mask_add = -1 if v[n - 1] < 0 else 0
    for i in range(n):
        v[i] = i32(v[i] + (m[i] & mask_add))
    for i in range(n - 1):
        v[i + 1] = i32(v[i + 1] + (v[i] >> 30))
        v[i] &= M30
    return v


def num_limbs30(bits):
    return (bits + 2 + 29) // 30   # room for (-2M, M)


def max_batches(bits):
    # half-delta divsteps bound for inputs below 2^bits (safegcd bounds, Pornin / Wuille): floor((45907 bits + 26313) / 19929)
    steps = (45907 * bits + 26313) // 19929
    return (steps + 29) // 30


def modinv_scaled(x, m_int, factor, bits):
    """factor * x^-1 mod m for 0 < x < m (m odd prime); returns (result, batches used)."""
    n = num_limbs30(bits)
    m = to_signed30(m_int, n)
    m_inv30 = pow(m_int, -1, 1 << 30)
    d = [0] * n
    e = to_signed30(factor, n)
    f = to_signed30(m_int, n)
    g = to_signed30(x, n)
    zeta = -1
    used = 0
    for _ in range(max_batches(bits)):
        zeta, t = divsteps_30(zeta, f[0] | (f[1] << 30), g[0] | (g[1] << 30))
        update_de(d, e, t, m, m_inv30)
        update_fg(f, g, t)
        used += 1
        if all(l == 0 for l in g):
            break
    assert all(l == 0 for l in g), "g != 0 after the bound"
    fv = from_signed30(f)
    assert fv in (1, -1), fv
    dv = from_signed30(d)
    assert -2 * m_int < dv < m_int
    r = normalize(d, f[n - 1] < 0, m)
    return from_signed30(r), used


def self_test(moduli, samples=200, seed=7):
    rnd = random.Random(seed)
    worst = {}
    for name, (p, bits) in moduli.items():
        r2 = pow(2, 2 * 32 * ((bits + 63) // 64 * 2), p)   # R^2 with R = 2^(64 * limbs64)
        mx = 0
        xs = [1, 2, p - 1, p - 2, (p + 1) // 2] + [rnd.randrange(1, p) for _ in range(samples)]
        for x in xs:
            got, used = modinv_scaled(x, p, r2, bits)
            assert got == r2 * pow(x, -1, p) % p, (name, x)
            mx = max(mx, used)
        worst[name] = (mx, max_batches(bits))
    return worst


if __name__ == "__main__":
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from constantine_b200.curves import FIELDS
    print(self_test({k: (f.modulus, f.bits) for k, f in FIELDS.items()}, samples=300))
