"""Instruction-level emulation of the specialised Montgomery multipliers of constantine_b200/csrc/field.cuh
(fe_dot2: two products under one reduction; fe_sqr: separate-operand-scanning squaring).

Every PTX helper the CUDA code uses (add.cc / addc.cc / addc, mad.lo.cc / madc.lo.cc / madc.hi.cc / madc.hi) is modelled on
32-bit limbs with an explicit carry flag, in exactly the order the CUDA source issues them. The two instructions that do
NOT write the carry flag (`addc`, `madc.hi`) assert that no carry would have been produced: a dropped carry is how a
carry-chain multiplier goes wrong silently for rare inputs, and random device tests would almost never hit it.
Test infrastructure only (tests/test_host_logic.py runs it on CPU for every base field)."""
M32 = 0xFFFFFFFF


class Flags:
    def __init__(self):
        self.c = 0

    def add_cc(self, a, b):
        t = a + b
        self.c = t >> 32
        return t & M32

    def addc_cc(self, a, b):
        t = a + b + self.c
        self.c = t >> 32
        return t & M32

    def addc(self, a, b):
        t = a + b + self.c
        assert t >> 32 == 0, "addc would drop a carry"
        self.c = 0
        return t & M32

    def mad_lo_cc(self, a, b, c):
        t = ((a * b) & M32) + c
        self.c = t >> 32
        return t & M32

    def madc_lo_cc(self, a, b, c):
        t = ((a * b) & M32) + c + self.c
        self.c = t >> 32
        return t & M32

    def madc_hi_cc(self, a, b, c):
        t = ((a * b) >> 32) + c + self.c
        self.c = t >> 32
        return t & M32

    def madc_hi(self, a, b, c):
        t = ((a * b) >> 32) + c + self.c
        assert t >> 32 == 0, "madc.hi would drop a carry"
        self.c = 0
        return t & M32


def limbs(x, n):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def value(l):
    return sum(v << (32 * i) for i, v in enumerate(l))


def field_consts(p, n):
    return limbs(p, n), (-pow(p, -1, 1 << 32)) & M32


def mont_round(E, O, n, P, inv, k):          # field.cuh mont_round
    m = (E[0] * inv) & M32
    O[0] = k.mad_lo_cc(m, P[1], O[0])
    O[1] = k.madc_hi_cc(m, P[1], O[1])
    for j in range(2, n, 2):
        O[j] = k.madc_lo_cc(m, P[j + 1], O[j])
        O[j + 1] = k.madc_hi_cc(m, P[j + 1], O[j + 1])
    assert k.c == 0, "carry out of the odd-limb chain in mont_round"
    E[0] = k.mad_lo_cc(m, P[0], E[0])
    E[1] = k.madc_hi_cc(m, P[0], E[1])
    for j in range(2, n, 2):
        E[j] = k.madc_lo_cc(m, P[j], E[j])
        E[j + 1] = k.madc_hi_cc(m, P[j], E[j + 1])
    O[n - 1] = k.addc(O[n - 1], 0)
    assert E[0] == 0


def add_product_row(E, O, c, di, n, k):      # field.cuh add_product_row
    O[0] = k.mad_lo_cc(c[1], di, O[0])
    O[1] = k.madc_hi_cc(c[1], di, O[1])
    for j in range(2, n - 2, 2):
        O[j] = k.madc_lo_cc(c[j + 1], di, O[j])
        O[j + 1] = k.madc_hi_cc(c[j + 1], di, O[j + 1])
    O[n - 2] = k.madc_lo_cc(c[n - 1], di, O[n - 2])
    O[n - 1] = k.madc_hi(c[n - 1], di, O[n - 1])
    E[0] = k.mad_lo_cc(c[0], di, E[0])
    E[1] = k.madc_hi_cc(c[0], di, E[1])
    for j in range(2, n, 2):
        E[j] = k.madc_lo_cc(c[j], di, E[j])
        E[j + 1] = k.madc_hi_cc(c[j], di, E[j + 1])
    O[n - 1] = k.addc(O[n - 1], 0)


def mont_step2(E, O, a, bi, c, di, n, P, inv, k):   # field.cuh mont_step2
    E[0] = k.add_cc(E[0], O[1])
    for j in range(0, n - 2, 2):
        O[j] = k.madc_lo_cc(a[j + 1], bi, O[j + 2])
        O[j + 1] = k.madc_hi_cc(a[j + 1], bi, O[j + 3])
    O[n - 2] = k.madc_lo_cc(a[n - 1], bi, 0)
    O[n - 1] = k.madc_hi(a[n - 1], bi, 0)
    E[0] = k.mad_lo_cc(a[0], bi, E[0])
    E[1] = k.madc_hi_cc(a[0], bi, E[1])
    for j in range(2, n, 2):
        E[j] = k.madc_lo_cc(a[j], bi, E[j])
        E[j + 1] = k.madc_hi_cc(a[j], bi, E[j + 1])
    O[n - 1] = k.addc(O[n - 1], 0)
    add_product_row(E, O, c, di, n, k)
    mont_round(E, O, n, P, inv, k)


def fe_dot2(p, n, a, b, c, d):
    """(a b + c d) R^-1 mod p as field.cuh fe_dot2 computes it; returns (canonical value, value before final_sub)."""
    P, inv = field_consts(p, n)
    a, b, c, d = (limbs(x, n) for x in (a, b, c, d))
    k = Flags()
    A, B = [0] * n, [0] * n
    for j in range(0, n, 2):
        A[j], A[j + 1] = (a[j] * b[0]) & M32, (a[j] * b[0]) >> 32
        B[j], B[j + 1] = (a[j + 1] * b[0]) & M32, (a[j + 1] * b[0]) >> 32
    add_product_row(A, B, c, d[0], n, k)
    mont_round(A, B, n, P, inv, k)
    for i in range(1, n):
        if i & 1:
            mont_step2(B, A, a, b[i], c, d[i], n, P, inv, k)
        else:
            mont_step2(A, B, a, b[i], c, d[i], n, P, inv, k)
    E, O = (B, A) if (n - 1) & 1 else (A, B)
    t = [0] * n
    t[0] = k.add_cc(O[0], E[1])
    for q in range(1, n - 1):
        t[q] = k.addc_cc(O[q], E[q + 1])
    t[n - 1] = k.addc(O[n - 1], 0)
    v = value(t)
    assert v < 3 * p, "fe_dot2 result not below 3p (two final subtractions would not suffice)"
    return v % p, v


def fe_sqr(p, n, a):
    """a^2 R^-1 mod p as field.cuh fe_sqr computes it; returns (canonical value, number of 32x32->64 multiply-accumulates)."""
    P, inv = field_consts(p, n)
    a = limbs(a, n)
    k = Flags()
    X, Y = [0] * (2 * n), [0] * (2 * n)
    macs = 0
    for i in range(n - 1):
        for acc, j0 in ((X, i + 1), (Y, i + 2)):
            end = None
            for j in range(j0, n, 2):
                acc[i + j] = k.mad_lo_cc(a[i], a[j], acc[i + j]) if j == j0 else k.madc_lo_cc(a[i], a[j], acc[i + j])
                acc[i + j + 1] = k.madc_hi_cc(a[i], a[j], acc[i + j + 1])
                end = i + j + 1
                macs += 1
            if end is not None:
                acc[end + 1] = k.addc(acc[end + 1], 0)
    T = [0] * (2 * n)
    T[0] = k.add_cc(X[0], Y[0])
    for q in range(1, 2 * n - 1):
        T[q] = k.addc_cc(X[q], Y[q])
    T[2 * n - 1] = k.addc(X[2 * n - 1], Y[2 * n - 1])
    T[0] = k.add_cc(T[0], T[0])
    for q in range(1, 2 * n - 1):
        T[q] = k.addc_cc(T[q], T[q])
    T[2 * n - 1] = k.addc(T[2 * n - 1], T[2 * n - 1])
    for i in range(n):
        T[2 * i] = k.mad_lo_cc(a[i], a[i], T[2 * i]) if i == 0 else k.madc_lo_cc(a[i], a[i], T[2 * i])
        T[2 * i + 1] = k.madc_hi_cc(a[i], a[i], T[2 * i + 1]) if i < n - 1 else k.madc_hi(a[i], a[i], T[2 * i + 1])
        macs += 1
    assert value(T) == value(a) ** 2
    C = [0] * (n + 1)
    for i in range(n):
        m = (T[i] * inv) & M32
        T[i] = k.mad_lo_cc(m, P[0], T[i])
        T[i + 1] = k.madc_hi_cc(m, P[0], T[i + 1])
        for j in range(2, n, 2):
            T[i + j] = k.madc_lo_cc(m, P[j], T[i + j])
            T[i + j + 1] = k.madc_hi_cc(m, P[j], T[i + j + 1])
        C[i] = k.addc(C[i], 0)
        T[i + 1] = k.mad_lo_cc(m, P[1], T[i + 1])
        T[i + 2] = k.madc_hi_cc(m, P[1], T[i + 2])
        for j in range(3, n, 2):
            T[i + j] = k.madc_lo_cc(m, P[j], T[i + j])
            T[i + j + 1] = k.madc_hi_cc(m, P[j], T[i + j + 1])
        C[i + 1] = k.addc(C[i + 1], 0)
        assert T[i] == 0
        macs += n
    t = [0] * n
    t[0] = k.add_cc(T[n], C[0])
    for q in range(1, n - 1):
        t[q] = k.addc_cc(T[n + q], C[q])
    t[n - 1] = k.addc(T[2 * n - 1], C[n - 1])
    assert C[n] == 0
    v = value(t)
    assert v < 2 * p
    return v % p, macs


# ---------------------------------------------------------------------------------------------------------------
