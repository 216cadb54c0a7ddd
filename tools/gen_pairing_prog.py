#!/usr/bin/env python3
"""Generate keaki_b200/csrc/pairing_prog_gen.cuh: the straight-line Fq2 programs the pairing VM
(keaki_b200/csrc/pairing_vm.cuh) executes, one pairing per thread.

Why a program: the optimal-ate pairing on BN254 with arkworks' final exponentiation
(`E::pairing`, src/kem.rs:30,58, src/kzg.rs:148 of the reference) is data-independent - the ate
loop digits and the bits of z are constants - so the whole computation is ONE fixed sequence of
Fq2 operations.  An Fq12 does not fit a thread's registers (96 of 255), so instead of letting the
compiler spill it to local memory the kernel keeps a per-thread file of Fq2 *slots* in shared
memory and interprets this sequence; the register allocation below (furthest-next-use eviction,
spills to an L2-resident global scratch) is done here, offline.

Standalone: plain Python integers, no import of oracle/.  tests/test_pairing_prog.py simulates the
emitted programs with `simulate()` and compares with the oracle's pairing bytes; tests/hostemu runs
them through the very interpreter the GPU runs.

Run:  python tools/gen_pairing_prog.py
"""
import os
import sys

Z = 4965661367192848881
Q = 36 * Z**4 + 36 * Z**3 + 24 * Z**2 + 6 * Z + 1
MONT = 1 << 256
ATE = [0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0,
       1, 1, 1, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0,
       0, 1, 0, 1, 1]
assert sum(d << i for i, d in enumerate(ATE)) == 6 * Z + 2

# ---------------------------------------------------------------------------------------------
# VM instruction set.  One instruction = two 64-bit words:
#   w0 = op | dst << 8 | A1 << 16 | A2 << 24 | B1 << 32 | B2 << 40 | F << 48 | G << 56
#   w1 = X0 | X1 << 8 | ... | X7 << 56                       (extra slot numbers of the macro operations)
# An operand byte is  slot | code << 5  (0xFF = absent); the code applies a cheap linear map while
# the operand is fetched: +x, -x, 2x, -2x, xi x, -xi x, 3x, 3 xi x (slot 31 is never used, so 0xFF is free).
#   MULX : dst = (A1 + A2) * (B1 + B2) + F + G        (Fq2 product)
#   SQRX : dst = (A1 + A2)^2 + F + G
#   MULF0: dst = (A1 + A2) * (B1).c0 + F + G          (Fq2 times an Fq scalar held in half a slot)
#   MULF1: dst = (A1 + A2) * (B1).c1 + F + G
#   LIN  : dst = A1 + A2 + B1 + B2 + F + G
#   CONJ : dst = conj(A1 + A2);   INV: dst = 1 / (A1 + A2)
#   LDC  : dst = constant[A1 | A2 << 8];  LDG: dst = global[A1 | A2 << 8];  STG: global[A1 | A2 << 8] = slot B1
# Macro operations (plain slot numbers, no codes) keep their intermediates in registers - no slot
# traffic, decoding or operand mapping per inner product:
#   F6MUL : (dst, X0, X1) = (A1, A2, B1) * (B2, F, G) in Fq6 = Fq2[v]/(v^3 - xi)            6 products
#   CYCSQR: (dst, X0..X4) = (A1, A2, B1, B2, F, G)^2 for a cyclotomic Fq12 element, w-basis    9 squarings
#   F6M01 : (dst, X0, X1) = (A1, A2, B1) * (B2 + F v)                                          5 products
#   F6SCL : (dst, X0, X1) = (A1, A2, B1) * B2   (Fq6 times an Fq2 scalar)                      3 products
# ---------------------------------------------------------------------------------------------
OPS = ["END", "MULX", "SQRX", "MULF0", "MULF1", "LIN", "CONJ", "INV", "LDC", "LDG", "STG", "NOP", "F6MUL", "CYCSQR", "F6M01", "F6SCL"]
OP = {n: i for i, n in enumerate(OPS)}
MACROS = ("F6MUL", "CYCSQR", "F6M01", "F6SCL")
# operand codes: coefficient (a, b) meaning a + b xi
CODES = [(1, 0), (-1, 0), (2, 0), (-2, 0), (0, 1), (0, -1), (3, 0), (0, 3)]
CODE = {c: i for i, c in enumerate(CODES)}
ABSENT = 0xFF
# cost in Fq multiplications (for the statistics printed into the header)
FQ_MULS = {"MULX": 3, "SQRX": 2, "MULF0": 2, "MULF1": 2, "INV": 2 + 2 + 314, "F6MUL": 18, "CYCSQR": 18, "F6M01": 15, "F6SCL": 9}


# ---------------------------------------------------------------------------------------------
# plain-integer Fq2 (for constants and for the simulator)
# ---------------------------------------------------------------------------------------------
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_mul(a, a)
        e >>= 1
    return r


def f2_inv(a):
    d = pow((a[0] * a[0] + a[1] * a[1]) % Q, Q - 2, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)


XI = (9, 1)
CONSTS = []          # list of Fq2 constants addressed by LDC
CONST_NAMES = []


def _const(name, v):
    CONSTS.append(v)
    CONST_NAMES.append(name)
    return len(CONSTS) - 1


C_ZERO = _const("zero", (0, 0))
C_ONE = _const("one", (1, 0))
C_FROB = [[_const("gamma_%d_%d" % (k, i), f2_pow(XI, i * (Q**k - 1) // 6)) for i in range(6)] for k in (1, 2, 3)]
C_TWX = _const("tw_x", f2_pow(XI, (Q - 1) // 3))
C_TWY = _const("tw_y", f2_pow(XI, (Q - 1) // 2))


# ---------------------------------------------------------------------------------------------
# tracer.  A V is a lazy linear combination  sum_k coef_k * value_k  (coef = a + b xi, small
# integers) of materialised values (SSA ids).  Additions, negations, doublings and xi-multiples
# cost nothing until the expression meets a multiplication, where up to two terms fold into the
# operand fetch and anything larger is first materialised by a LIN instruction.
# ---------------------------------------------------------------------------------------------
class Ins:
    __slots__ = ("op", "dsts", "opnds", "extra", "imm")

    def __init__(self, op, dsts, opnds=(), imm=0, extra=()):
        self.op, self.dsts, self.imm = op, list(dsts), imm
        self.opnds = list(opnds) + [None] * (6 - len(opnds))   # A1 A2 B1 B2 F G: (id, code) or None
        self.extra = list(extra)                                # further source ids of a macro operation

    @property
    def dst(self):
        return self.dsts[0]

    def srcs(self):
        return [o[0] for o in self.opnds if o is not None] + self.extra


class Trace:
    def __init__(self):
        self.ins = []
        self.nvals = 0
        self.inputs = []    # value ids preloaded by the kernel prologue, in slot order
        self.outputs = []
        self.memo = {}      # frozenset(terms) -> materialised id
        self.const_id = {}  # const index -> value id

    def new(self):
        self.nvals += 1
        return self.nvals - 1

    def emit(self, op, opnds=(), imm=0):
        d = self.new()
        self.ins.append(Ins(op, [d], opnds, imm))
        return d

    def emit_macro(self, op, srcs, ndst):
        """srcs: materialised single-term values (V); returns ndst new values"""
        ids = []
        for x in srcs:
            x = x.mat()
            if x.is_zero():
                x = zero_value(self)
            (vid, c), = x.terms.items()
            assert c == (1, 0)
            ids.append(vid)
        dsts = [self.new() for _ in range(ndst)]
        plain = CODE[(1, 0)]
        self.ins.append(Ins(op, dsts, [(v, plain) for v in ids[:6]], extra=ids[6:]))
        return [V.of(self, d) for d in dsts]


def _decompose(c):
    """coefficient (a, b) -> list of operand codes summing to it"""
    a, b = c
    out = []
    while a >= 3:
        out.append(CODE[(3, 0)]); a -= 3
    while a <= -2:
        out.append(CODE[(-2, 0)]); a += 2
    if a == 2:
        out.append(CODE[(2, 0)])
    elif a == 1:
        out.append(CODE[(1, 0)])
    elif a == -1:
        out.append(CODE[(-1, 0)])
    while b >= 3:
        out.append(CODE[(0, 3)]); b -= 3
    while b > 0:
        out.append(CODE[(0, 1)]); b -= 1
    while b < 0:
        out.append(CODE[(0, -1)]); b += 1
    return out


class V:
    __slots__ = ("t", "terms")

    def __init__(self, t, terms):
        self.t = t
        self.terms = {k: c for k, c in terms.items() if c != (0, 0)}

    @staticmethod
    def of(t, vid):
        return V(t, {vid: (1, 0)})

    def is_zero(self):
        return not self.terms

    def code_terms(self):
        out = []
        for vid in sorted(self.terms):
            out += [(vid, c) for c in _decompose(self.terms[vid])]
        return out

    def _lin(self, o, sgn):
        r = dict(self.terms)
        for k, (a, b) in o.terms.items():
            x = r.get(k, (0, 0))
            r[k] = (x[0] + sgn * a, x[1] + sgn * b)
        return V(self.t, r)

    def __add__(self, o): return self._lin(o, 1)
    def __sub__(self, o): return self._lin(o, -1)
    def __neg__(self): return V(self.t, {k: (-a, -b) for k, (a, b) in self.terms.items()})
    def scale(self, n): return V(self.t, {k: (n * a, n * b) for k, (a, b) in self.terms.items()})
    def dbl(self): return self.scale(2)
    def tpl(self): return self.scale(3)

    def mat(self):
        """force materialisation: returns a single-term (+1) expression"""
        if self.is_zero():
            return self
        if len(self.terms) == 1 and next(iter(self.terms.values())) == (1, 0):
            return self
        key = frozenset(self.terms.items())
        vid = self.t.memo.get(key)
        if vid is None:
            ct = self.code_terms()
            while len(ct) > 6:
                part = self.t.emit("LIN", ct[:6])
                ct = [(part, CODE[(1, 0)])] + ct[6:]
            vid = self.t.emit("LIN", ct)
            self.t.memo[key] = vid
        return V.of(self.t, vid)

    def mulxi(self):
        x = self
        if len(x.terms) > 1 or any(b != 0 for _, b in x.terms.values()):
            x = x.mat()
        return V(self.t, {k: (0, a) for k, (a, b) in x.terms.items()})

    def fold(self):
        """operand pair for a multiplication: at most two code terms, else materialise"""
        ct = self.code_terms()
        if len(ct) > 2:
            ct = self.mat().code_terms()
        return ct + [None] * (2 - len(ct))

    def __mul__(self, o):
        if self.is_zero() or o.is_zero():
            return V(self.t, {})
        return V.of(self.t, self.t.emit("MULX", self.fold() + o.fold()))

    def sqr(self):
        if self.is_zero():
            return self
        return V.of(self.t, self.t.emit("SQRX", self.fold()))

    def mulfq(self, o, half):
        """self * (o.c0 or o.c1), o a materialised value"""
        if self.is_zero():
            return self
        (oid, oc), = o.terms.items()
        assert oc == (1, 0)
        return V.of(self.t, self.t.emit("MULF%d" % half, self.fold() + [(oid, CODE[(1, 0)])]))

    def conj(self):
        if self.is_zero():
            return self
        return V.of(self.t, self.t.emit("CONJ", self.fold()))

    def inv(self):
        return V.of(self.t, self.t.emit("INV", self.fold()))


def zero_value(t):
    """a materialised zero (the lazy zero has no terms and no slot)"""
    one = const(t, C_ONE)
    oid = next(iter(one.terms))
    return V.of(t, t.emit("LIN", [(oid, CODE[(1, 0)]), (oid, CODE[(-1, 0)])]))


def const(t, idx):
    if idx == C_ZERO:
        return V(t, {})
    if idx not in t.const_id:
        t.const_id[idx] = t.emit("LDC", imm=idx)
    return V.of(t, t.const_id[idx])


# ---------------------------------------------------------------------------------------------
# tower formulas over V.  Fq6 = triple over v (v^3 = xi); Fq12 = list g[0..5] in the w-basis
# (w^6 = xi), i.e. tower halves c0 = (g0, g2, g4), c1 = (g1, g3, g5).
# ---------------------------------------------------------------------------------------------
def f6_add(a, b): return tuple(x + y for x, y in zip(a, b))
def f6_sub(a, b): return tuple(x - y for x, y in zip(a, b))
def f6_neg(a): return tuple(-x for x in a)
def f6_mul_v(a): return (a[2].mulxi(), a[0], a[1])


USE_MACROS = True


def f6_mul(a, b):
    if USE_MACROS and not any(x.is_zero() for x in a + b):
        return tuple(a[0].t.emit_macro("F6MUL", list(a) + list(b), 3))
    v0, v1, v2 = a[0] * b[0], a[1] * b[1], a[2] * b[2]
    c0 = v0 + ((a[1] + a[2]) * (b[1] + b[2]) - v1 - v2).mulxi()
    c1 = ((a[0] + a[1]) * (b[0] + b[1]) - v0 - v1).mat() + v2.mulxi()
    c2 = ((a[0] + a[2]) * (b[0] + b[2]) - v0 - v2).mat() + v1
    return (c0, c1, c2)


def f6_sqr(a):  # Chung-Hasan SQR2
    s0 = a[0].sqr()
    s1 = (a[0] * a[1]).dbl()
    s2 = (a[0] - a[1] + a[2]).sqr()
    s3 = (a[1] * a[2]).dbl()
    s4 = a[2].sqr()
    return (s0 + s3.mulxi(), s1 + s4.mulxi(), s1 + s2 + s3 - s0 - s4)


def f6_mul_f2(a, k):
    if USE_MACROS and not any(x.is_zero() for x in a) and not k.is_zero():
        return tuple(a[0].t.emit_macro("F6SCL", list(a) + [k], 3))
    return (a[0] * k, a[1] * k, a[2] * k)


def f6_mul_01(a, b0, b1):  # a * (b0 + b1 v)
    if USE_MACROS and not any(x.is_zero() for x in a) and not b0.is_zero() and not b1.is_zero():
        return tuple(a[0].t.emit_macro("F6M01", list(a) + [b0, b1], 3))
    aa, bb = a[0] * b0, a[1] * b1
    c0 = ((a[1] + a[2]) * b1 - bb).mulxi() + aa
    c1 = ((a[0] + a[1]) * (b0 + b1) - aa - bb).mat()
    c2 = ((a[0] + a[2]) * b0 - aa).mat() + bb
    return (c0, c1, c2)


def f6_inv(a):
    t0 = a[0].sqr() - (a[1] * a[2]).mulxi()
    t1 = a[2].sqr().mulxi() - a[0] * a[1]
    t2 = a[1].sqr() - a[0] * a[2]
    d = (a[0] * t0 + (a[2] * t1 + a[1] * t2).mulxi()).inv()
    return (t0 * d, t1 * d, t2 * d)


def halves(g): return (g[0], g[2], g[4]), (g[1], g[3], g[5])
def from_halves(c0, c1): return [x.mat() for x in (c0[0], c1[0], c0[1], c1[1], c0[2], c1[2])]


def f12_mul(a, b):
    a0, a1 = halves(a)
    b0, b1 = halves(b)
    t0, t1 = f6_mul(a0, b0), f6_mul(a1, b1)
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a0, a1), f6_add(b0, b1)), t0), t1)
    c0 = f6_add(t0, f6_mul_v(t1))
    return from_halves(c0, c1)


def f12_sqr(a):
    a0, a1 = halves(a)
    t = f6_mul(a0, a1)
    c0 = f6_sub(f6_sub(f6_mul(f6_add(a0, a1), f6_add(a0, f6_mul_v(a1))), t), f6_mul_v(t))
    c1 = f6_add(t, t)
    return from_halves(c0, c1)


def f12_conj(a): return [a[0], -a[1], a[2], -a[3], a[4], -a[5]]   # lazy: signs fold into the operand fetch


def f12_inv(a):
    a0, a1 = halves(a)
    d = f6_inv(f6_sub(f6_sqr(a0), f6_mul_v(f6_sqr(a1))))
    return from_halves(f6_mul(a0, d), f6_neg(f6_mul(a1, d)))


def f12_frobenius(t, a, k):
    out = []
    for i in range(6):
        x = a[i].conj() if (k & 1) else a[i]
        if i == 0:
            out.append(x)
        elif k == 2:  # gamma_{2,i} lies in Fq
            out.append(x.mulfq(const(t, C_FROB[1][i]), 0))
        else:
            out.append(x * const(t, C_FROB[k - 1][i]))
    return out


def fq4_sqr(a, b):
    a2, b2 = a.sqr(), b.sqr()
    return (a2 + b2.mulxi()).mat(), ((a + b).sqr() - a2 - b2).mat()


def cyclotomic_sqr(g):  # Granger-Scott; same arrangement as tower.cuh (checked there)
    if USE_MACROS:
        return g[0].t.emit_macro("CYCSQR", g, 6)
    A0, A1 = fq4_sqr(g[0], g[3])
    B0, B1 = fq4_sqr(g[1], g[4])
    C0, C1 = fq4_sqr(g[2], g[5])
    h = [None] * 6
    h[0] = (A0 - g[0]).dbl() + A0
    h[3] = (A1 + g[3]).dbl() + A1
    sC0, sC1 = C1.mulxi(), C0
    h[1] = (sC0 + g[1]).dbl() + sC0
    h[4] = (sC1 - g[4]).dbl() + sC1
    h[2] = (B0 - g[2]).dbl() + B0
    h[5] = (B1 + g[5]).dbl() + B1
    return [x.mat() for x in h]


def mul_by_line(f, l0, l1, l3):
    """f * (l0 + l1 w + l3 w^3): the line is c0 = (l0, 0, 0), c1 = (l1, l3, 0) in the tower;
    13 Fq2 products."""
    f0, f1 = halves(f)
    a = f6_mul_f2(f0, l0)
    b = f6_mul_01(f1, l1, l3)
    e = f6_mul_01(f6_add(f0, f1), l0 + l1, l3)
    c1 = f6_sub(f6_sub(e, a), b)
    c0 = f6_add(a, f6_mul_v(b))
    return from_halves(c0, c1)


# ---------------------------------------------------------------------------------------------
# Miller loop (Jacobian accumulator on the twist; lines scaled by subfield elements only) and the
# arkworks final exponentiation (easy part, Fuentes-Castaneda hard part in the y0..y16 arrangement)
# ---------------------------------------------------------------------------------------------
def line_dbl(T):
    X, Y, Zc = T
    a, b = X.sqr(), Y.sqr()
    c = b.sqr()
    d = ((X + b).sqr() - a - c).mat().dbl()
    e = a.tpl()
    zz = Zc.sqr()
    z3 = (Y * Zc).dbl()
    l3 = e * X - b.dbl()
    l1 = -(e * zz)
    l0 = z3 * zz
    x3 = e.sqr() - d.dbl()
    x3 = x3.mat()
    y3 = e * (d - x3) - c.dbl().dbl().dbl()
    return (l0, l1, l3.mat()), (x3, y3.mat(), z3.mat())


def line_add(T, x2, y2):
    X, Y, Zc = T
    zz = Zc.sqr()
    h = (x2 * zz - X).mat()
    r = (y2 * (Zc * zz) - Y).mat()
    z3 = Zc * h
    l3 = r * x2 - z3 * y2
    h2 = h.sqr()
    h3 = h * h2
    v = X * h2
    x3 = (r.sqr() - h3 - v.dbl()).mat()
    y3 = r * (v - x3) - Y * h3
    return (z3, -r, l3.mat()), (x3, y3.mat(), z3)


def line_add_first(x1, y1, x2, y2):
    """chord through two affine points (Z = 1)"""
    h = x2 - x1
    r = y2 - y1
    l3 = r * x2 - h * y2
    h2 = h.sqr()
    h3 = h * h2
    v = x1 * h2
    x3 = r.sqr() - h3 - v.dbl()
    y3 = r * (v - x3) - y1 * h3
    return (h, -r, l3), (x3, y3, h)


def apply_line(f, line, P):
    l0, l1, l3 = line
    return mul_by_line(f, l0.mulfq(P, 1), l1.mulfq(P, 0), l3)


def miller(t, P, Qx, Qy):
    nQy = -Qy
    # first iteration (i = 63): f = 1, T = Q affine
    X, Y = Qx, Qy
    a, b = X.sqr(), Y.sqr()
    c = b.sqr()
    d = ((X + b).sqr() - a - c).mat().dbl()
    e = a.tpl()
    z3 = Y.dbl()
    l3 = e * X - b.dbl()
    x3 = (e.sqr() - d.dbl()).mat()
    y3 = (e * (d - x3) - c.dbl().dbl().dbl()).mat()
    T = (x3, y3, z3.mat())
    l3 = l3.mat()
    zero = const(t, C_ZERO)
    f = [z3.mulfq(P, 1), (-e).mulfq(P, 0), zero, l3, zero, zero]
    for i in range(63, -1, -1):
        if i != 63:
            f = f12_sqr(f)
            line, T = line_dbl(T)
            f = apply_line(f, line, P)
        dgt = ATE[i]
        if dgt:
            line, T = line_add(T, Qx, Qy if dgt > 0 else nQy)
            f = apply_line(f, line, P)
    twx, twy = const(t, C_TWX), const(t, C_TWY)
    q1x, q1y = Qx.conj() * twx, Qy.conj() * twy
    q2x, q2y = q1x.conj() * twx, -(q1y.conj() * twy)
    line, T = line_add(T, q1x, q1y)
    f = apply_line(f, line, P)
    line, T = line_add(T, q2x, q2y)
    f = apply_line(f, line, P)
    return f


def wnaf(k, w):
    """width-w non-adjacent form, LSB first: non-zero digits are odd, |d| < 2^(w-1)"""
    out = []
    while k:
        if k & 1:
            d = k % (1 << w)
            if d >= 1 << (w - 1):
                d -= 1 << w
            k -= d
        else:
            d = 0
        out.append(d)
        k >>= 1
    return out


# Exponentiation by z in the cyclotomic subgroup (inverse = conjugate, free): width-4 wNAF of z has 14 non-zero
# digits in {+-1, +-3, +-5, +-7} against 24 for the plain NAF, so 13 products + 4 for the table (a^2, a^3, a^5, a^7)
# replace 23; the table lives in slots / the global scratch like every other value.
Z_WNAF_W = 4
Z_WNAF = wnaf(Z, Z_WNAF_W)
assert sum(d << i for i, d in enumerate(Z_WNAF)) == Z and Z_WNAF[-1] == 1


def cyclotomic_exp_z(a):
    tab = {1: a}
    top = max(abs(d) for d in Z_WNAF)
    if top > 1:
        a2 = cyclotomic_sqr(a)
        for d in range(3, top + 1, 2):
            tab[d] = f12_mul(tab[d - 2], a2)
    r = a
    for i in range(len(Z_WNAF) - 2, -1, -1):
        r = cyclotomic_sqr(r)
        d = Z_WNAF[i]
        if d > 0:
            r = f12_mul(r, tab[d])
        elif d < 0:
            r = f12_mul(r, f12_conj(tab[-d]))
    return r


def final_exp(t, f):
    r = f12_mul(f12_conj(f), f12_inv(f))
    r = f12_mul(f12_frobenius(t, r, 2), r)
    y0 = f12_conj(cyclotomic_exp_z(r))
    y1 = cyclotomic_sqr(y0)
    y2 = cyclotomic_sqr(y1)
    y3 = f12_mul(y2, y1)
    y4 = f12_conj(cyclotomic_exp_z(y3))
    y5 = cyclotomic_sqr(y4)
    y6 = cyclotomic_exp_z(y5)
    y3 = f12_conj(y3)
    y7 = f12_mul(y6, y4)
    y8 = f12_mul(y7, y3)
    y9 = f12_mul(y8, y1)
    y10 = f12_mul(y8, y4)
    y11 = f12_mul(y10, r)
    y12 = f12_frobenius(t, y9, 1)
    y13 = f12_mul(y12, y11)
    y8 = f12_frobenius(t, y8, 2)
    y14 = f12_mul(y8, y13)
    y15 = f12_frobenius(t, f12_mul(f12_conj(r), y9), 3)
    return f12_mul(y15, y14)


def trace_pairing(what="pairing"):
    """inputs (slot order): P = (xP, yP) packed as one Fq2, Q.x, Q.y;  outputs: the 6 Fq2 of GT in
    ark-serialize order (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2)."""
    t = Trace()
    ids = [t.new() for _ in range(3)]
    t.inputs = ids
    P, Qx, Qy = (V.of(t, i) for i in ids)
    f = miller(t, P, Qx, Qy)
    if what == "pairing":
        f = final_exp(t, f)
    c0, c1 = halves(f)
    t.outputs = []
    for x in c0 + c1:
        x = x.mat()
        if x.is_zero():
            x = zero_value(t)
        t.outputs.append(next(iter(x.terms)))
    return t


# ---------------------------------------------------------------------------------------------
# passes: fuse "product - f - g" into the product, dead code, slot allocation with
# furthest-next-use eviction (spills go to the global scratch; constants are rematerialised)
# ---------------------------------------------------------------------------------------------
def fuse_posts(t):
    use_count = {}
    for i in t.ins:
        for s_ in i.srcs():
            use_count[s_] = use_count.get(s_, 0) + 1
    for o in t.outputs:
        use_count[o] = use_count.get(o, 0) + 1
    defs = {i.dst: i for i in t.ins if len(i.dsts) == 1}
    dead = set()
    for L in t.ins:
        if L.op != "LIN":
            continue
        terms = [o for o in L.opnds if o is not None]
        if not 1 <= len(terms) <= 3:
            continue
        for k, (vid, code) in enumerate(terms):
            M = defs.get(vid)
            if (code == CODE[(1, 0)] and M is not None and M.op in ("MULX", "SQRX", "MULF0", "MULF1")
                    and M.opnds[4] is None and M.opnds[5] is None and use_count.get(vid) == 1 and id(M) not in dead):
                rest = terms[:k] + terms[k + 1:]
                L.op, L.imm = M.op, M.imm
                L.opnds = M.opnds[:4] + rest + [None] * (2 - len(rest))
                dead.add(id(M))
                break
    t.ins = [i for i in t.ins if id(i) not in dead]


def eliminate_dead(t):
    live = set(t.outputs)
    keep = []
    for ins in reversed(t.ins):
        if any(d in live for d in ins.dsts):
            keep.append(ins)
            live.update(ins.srcs())
    t.ins = keep[::-1]


def encode(op, dst=0, ob=(ABSENT,) * 6, xb=()):
    """-> (w0, w1)"""
    w0 = OP[op] | (dst << 8)
    for k, b in enumerate(ob):
        w0 |= b << (16 + 8 * k)
    w1 = 0
    xb = list(xb) + [ABSENT] * (8 - len(xb))
    for k, b in enumerate(xb):
        w1 |= b << (8 * k)
    return (w0, w1)


def allocate(t, nslots):
    """Returns (words, n_global_slots, out_slots, stats): two 64-bit words per instruction.
    Slots 0..len(inputs)-1 hold the inputs.  Every lane reads all operands of an instruction before any lane
    stores a result (the interpreter synchronises in between), so destinations may reuse the slots of sources
    that die at the instruction."""
    assert nslots <= 31
    real = [i for i in t.ins if i.op != "LDC"]
    const_of = {i.dst: i.imm for i in t.ins if i.op == "LDC"}
    INF = 1 << 60
    uses = {}
    for pos, i in enumerate(real):
        for s_ in i.srcs():
            uses.setdefault(s_, []).append(pos)
    for o in t.outputs:
        uses.setdefault(o, []).append(len(real))
    ptr = {v: 0 for v in uses}

    def next_use(v, pos):
        u = uses.get(v)
        if not u:
            return INF
        k = ptr[v]
        while k < len(u) and u[k] < pos:
            k += 1
        ptr[v] = k
        return u[k] if k < len(u) else INF

    slot_of, val_in = {}, [None] * nslots
    gslot_of, gfree, gcount = {}, [], 0
    for k, v in enumerate(t.inputs):
        slot_of[v] = k
        val_in[k] = v
    out = []
    stats = {"spill_st": 0, "spill_ld": 0}

    def release(v):
        s_ = slot_of.pop(v)
        val_in[s_] = None
        if v in gslot_of:
            gfree.append(gslot_of.pop(v))

    def get_slot(pos, pinned):
        nonlocal gcount
        for s_ in range(nslots):
            if val_in[s_] is None:
                return s_
        best, bu = None, -1
        for s_ in range(nslots):
            v = val_in[s_]
            if v in pinned:
                continue
            u = next_use(v, pos)
            if v in const_of or v in gslot_of:
                u += 0.5   # no store needed: prefer on ties
            if u > bu:
                best, bu = s_, u
        assert best is not None, "too few slots"
        v = val_in[best]
        if v not in const_of and v not in gslot_of:
            g = gfree.pop() if gfree else gcount
            if g == gcount:
                gcount += 1
            gslot_of[v] = g
            out.append(encode("STG", 0, (g & 255, g >> 8, best, ABSENT, ABSENT, ABSENT)))
            stats["spill_st"] += 1
        del slot_of[v]
        val_in[best] = None
        return best

    def ensure(v, pos, pinned):
        if v in slot_of:
            return slot_of[v]
        s_ = get_slot(pos, pinned)
        if v in const_of:
            c = const_of[v]
            out.append(encode("LDC", s_, (c & 255, c >> 8, ABSENT, ABSENT, ABSENT, ABSENT)))
        else:
            g = gslot_of[v]
            out.append(encode("LDG", s_, (g & 255, g >> 8, ABSENT, ABSENT, ABSENT, ABSENT)))
            stats["spill_ld"] += 1
        slot_of[v] = s_
        val_in[s_] = v
        return s_

    for pos, i in enumerate(real):
        pinned = set(i.srcs())
        for v in sorted(pinned):
            ensure(v, pos, pinned)
        ob = [ABSENT if o is None else (slot_of[o[0]] | (o[1] << 5)) for o in i.opnds]
        xsrc = [slot_of[v] for v in i.extra]
        for v in sorted(pinned):
            if next_use(v, pos + 1) == INF:
                release(v)
        pin2 = set(pinned)
        ds = []
        for d in i.dsts:
            sl = get_slot(pos + 1, pin2)
            slot_of[d] = sl
            val_in[sl] = d
            pin2.add(d)
            ds.append(sl)
        out.append(encode(i.op, ds[0], ob, xsrc + ds[1:]))
        for d in i.dsts:   # results nobody reads (a macro operation always produces all of its outputs)
            if next_use(d, pos + 1) == INF:
                release(d)
    pinned = set(t.outputs)
    out_slots = [ensure(v, len(real), pinned) for v in t.outputs]
    out.append(encode("END"))
    flat = [w for pair in out for w in pair]
    stats["n_ins"] = len(out)
    stats["fq_muls"] = sum(FQ_MULS.get(OPS[w0 & 255], 0) for w0, _ in out)
    hist = {}
    for w0, _ in out:
        hist[OPS[w0 & 255]] = hist.get(OPS[w0 & 255], 0) + 1
    stats["hist"] = hist
    return flat, gcount, out_slots, stats


# ---------------------------------------------------------------------------------------------
# simulator (plain integers, canonical values) - used by the tests.  The macro operations are
# evaluated as plain polynomial arithmetic over Fq2, independently of the formulas above.
# ---------------------------------------------------------------------------------------------
def _apply(code, x):
    a, b = CODES[code]
    r = (a * x[0] % Q, a * x[1] % Q)
    if b:
        y = f2_mul(x, XI)
        r = ((r[0] + b * y[0]) % Q, (r[1] + b * y[1]) % Q)
    return r


def _f2_add(x, y): return ((x[0] + y[0]) % Q, (x[1] + y[1]) % Q)


def _polymul_mod(a, b, n):
    """(sum a_i X^i)(sum b_j X^j) mod X^n - xi over Fq2"""
    r = [(0, 0)] * n
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            pr = f2_mul(x, y)
            k = i + j
            while k >= n:
                pr = f2_mul(pr, XI)
                k -= n
            r[k] = _f2_add(r[k], pr)
    return r


def simulate(words, nslots, out_slots, inputs):
    """inputs: list of Fq2 (tuples of ints) for slots 0..; returns the Fq2 values of out_slots."""
    S = [(0, 0)] * 32
    G = {}
    for k, v in enumerate(inputs):
        S[k] = v
    for base in range(0, len(words), 2):
        w, w1 = words[base], words[base + 1]
        op, d = OPS[w & 255], (w >> 8) & 255
        ob = [(w >> (16 + 8 * k)) & 255 for k in range(6)]
        xb = [(w1 >> (8 * k)) & 255 for k in range(8)]
        if op == "END":
            break
        if op == "NOP":
            continue
        if op == "LDC":
            S[d] = CONSTS[ob[0] | (ob[1] << 8)]
            continue
        if op == "LDG":
            S[d] = G[ob[0] | (ob[1] << 8)]
            continue
        if op == "STG":
            G[ob[0] | (ob[1] << 8)] = S[ob[2]]
            continue
        assert d < nslots and all(b == ABSENT or (b & 31) < nslots for b in ob)
        if op in MACROS:
            assert all(b == ABSENT or b >> 5 == 0 for b in ob)
            src = [None if b == ABSENT else S[b] for b in ob]
            if op == "F6MUL":
                res = _polymul_mod(src[:3], src[3:], 3)
                dsts = [d, xb[0], xb[1]]
            elif op == "CYCSQR":
                res = _polymul_mod(src, src, 6)
                dsts = [d] + xb[:5]
            elif op == "F6M01":
                res = _polymul_mod(src[:3], [src[3], src[4], (0, 0)], 3)
                dsts = [d, xb[0], xb[1]]
            else:
                res = _polymul_mod(src[:3], [src[3]], 3)
                dsts = [d, xb[0], xb[1]]
            assert len(set(dsts)) == len(dsts) and all(x < nslots for x in dsts)
            for k, x in zip(dsts, res):
                S[k] = x
            continue
        val = [(0, 0) if b == ABSENT else _apply(b >> 5, S[b & 31]) for b in ob]
        A, B = _f2_add(val[0], val[1]), _f2_add(val[2], val[3])
        if op == "MULX": r = f2_mul(A, B)
        elif op == "SQRX": r = f2_mul(A, A)
        elif op == "MULF0": r = (A[0] * S[ob[2] & 31][0] % Q, A[1] * S[ob[2] & 31][0] % Q)
        elif op == "MULF1": r = (A[0] * S[ob[2] & 31][1] % Q, A[1] * S[ob[2] & 31][1] % Q)
        elif op == "LIN": r = _f2_add(A, B)
        elif op == "CONJ": r = (A[0], -A[1] % Q)
        elif op == "INV": r = f2_inv(A) if A != (0, 0) else (0, 0)
        else:
            raise ValueError(op)
        if op in ("MULX", "SQRX", "MULF0", "MULF1", "LIN"):
            r = _f2_add(_f2_add(r, val[4]), val[5])
        S[d] = r
    return [S[s_] for s_ in out_slots]


def build(what, nslots):
    t = trace_pairing(what)
    fuse_posts(t)
    eliminate_dead(t)
    return allocate(t, nslots)


# ---------------------------------------------------------------------------------------------
# emit
# ---------------------------------------------------------------------------------------------
def limbs(x):
    return ", ".join("0x%08xu" % ((x >> (32 * i)) & 0xFFFFFFFF) for i in range(8))


# (what, slots): the kernel picks by slot count
VARIANTS = [("pairing", 14), ("pairing", 15), ("pairing", 16), ("pairing", 18), ("pairing", 28)]


def variant_name(what, ns):
    return "%s_S%d" % (what.upper(), ns)


def main():
    out = ["// GENERATED by tools/gen_pairing_prog.py - do not edit.",
           "// Straight-line programs for the pairing VM (pairing_vm.cuh): optimal-ate Miller loop + arkworks final",
           "// exponentiation, slot-allocated for S shared-memory slots per pairing.  Two 64-bit words per instruction:",
           "// w0 = op | dst<<8 | A1<<16 | A2<<24 | B1<<32 | B2<<40 | F<<48 | G<<56 (operand = slot | code<<5, 0xFF absent),",
           "// w1 = eight further slot numbers of the macro operations.",
           "#pragma once", "#include <stdint.h>", "namespace kb { namespace vmprog {",
           "enum Op : uint32_t { " + ", ".join("OP_%s = %d" % (n, i) for i, n in enumerate(OPS)) + " };",
           "// operand codes: " + ", ".join("%d: %+d%+dxi" % (i, a, b) for i, (a, b) in enumerate(CODES)),
           "static const int NUM_CONSTS = %d;" % len(CONSTS),
           "// Fq2 constants (Montgomery form): " + ", ".join(CONST_NAMES),
           "static const uint32_t CONSTS[%d * 16] = {" % len(CONSTS)]
    for c in CONSTS:
        out.append("    " + limbs(c[0] * MONT % Q) + ", " + limbs(c[1] * MONT % Q) + ",")
    out.append("};")
    out.append("struct Program { int slots, gslots, len; uint8_t out[6]; const uint64_t* words; };  // len in 64-bit words")
    names = []
    for what, ns in VARIANTS:
        words, gcount, out_slots, st = build(what, ns)
        name = variant_name(what, ns)
        names.append((name, ns, gcount, len(words), out_slots))
        out.append("// %s: %d instructions, %d Fq multiplications, %d spill stores / %d spill loads, %d global slots; %s"
                   % (name, st["n_ins"], st["fq_muls"], st["spill_st"], st["spill_ld"], gcount,
                      " ".join("%s=%d" % kv for kv in sorted(st["hist"].items()))))
        out.append("static const uint64_t %s_WORDS[%d] = {" % (name, len(words)))
        for i in range(0, len(words), 6):
            out.append("    " + ", ".join("0x%016xull" % w for w in words[i:i + 6]) + ",")
        out.append("};")
        print(name, st, "gslots", gcount, file=sys.stderr)
    out.append("static const Program PROGRAMS[%d] = {" % len(names))
    for name, ns, gcount, ln, outs in names:
        out.append("    {%d, %d, %d, {%s}, %s_WORDS}," % (ns, gcount, ln, ", ".join(str(s_) for s_ in outs), name))
    out.append("};")
    out.append("static const int NUM_PROGRAMS = %d;" % len(names))
    out.append("}}  // namespace kb::vmprog")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "keaki_b200", "csrc", "pairing_prog_gen.cuh")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
