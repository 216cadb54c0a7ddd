#!/usr/bin/env python3
"""Generate keaki_b200/csrc/pairing_warp_gen.cuh: lane-parallel schedules of the pairing for the WARP-COOPERATIVE kernel
(keaki_b200/csrc/pairing_warp.cuh) - one warp per pairing, the 32 lanes executing independent Fq2 operations of the same
pairing side by side.  It is the low-latency form (one pairing in about a millisecond instead of nine for a lone thread of
the throughput kernel pairing_st): used for small batches, for the per-commitment pairing of kb_encrypt_batch and for the
window-base chain of its GT tables.

The tower formulas are those of tools/gen_pairing_prog.py (Miller loop, final exponentiation, width-4 wNAF ...), traced
here with an EAGER value class: every operation becomes a node of kind
    MUL  d = (a1 + a2) * b          Fq2 product; the a-side sum is not reduced (a2 = the zero slot when absent)
    LIN  d = s1 x1 + s2 x2 + s3 x3 + s4 x4,  s in {+1, -1}   (absent terms point at the zero slot)
    XI   d = (9 + u) x              CONJ  d = conj(x)              INV  d = 1 / x
and the DAG is list-scheduled into STEPS of up to 32 nodes of one kind (critical-path priority); all lanes of a step read
their operands, synchronise, then write, so a slot freed in a step can be reused by that step's results.  Values live in
the warp's shared-memory slot file (allocated here, first-fit).

Standalone: plain Python integers, no import of oracle/.  tests/test_pairing_warp.py simulates the emitted schedules with
`simulate()` against the oracle; tests/hostemu runs them through the very interpreter the GPU runs.

Run:  python tools/gen_pairing_warp.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_pairing_prog as gp  # noqa: E402  (formulas, constants, integer Fq2 helpers)

Q = gp.Q
KINDS = ["NOP", "MUL", "LIN", "CONJ", "INV"]
KIND = {k: i for i, k in enumerate(KINDS)}
LANES = 32
ZERO_SLOT = 0          # slot 0 always holds 0
MAX_SLOTS = 440        # 28 KB of shared memory per warp


# ---------------------------------------------------------------------------------------------
# tracer: values are LAZY linear combinations sum (a_i + b_i xi) x_i of materialised nodes (small integer a_i, b_i);
# a combination is materialised - one LIN node - only where a product, a conjugation, an inversion or an output needs
# the value itself, so that exactly one LIN level separates two product levels.
# ---------------------------------------------------------------------------------------------
MAX_TERMS = 14         # terms c * x one LIN descriptor holds
MAX_COEF = 32
LIN_BOUND = 120        # |integer value of a LIN| < LIN_BOUND * q
GT_PROD_N = 48         # operands of the GT product schedule (32 window entries of A or A', 16 of gT; unused ones = 1)


class Node:
    __slots__ = ("kind", "srcs", "u", "w", "phase")

    def __init__(self, kind, srcs=(), u=(), w=()):
        self.kind, self.srcs, self.u, self.w, self.phase = kind, list(srcs), list(u), list(w), 0

    def deps(self):
        return [s for s in self.srcs if s >= 0] + [v for v, _ in self.u] + [v for v, _ in self.w]


class WTrace:
    def __init__(self):
        self.nodes = []          # Node per value id (None for inputs / constants)
        self.inputs = []         # value ids preloaded by the kernel, in slot order
        self.consts = {}         # const index -> value id
        self.memo = {}
        self.outputs = []
        self.split = {}
        self.phase = 0           # nodes of a later phase are scheduled after all nodes of the earlier ones (bounds live values)

    def new_input(self):
        self.nodes.append(None)
        self.inputs.append(len(self.nodes) - 1)
        return len(self.nodes) - 1

    def emit(self, node, key):
        if key is not None and key in self.memo:
            return self.memo[key]
        node.phase = self.phase
        self.nodes.append(node)
        vid = len(self.nodes) - 1
        if key is not None:
            self.memo[key] = vid
        return vid

    def emit_mul(self, a1, a2, b1, b2):
        return self.emit(Node("MUL", [a1, a2, b1, b2]), ("MUL", a1, a2, b1, b2))

    def emit_lin(self, u, w):
        u, w = sorted(u), sorted(w)
        return self.emit(Node("LIN", u=u, w=w), ("LIN", tuple(u), tuple(w)))


def _units(terms):
    """term lists (U, W) of a coefficient dict: (vid, coefficient)"""
    u = [(v, a) for v, (a, b) in sorted(terms.items()) if a]
    w = [(v, b) for v, (a, b) in sorted(terms.items()) if b]
    return u, w


def _nunits(terms):
    return sum((a != 0) + (b != 0) for a, b in terms.values())


class EV:
    """same interface as gen_pairing_prog.V, for the formulas of that module"""
    __slots__ = ("t", "terms")

    def __init__(self, t, terms):
        self.t = t
        self.terms = {v: c for v, c in terms.items() if c != (0, 0)}

    @staticmethod
    def of(t, vid):
        return EV(t, {vid: (1, 0)})

    def is_zero(self):
        return not self.terms

    def single(self):
        """(vid, a) when the value is a * x with a plain integer a"""
        if len(self.terms) == 1:
            (v, (a, b)), = self.terms.items()
            if b == 0:
                return v, a
        return None

    def mat(self):
        return self

    def force(self):
        """the value as ONE materialised node with coefficient 1"""
        s = self.single()
        if s is not None and s[1] == 1:
            return self
        terms = dict(self.terms)
        while any(abs(a) > MAX_COEF or abs(b) > MAX_COEF for a, b in terms.values()):     # big multiples: through 32 x
            v, (a, b) = next((v, c) for v, c in terms.items() if abs(c[0]) > MAX_COEF or abs(c[1]) > MAX_COEF)
            ha, hb = int(a / 32), int(b / 32)
            d = self.t.emit_lin([(v, 32)], [])
            terms[v] = (a - 32 * ha, b - 32 * hb)
            a0, b0 = terms.get(d, (0, 0))
            terms[d] = (a0 + ha, b0 + hb)
            terms = {k: c for k, c in terms.items() if c != (0, 0)}
        def weight(c):
            return abs(c[0]) + 10 * abs(c[1])
        while _nunits(terms) > MAX_TERMS or sum(weight(c) for c in terms.values()) > LIN_BOUND:
            # split: materialise the heaviest part that fits one descriptor, keep going with it as one term
            part, n, wsum = {}, 0, 0
            for v, c in sorted(terms.items(), key=lambda kv: -weight(kv[1])):
                k = (c[0] != 0) + (c[1] != 0)
                if n + k <= MAX_TERMS and wsum + weight(c) <= LIN_BOUND:
                    part[v] = c
                    n += k
                    wsum += weight(c)
            assert wsum >= 2
            for v in part:
                del terms[v]
            u, w = _units(part)
            pv = self.t.emit_lin(u, w)
            a, b = terms.get(pv, (0, 0))
            terms[pv] = (a + 1, b)
        u, w = _units(terms)
        # the value exists from now on: later uses of this object name the node instead of re-expanding the combination
        self.terms = {self.t.emit_lin(u, w): (1, 0)}
        return self

    def _comb(self, o, sgn):
        terms = dict(self.terms)
        for v, (a, b) in o.terms.items():
            a0, b0 = terms.get(v, (0, 0))
            terms[v] = (a0 + sgn * a, b0 + sgn * b)
        r = EV(self.t, terms)
        if _nunits(r.terms) > MAX_TERMS:
            # materialise the heavier side and retry
            if _nunits(self.terms) >= _nunits(o.terms) and _nunits(self.terms) > 1:
                return self.force()._comb(o, sgn)
            return self._comb(o.force(), sgn)
        return r

    def __add__(self, o): return self._comb(o, 1)
    def __sub__(self, o): return self._comb(o, -1)
    def __neg__(self): return EV(self.t, {v: (-a, -b) for v, (a, b) in self.terms.items()})

    def scale(self, n):
        r = EV(self.t, {v: (a * n, b * n) for v, (a, b) in self.terms.items()})
        if _nunits(r.terms) > MAX_TERMS:
            return self.force().scale(n)
        return r

    def dbl(self): return self.scale(2)
    def tpl(self): return self.scale(3)

    def mulxi(self):
        if any(b for _, b in self.terms.values()):
            return self.force().mulxi()
        return EV(self.t, {v: (0, a) for v, (a, _) in self.terms.items()})

    def _operand(self):
        """(coefficient (a, b), v1, v2): the value as coefficient * (v1 + v2) with materialised v1, v2 (v2 = ZERO_VID)"""
        if len(self.terms) == 1:
            (v, c), = self.terms.items()
            if max(abs(c[0]), abs(c[1])) <= 4:      # small multiples ride on the product's coefficient
                return c, v, ZERO_VID
        if FOLD_SUMS and len(self.terms) == 2 and all(c == (1, 0) for c in self.terms.values()):
            v1, v2 = sorted(self.terms)
            return (1, 0), v1, v2
        return (1, 0), next(iter(self.force().terms)), ZERO_VID

    def __mul__(self, o):
        if self.is_zero() or o.is_zero():
            return EV(self.t, {})
        (a0, b0), x1, x2 = self._operand()
        (a1, b1), y1, y2 = o._operand()
        if b0 and b1:   # xi^2: materialise one side
            return self.force() * o
        # a-side (summed without reduction) and b-side (summed mod q) are interchangeable: canonical order for the memo
        (x1, x2), (y1, y2) = sorted([(x1, x2), (y1, y2)])
        return EV(self.t, {self.t.emit_mul(x1, x2, y1, y2): (a0 * a1, a0 * b1 + a1 * b0)})

    def sqr(self):
        return self * self

    def mulfq(self, o, half):
        """self * (o.c0 or o.c1): o is the input P, whose coordinates the kernel prologue also stores as (xP, 0) and
        (yP, 0), or a constant of Fq, (c, 0) itself"""
        if self.is_zero():
            return self
        ov = next(iter(o.terms))
        if (ov, half) in self.t.split:
            return self * EV.of(self.t, self.t.split[(ov, half)])
        assert half == 0 and ov in self.t.consts.values()
        return self * o

    def conj(self):
        if self.is_zero():
            return self
        x = next(iter(self.force().terms))
        return EV.of(self.t, self.t.emit(Node("CONJ", [x]), ("CONJ", x)))

    def inv(self):
        x = next(iter(self.force().terms))
        return EV.of(self.t, self.t.emit(Node("INV", [x]), None))


FOLD_SUMS = True
SHALLOW = True        # depth-1 Fq12 routines and the shallow point chain (below)
ZERO_VID = -1   # "the zero slot" inside MUL sources


def _const(t, idx):
    if idx == gp.C_ZERO:
        return EV(t, {})
    if idx not in t.consts:
        t.nodes.append(None)
        t.consts[idx] = len(t.nodes) - 1
    return EV.of(t, t.consts[idx])


def _zero_value(t):
    return EV(t, {})


# ---------------------------------------------------------------------------------------------
# Miller loop with a SHALLOW point chain.  The G2 accumulator T is independent of f, and with the Jacobian formulas of
# gen_pairing_prog.py its dependent chain (3 product levels per doubling, 5 per addition: 372 levels) is what bounds the
# Miller loop of a lane-parallel schedule, not the 156 levels of f.  Homogeneous projective coordinates on the twist
# y^2 = x^3 + b' (b' = 3 / xi) as in arkworks' own BN `doubling_step` / `addition_step` (ark-ec 0.4.2 models/bn/g2.rs),
# rearranged for depth: T = (X, Y, Z, Zb = b' Z); doubling in 2 levels (11 products), mixed addition in 3 (20 products).
# Line functions differ from the Jacobian ones by factors in Fq2, which the final exponentiation kills: GT is the same.
# ---------------------------------------------------------------------------------------------
C_TWIST_B = len(gp.CONSTS)                                   # index in WP_CONSTS
WP_CONSTS = list(gp.CONSTS) + [gp.f2_mul((3, 0), gp.f2_inv(gp.XI))]
WP_CONST_NAMES = list(gp.CONST_NAMES) + ["twist_b"]


def proj_dbl(T):
    """2 T and the tangent line (l0, l1, l3): l0 yP + l1 xP w + l3 w^3"""
    X, Y, Z, Zb = T
    xy, b, c, yz, yzb, j = X * Y, Y.sqr(), Z * Zb, Y * Z, Y * Zb, X.sqr()     # c = b' Z^2
    e = c.tpl()                     # 3 b' Z^2
    f = e.tpl()
    h = yz.dbl()                    # 2 Y Z
    g = b + f
    x3 = (xy * (b - f)).dbl()       # 4 x arkworks' (X3, Y3, Z3): a = XY/2, g = (b + f)/2
    y3 = g.sqr() - e.sqr().scale(12)
    z3 = (b * h).scale(4)
    zb3 = (b * yzb).scale(8)
    return (h, -j.tpl(), b - e), (x3, y3, z3, zb3)


def proj_add(T, x2, y2):
    """T + Q for affine Q = (x2, y2), and the chord line"""
    X, Y, Z, Zb = T
    theta = Y - y2 * Z
    lam = X - x2 * Z
    c, d = theta.sqr(), lam.sqr()
    zl, xl, tx, tl, tz, yl, zbl = Z * lam, X * lam, theta * X, theta * lam, theta * Z, Y * lam, Zb * lam
    jj = theta * x2 - lam * y2
    x3 = d.sqr() + zl * c - (xl * d).dbl()                 # lam (lam^3 + Z theta^2 - 2 X lam^2)
    y3 = (tx.tpl() - tl - yl) * d - tz * c                 # theta (3 X lam^2 - lam^3 - Z theta^2) - Y lam^3
    z3 = zl * d
    zb3 = zbl * d
    return (lam, -theta, jj), (x3, y3, z3, zb3)


def miller_shallow(t, P, Qx, Qy):
    one, bt = _const(t, gp.C_ONE), _const(t, C_TWIST_B)
    T = (Qx, Qy, one, bt)
    nQy = -Qy
    f = None
    for i in range(63, -1, -1):
        line, T = proj_dbl(T)
        if f is None:
            zero = EV(t, {})
            f = [line[0].mulfq(P, 1), line[1].mulfq(P, 0), zero, line[2], zero, zero]
        else:
            f = gp.apply_line(gp.f12_sqr(f), line, P)
        dgt = gp.ATE[i]
        if dgt:
            line, T = proj_add(T, Qx, Qy if dgt > 0 else nQy)
            f = gp.apply_line(f, line, P)
    twx, twy = _const(t, gp.C_TWX), _const(t, gp.C_TWY)
    q1x, q1y = Qx.conj() * twx, Qy.conj() * twy
    q2x, q2y = q1x.conj() * twx, -(q1y.conj() * twy)
    line, T = proj_add(T, q1x, q1y)
    f = gp.apply_line(f, line, P)
    line, T = proj_add(T, q2x, q2y)
    f = gp.apply_line(f, line, P)
    return f


# Fq12 routines with ONE product level and no combination level in front of it: every product operand is a materialised
# value or the plain sum of two (folded into the product step).  They spend products (idle lanes) to save depth: 18 instead
# of 12 for the square, 16 instead of 13 for the sparse product, 24 instead of 18 for the full product.
def f12_sqr_shallow(a):
    a0, a1 = gp.halves(a)
    s0, s1, m = gp.f6_mul(a0, a0), gp.f6_mul(a1, a1), gp.f6_mul(a0, a1)
    return gp.from_halves(gp.f6_add(s0, gp.f6_mul_v(s1)), gp.f6_add(m, m))


def f12_mul_shallow(a, b):
    a0, a1 = gp.halves(a)
    b0, b1 = gp.halves(b)
    c0 = gp.f6_add(gp.f6_mul(a0, b0), gp.f6_mul_v(gp.f6_mul(a1, b1)))
    c1 = gp.f6_add(gp.f6_mul(a0, b1), gp.f6_mul(a1, b0))
    return gp.from_halves(c0, c1)


def mul_by_line_shallow(f, l0, l1, l3):
    f0, f1 = gp.halves(f)
    c0 = gp.f6_add(gp.f6_mul_f2(f0, l0), gp.f6_mul_v(gp.f6_mul_01(f1, l1, l3)))
    c1 = gp.f6_add(gp.f6_mul_01(f0, l1, l3), gp.f6_mul_f2(f1, l0))
    return gp.from_halves(c0, c1)


def trace(what):
    """what: "pairing" (inputs P, Qx, Qy and the split coordinates (xP, 0), (yP, 0); outputs GT in tower order) or
    "gt_bases" (input: a cyclotomic Fq12 in tower order; outputs a^(2^(8 w)), w = 0..31, tower order each)."""
    t = WTrace()
    saved = (gp.const, gp.zero_value, gp.USE_MACROS, gp.f12_sqr, gp.f12_mul, gp.mul_by_line)
    gp.const, gp.zero_value, gp.USE_MACROS = _const, _zero_value, False
    if SHALLOW:
        gp.f12_sqr, gp.f12_mul, gp.mul_by_line = f12_sqr_shallow, f12_mul_shallow, mul_by_line_shallow
    try:
        if what == "pairing":
            p, qx, qy = t.new_input(), t.new_input(), t.new_input()
            xp, yp = t.new_input(), t.new_input()
            t.split = {(p, 0): xp, (p, 1): yp}
            f = miller_shallow(t, EV.of(t, p), EV.of(t, qx), EV.of(t, qy))
            f = gp.final_exp(t, f)
            c0, c1 = gp.halves(f)
            outs = list(c0 + c1)
        elif what == "gt_prod":
            # product of GT_PROD_N Fq12 values (tower order each): chunks of 8 reduced one after the other (phases), so that
            # at most one chunk's 96 level-1 products are alive next to the inputs
            elems = []
            for _ in range(GT_PROD_N):
                tower = [EV.of(t, t.new_input()) for _ in range(6)]
                elems.append([tower[0], tower[3], tower[1], tower[4], tower[2], tower[5]])

            def tree(xs):
                while len(xs) > 1:
                    nxt = [[c.force() for c in gp.f12_mul(xs[i], xs[i + 1])] for i in range(0, len(xs) - 1, 2)]
                    if len(xs) & 1:
                        nxt.append(xs[-1])
                    xs = nxt
                return xs[0]
            partial = []
            for c in range(0, GT_PROD_N, 8):
                t.phase = c // 8
                partial.append(tree(elems[c:c + 8]))
            t.phase = GT_PROD_N // 8
            c0, c1 = gp.halves(tree(partial))
            outs = list(c0 + c1)
        else:
            ids = [t.new_input() for _ in range(6)]
            tower = [EV.of(t, i) for i in ids]
            g = [tower[0], tower[3], tower[1], tower[4], tower[2], tower[5]]   # tower order -> w-basis
            outs = []
            for w in range(32):
                g = [x.force() for x in g]
                c0, c1 = gp.halves(g)
                outs += list(c0 + c1)
                if w < 31:
                    for _ in range(8):
                        g = gp.cyclotomic_sqr(g)
    finally:
        gp.const, gp.zero_value, gp.USE_MACROS, gp.f12_sqr, gp.f12_mul, gp.mul_by_line = saved
    t.outputs = []
    for x in outs:
        if x.is_zero():
            t.outputs.append(ZERO_VID)
        else:
            v = next(iter(x.force().terms))
            if t.nodes[v] is None:       # an input / constant passed through: copy it so that outputs are program results
                v = t.emit_lin([(v, 1)], [])
            t.outputs.append(v)
    return t


# ---------------------------------------------------------------------------------------------
# scheduling and slot allocation
# ---------------------------------------------------------------------------------------------
def lin_cost(nu, nw):
    return 150 + 40 * nu + 75 * nw + 110


def schedule(t):
    """ASAP levels; inside a level one step per kind and 32 nodes (LIN nodes grouped by length: a step costs as much as its
    longest lane, and its plain + xi term counts must fit one descriptor)"""
    live = set()
    stack = [v for v in t.outputs if v >= 0]
    while stack:
        v = stack.pop()
        if v in live:
            continue
        live.add(v)
        nd = t.nodes[v]
        if nd is not None:
            stack += nd.deps()
    users = {v: [] for v in live}
    level = {}
    # rounds: [LIN-phase steps, MUL steps]; level = 2 * round + (1 for MUL)
    for v in sorted(live):                 # ids are topologically ordered
        nd = t.nodes[v]
        if nd is None:
            level[v] = -1
            continue
        for s in set(nd.deps()):
            users[s].append(v)
        lv = 0
        for s in nd.deps():
            ls = level[s]
            if ls < 0:
                continue
            if nd.kind == "MUL":
                # after a LIN of the same round, or a MUL of the round before
                lv = max(lv, ls + 1 if ls % 2 == 0 else ls + 2)
            else:
                lv = max(lv, ls + 2 if ls % 2 == 0 else ls + 1)
        if nd.kind == "MUL" and lv % 2 == 0:
            lv += 1
        if nd.kind != "MUL" and lv % 2 == 1:
            lv += 1
        level[v] = lv
    # ALAP adjustment for LIN nodes would go here; ASAP keeps slots busy a little longer but is simple
    by_level = {}
    for v in live:
        if t.nodes[v] is not None:
            by_level.setdefault((t.nodes[v].phase, level[v]), []).append(v)
    steps = []
    for lv in sorted(by_level):
        nodes = by_level[lv]
        for kind in KINDS[1:]:
            group = sorted(v for v in nodes if t.nodes[v].kind == kind)
            if not group:
                continue
            if kind != "LIN":
                for i in range(0, len(group), LANES):
                    steps.append((kind, group[i:i + LANES]))
                continue
            group.sort(key=lambda v: (len(t.nodes[v].w) > 0, len(t.nodes[v].u) + len(t.nodes[v].w)))
            cur, mu, mw = [], 0, 0
            for v in group:
                nd = t.nodes[v]
                nu2, nw2 = max(mu, len(nd.u)), max(mw, len(nd.w))
                # open a new step when the descriptor would overflow, the step is full, or when mixing would cost more
                # than a second step
                split = len(cur) == LANES // 2 or nu2 + nw2 > MAX_TERMS
                if cur and not split and lin_cost(nu2, nw2) > lin_cost(mu, mw) + 400 and len(group) > LANES // 2:
                    split = True
                if split:
                    steps.append(("LIN", cur))
                    cur, mu, mw = [], 0, 0
                    nu2, nw2 = len(nd.u), len(nd.w)
                cur.append(v)
                mu, mw = nu2, nw2
            if cur:
                steps.append(("LIN", cur))
    return steps, users, live


def allocate(t, steps, users, live):
    """first-fit slots; inputs and constants occupy the first slots after the zero slot"""
    step_of = {}
    for si, (_, take) in enumerate(steps):
        for v in take:
            step_of[v] = si
    last_use = {}
    for v in live:
        us = [step_of[u] for u in users[v]]
        last_use[v] = max(us) if us else -1
    for o in t.outputs:
        if o >= 0:
            last_use[o] = len(steps) + 1   # outputs stay to the end
    slot = {}
    nxt = 1
    for v in t.inputs:
        slot[v] = nxt
        nxt += 1
    const_slots = {}
    for idx, v in sorted(t.consts.items()):
        if v in live:
            slot[v] = nxt
            const_slots[idx] = nxt
            nxt += 1
    n_fixed = nxt
    free, high = [], n_fixed
    dying = {}
    for v in live:
        if 0 <= last_use[v] <= len(steps):
            dying.setdefault(last_use[v], []).append(v)
    for si, (_, take) in enumerate(steps):
        for v in dying.get(si, []):       # operands read for the last time in this step: their slots can take results
            if slot[v] >= n_fixed:
                free.append(slot[v])
        free.sort(reverse=True)
        for v in take:
            if free:
                slot[v] = free.pop()
            else:
                slot[v] = high
                high += 1
    assert high <= MAX_SLOTS, "slot file too small: %d" % high
    return slot, const_slots, high


def build(what):
    """descriptor words, 8 x u32 per lane per step:
         w0 = kind | active << 4 | dst << 8 | half << 17 | nU << 20 | nW << 24   (nU, nW: the step's loop bounds, same in
              all lanes; half: a LIN node occupies two adjacent lanes, lane `half` computes coordinate c_half of the result)
         MUL: w1 = a1 | a2 << 16, w2 = b1 | b2 << 16:  d = (a1 + a2) * (b1 + b2), the a-side sum not reduced
         LIN: terms t0..t13 of 16 bits (slot | sign << 9 | (|c| - 1) << 10), two per word in w1..w7; t0..t(nU-1) are the
              plain terms, the next nW the terms multiplied by xi; absent terms name the zero slot (coefficient 1).
              The integer value sum |c_u| + 10 sum |c_w| stays below LIN_BOUND (the kernel adds 128 q before reducing)
         CONJ / INV: w1 = source slot"""
    t = trace(what)
    steps, users, live = schedule(t)
    slot, const_slots, nslots = allocate(t, steps, users, live)
    assert nslots <= 512

    def s_of(v):
        return ZERO_SLOT if v < 0 else slot[v]
    words = []
    hist, cost = {}, 0
    for kind, take in steps:
        hist[kind] = hist.get(kind, 0) + 1
        nu = max([len(t.nodes[v].u) for v in take]) if kind == "LIN" else 0
        nw = max([len(t.nodes[v].w) for v in take]) if kind == "LIN" else 0
        assert nu + nw <= MAX_TERMS, (nu, nw)
        cost += {"MUL": 800, "CONJ": 100, "INV": 40000}.get(kind, lin_cost(nu, nw))
        per_node = 2 if kind == "LIN" else 1        # a LIN node takes two lanes: one per coordinate of the Fq2 result
        assert len(take) * per_node <= LANES
        for lane in range(LANES):
            w0 = KIND[kind] | (nu << 20) | (nw << 24)
            if lane >= len(take) * per_node:
                words += [w0, 0, 0, 0, 0, 0, 0, 0]
                continue
            v = take[lane // per_node]
            nd = t.nodes[v]
            w0 |= (1 << 4) | (slot[v] << 8)
            if kind == "LIN":
                w0 |= (lane & 1) << 17
                assert sum(abs(c) for _, c in nd.u) + 10 * sum(abs(c) for _, c in nd.w) <= LIN_BOUND

                def enc(x, c):
                    assert 1 <= abs(c) <= MAX_COEF
                    return s_of(x) | (int(c < 0) << 9) | ((abs(c) - 1) << 10)
                tt = [0] * 14
                for k, (x, c) in enumerate(nd.u):
                    tt[k] = enc(x, c)
                for k, (x, c) in enumerate(nd.w):
                    tt[nu + k] = enc(x, c)
                words += [w0] + [tt[2 * k] | (tt[2 * k + 1] << 16) for k in range(7)]
            elif kind == "MUL":
                a1, a2, b1, b2 = (s_of(x) for x in nd.srcs)
                words += [w0, a1 | (a2 << 16), b1 | (b2 << 16), 0, 0, 0, 0, 0]
            else:
                words += [w0, s_of(nd.srcs[0]), 0, 0, 0, 0, 0, 0]
    outs = [s_of(v) for v in t.outputs]
    stats = {"steps": len(steps), "hist": hist, "nodes": sum(len(tk) for _, tk in steps), "slots": nslots,
             "mul_nodes": sum(len(tk) for k, tk in steps if k == "MUL"), "model_instr": cost}
    return {"words": words, "nsteps": len(steps), "nslots": nslots, "inputs": [slot[v] for v in t.inputs],
            "consts": const_slots, "outs": outs, "stats": stats}


# ---------------------------------------------------------------------------------------------
# simulator (plain integers) - used by the tests
# ---------------------------------------------------------------------------------------------
def simulate(prog, inputs):
    """inputs: list of Fq2 (tuples) for prog["inputs"]; returns the Fq2 values of prog["outs"]"""
    S = [(0, 0)] * prog["nslots"]
    for s, v in zip(prog["inputs"], inputs):
        S[s] = v
    for idx, s in prog["consts"].items():
        S[s] = WP_CONSTS[idx]
    w = prog["words"]
    for step in range(prog["nsteps"]):
        writes = []
        for lane in range(LANES):
            d8 = w[8 * (step * LANES + lane): 8 * (step * LANES + lane) + 8]
            w0, w1, w2 = d8[0], d8[1], d8[2]
            if not (w0 >> 4) & 1:
                continue
            kind, dst, nu, nw, half = KINDS[w0 & 15], (w0 >> 8) & 511, (w0 >> 20) & 15, (w0 >> 24) & 15, (w0 >> 17) & 1
            if kind == "MUL":
                a1, a2, b1, b2 = w1 & 0xFFFF, w1 >> 16, w2 & 0xFFFF, w2 >> 16
                a = ((S[a1][0] + S[a2][0]) % Q, (S[a1][1] + S[a2][1]) % Q)
                b = ((S[b1][0] + S[b2][0]) % Q, (S[b1][1] + S[b2][1]) % Q)
                r = gp.f2_mul(a, b)
            elif kind == "LIN":
                tt = [(x >> (16 * k)) & 0xFFFF for x in d8[1:] for k in range(2)]
                acc = 0
                for k in range(nu + nw):
                    c = (-1 if (tt[k] >> 9) & 1 else 1) * (((tt[k] >> 10) & 63) + 1)
                    x = S[tt[k] & 511]
                    if k < nu:
                        acc += c * x[half]
                    else:
                        acc += 9 * c * x[half] + (c if half else -c) * x[1 - half]
                writes.append((dst, half, acc % Q))
                continue
            elif kind == "CONJ":
                r = (S[w1][0], -S[w1][1] % Q)
            elif kind == "INV":
                r = gp.f2_inv(S[w1]) if S[w1] != (0, 0) else (0, 0)
            else:
                raise ValueError(kind)
            writes.append((dst, 0, r[0]))
            writes.append((dst, 1, r[1]))
        for dst, half, r in writes:
            S[dst] = (r, S[dst][1]) if half == 0 else (S[dst][0], r)
    return [S[s] for s in prog["outs"]]


# ---------------------------------------------------------------------------------------------
# emit
# ---------------------------------------------------------------------------------------------
def compact(prog):
    """(step headers, node stream) of the dense descriptor array: per step one header word kind | nactive << 4 | nU << 20 |
    nW << 24, per active lane (LIN: per pair of lanes) the destination slot followed by its payload (MUL: 2 words, LIN: the
    nU + nW terms two per word, CONJ / INV: 1 word).  wpprog::expand() (emitted below) rebuilds the dense array the kernel reads."""
    w = prog["words"]
    hdr, stream = [], []
    for step in range(prog["nsteps"]):
        d0 = w[8 * LANES * step]
        kind, nu, nw = d0 & 15, (d0 >> 20) & 15, (d0 >> 24) & 15
        nact = 0
        for lane in range(LANES):
            d = w[8 * (LANES * step + lane): 8 * (LANES * step + lane) + 8]
            if not (d[0] >> 4) & 1:
                assert lane >= nact   # active lanes come first
                continue
            assert lane == nact
            nact += 1
            if KINDS[kind] == "LIN" and lane & 1:
                continue                      # the second lane of a LIN node repeats the first
            stream.append((d[0] >> 8) & 511)
            if KINDS[kind] == "MUL":
                stream += [d[1], d[2]]
            elif KINDS[kind] == "LIN":
                stream += d[1:1 + (nu + nw + 1) // 2]
            else:
                stream.append(d[1])
        hdr.append(kind | (nact << 4) | (nu << 20) | (nw << 24))
    return hdr, stream


EXPAND = """// dense descriptor array (8 words per lane, 32 lanes per step) from the compact form
static inline void expand(const Program& p, uint32_t* words) {
  const uint32_t* s = p.stream;
  for (int step = 0; step < p.nsteps; step++) {
    const uint32_t h = p.hdr[step], kind = h & 15u, nact = (h >> 4) & 63u, nu = (h >> 20) & 15u, nw = (h >> 24) & 15u;
    for (uint32_t lane = 0; lane < 32; lane++) {
      uint32_t* d = words + 8 * (32 * (size_t)step + lane);
      for (int k = 0; k < 8; k++) d[k] = 0;
      d[0] = kind | (nu << 20) | (nw << 24);
      if (lane >= nact) continue;
      if (kind == K_LIN && (lane & 1u)) {   // second lane of a LIN node: same descriptor, other coordinate
        for (int k = 0; k < 8; k++) d[k] = d[k - 8];
        d[0] |= 1u << 17;
        continue;
      }
      d[0] |= (1u << 4) | (*s++ << 8);
      if (kind == K_MUL) { d[1] = *s++; d[2] = *s++; }
      else if (kind == K_LIN) { for (uint32_t k = 0; k < (nu + nw + 1) / 2; k++) d[1 + k] = *s++; }
      else d[1] = *s++;
    }
  }
}"""


def expand_py(prog, hdr, stream):
    """Python twin of the emitted expand() (the generator checks its own compaction with it)"""
    out, pos = [], 0
    for step in range(prog["nsteps"]):
        h = hdr[step]
        kind, nact, nu, nw = h & 15, (h >> 4) & 63, (h >> 20) & 15, (h >> 24) & 15
        for lane in range(LANES):
            d = [kind | (nu << 20) | (nw << 24)] + [0] * 7
            if lane < nact:
                if KINDS[kind] == "LIN" and lane & 1:
                    d = list(out[-8:])
                    d[0] |= 1 << 17
                    out += d
                    continue
                d[0] |= (1 << 4) | (stream[pos] << 8)
                pos += 1
                if KINDS[kind] == "MUL":
                    d[1], d[2] = stream[pos], stream[pos + 1]
                    pos += 2
                elif KINDS[kind] == "LIN":
                    n = (nu + nw + 1) // 2
                    d[1:1 + n] = stream[pos:pos + n]
                    pos += n
                else:
                    d[1] = stream[pos]
                    pos += 1
            out += d
    assert pos == len(stream)
    return out


def main():
    out = ["// GENERATED by tools/gen_pairing_warp.py - do not edit.",
           "// Lane-parallel schedules for the warp-cooperative pairing kernel (pairing_warp.cuh); descriptor format: see build()",
           "// in the generator and wp::lane_compute in pairing_warp.cuh.  Stored compactly (compact() in the generator).",
           "#pragma once", "#include <stddef.h>", "#include <stdint.h>", "namespace kb { namespace wpprog {",
           "enum Kind : uint32_t { " + ", ".join("K_%s = %d" % (n, i) for i, n in enumerate(KINDS)) + " };",
           "struct Program { int nsteps, nslots, ninputs, nconsts, nouts; const uint16_t* inputs; const uint16_t* const_idx; const uint16_t* const_slot; "
           "const uint16_t* outs; const uint32_t* hdr; const uint32_t* stream; };"]
    out.append("static const int NUM_CONSTS = %d;" % len(WP_CONSTS))
    out.append("// Fq2 constants (Montgomery form): " + ", ".join(WP_CONST_NAMES))
    out.append("static const uint32_t CONSTS[%d * 16] = {" % len(WP_CONSTS))
    for c in WP_CONSTS:
        out.append("    " + gp.limbs(c[0] * gp.MONT % Q) + ", " + gp.limbs(c[1] * gp.MONT % Q) + ",")
    out.append("};")
    names = []
    for what in ("pairing", "gt_bases", "gt_prod"):
        p = build(what)
        hdr, stream = compact(p)
        assert expand_py(p, hdr, stream) == p["words"]
        name = what.upper()
        st = p["stats"]
        out.append("// %s: %d steps (%s), %d nodes (%d products), %d slots, modelled %d instructions on the critical lane"
                   % (name, st["steps"], " ".join("%s=%d" % kv for kv in sorted(st["hist"].items())), st["nodes"], st["mul_nodes"], st["slots"], st["model_instr"]))
        print(name, st, file=sys.stderr)
        out.append("static const uint16_t %s_INPUTS[%d] = {%s};" % (name, len(p["inputs"]), ", ".join(map(str, p["inputs"]))))
        cs = sorted(p["consts"].items())
        out.append("static const uint16_t %s_CONST_IDX[%d] = {%s};" % (name, max(1, len(cs)), ", ".join(str(i) for i, _ in cs) or "0"))
        out.append("static const uint16_t %s_CONST_SLOT[%d] = {%s};" % (name, max(1, len(cs)), ", ".join(str(s) for _, s in cs) or "0"))
        out.append("static const uint16_t %s_OUTS[%d] = {%s};" % (name, len(p["outs"]), ", ".join(map(str, p["outs"]))))
        for arr, vals in (("HDR", hdr), ("STREAM", stream)):
            out.append("static const uint32_t %s_%s[%d] = {" % (name, arr, len(vals)))
            for i in range(0, len(vals), 10):
                out.append("    " + ", ".join("0x%xu" % x for x in vals[i:i + 10]) + ",")
            out.append("};")
        names.append((name, p, len(cs)))
    for name, p, ncs in names:
        out.append("static const Program %s = {%d, %d, %d, %d, %d, %s_INPUTS, %s_CONST_IDX, %s_CONST_SLOT, %s_OUTS, %s_HDR, %s_STREAM};"
                   % (name, p["nsteps"], p["nslots"], len(p["inputs"]), ncs, len(p["outs"]), name, name, name, name, name, name))
    out.append(EXPAND)
    out.append("}}  // namespace kb::wpprog")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "keaki_b200", "csrc", "pairing_warp_gen.cuh")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
