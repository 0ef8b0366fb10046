#!/usr/bin/env python3
"""Generator + bit-exact emulator for the sm_100a Fp (BLS12-381 base field) inline-PTX primitives.

Fp elements are 12 x 32-bit limbs (little-endian limb order) held in registers, Montgomery form with
R = 2^384.  The hot primitive is the Montgomery product built from `mad.lo.cc.u32` / `madc.hi.cc.u32`
carry chains.  A 32x32 product's low half lands in column j and its high half in column j+1, so the
products a[j]*b_i for EVEN j form one unbroken carry chain over columns 0..11 and those for ODD j
another over columns 1..12.  Two accumulators (`e` aligned at column 0, `o` aligned at column 1) are
kept, their roles swap after every row (one limb of the running sum is retired per row by the
Montgomery reduction step), and they are merged once at the end.  The modulus limbs and -p^-1 mod 2^32
are immediates.

This script
  * builds each primitive as a list of PTX instructions,
  * EMULATES that exact instruction list in Python (32-bit registers + the CC.CF flag) against
    big-integer arithmetic on random and edge-case operands (no GPU exists in the build container), and
  * writes crypto12381_b200/csrc/fp_ptx.inc (device asm wrappers).

Usage: python tools/gen_fp_ptx.py [--check-only]
"""
import os
import random
import sys

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
N = 12
MASK = 0xFFFFFFFF
PL = [(P >> (32 * i)) & MASK for i in range(N)]
M0 = (-pow(P, -1, 1 << 32)) & MASK
RMONT = 1 << 384


class Prog:
    """A straight-line PTX fragment over named 32-bit registers."""

    def __init__(self):
        self.ins = []
        self.regs = []
        self.preds = []

    def reg(self, name):
        if name not in self.regs:
            self.regs.append(name)
        return name

    def pred(self, name):
        if name not in self.preds:
            self.preds.append(name)
        return name

    def emit(self, op, *args):
        self.ins.append((op, args))

    # ---- emulator -------------------------------------------------------------------------------
    def run(self, env):
        cf = 0
        preds = {}

        def val(x):
            if isinstance(x, int):
                return x & MASK
            return env[x]

        for op, a in self.ins:
            if op == "mov.u32":
                env[a[0]] = val(a[1])
            elif op == "mul.lo.u32":
                env[a[0]] = (val(a[1]) * val(a[2])) & MASK
            elif op == "mul.hi.u32":
                env[a[0]] = (val(a[1]) * val(a[2])) >> 32
            elif op in ("mad.lo.cc.u32", "madc.lo.cc.u32", "madc.lo.u32", "mad.lo.u32"):
                t = ((val(a[1]) * val(a[2])) & MASK) + val(a[3]) + (cf if op.startswith("madc") else 0)
                env[a[0]] = t & MASK
                if ".cc" in op:
                    cf = t >> 32
            elif op in ("mad.hi.cc.u32", "madc.hi.cc.u32", "madc.hi.u32", "mad.hi.u32"):
                t = ((val(a[1]) * val(a[2])) >> 32) + val(a[3]) + (cf if op.startswith("madc") else 0)
                env[a[0]] = t & MASK
                if ".cc" in op:
                    cf = t >> 32
            elif op in ("add.cc.u32", "addc.cc.u32", "addc.u32", "add.u32"):
                t = val(a[1]) + val(a[2]) + (cf if op.startswith("addc") else 0)
                env[a[0]] = t & MASK
                if ".cc" in op:
                    cf = t >> 32
            elif op in ("sub.cc.u32", "subc.cc.u32", "subc.u32", "sub.u32"):
                t = val(a[1]) - val(a[2]) - (cf if op.startswith("subc") else 0)
                env[a[0]] = t & MASK
                if ".cc" in op:
                    cf = 1 if t < 0 else 0
            elif op == "setp.ne.u32":
                preds[a[0]] = val(a[1]) != val(a[2])
            elif op == "setp.eq.u32":
                preds[a[0]] = val(a[1]) == val(a[2])
            elif op == "selp.u32":
                env[a[0]] = val(a[1]) if preds[a[3]] else val(a[2])
            elif op == "assert.nocarry":
                assert cf == 0, "carry cell overflowed"
            elif op == "assert.zero":
                assert env[a[0]] == 0, "carry out of the top column"
            elif op == "and.b32":
                env[a[0]] = val(a[1]) & val(a[2])
            elif op == "or.b32":
                env[a[0]] = val(a[1]) | val(a[2])
            else:
                raise ValueError(op)
        return env

    # ---- PTX text -------------------------------------------------------------------------------
    def ptx(self, operand_map):
        """operand_map: register name -> '%k' for asm operands; other names become local .reg."""
        local = [r for r in self.regs if r not in operand_map]
        lines = ["{"]
        if local:
            lines.append(".reg .u32 " + ", ".join(local) + ";")
        if self.preds:
            lines.append(".reg .pred " + ", ".join(self.preds) + ";")

        def fmt(x):
            if isinstance(x, int):
                return "0x%08x" % (x & MASK)
            return operand_map.get(x, x)

        for op, a in self.ins:
            if op.startswith("assert."):
                continue
            lines.append("%s %s;" % (op, ", ".join(fmt(x) for x in a)))
        lines.append("}")
        return lines


def final_reduce(pr, src, dst, extra_hi=None):
    """dst = src - p if src >= p else src  (src < 2p).  extra_hi: optional 13th limb register (0/1)."""
    t = [pr.reg("t%d" % i) for i in range(N)]
    pr.emit("sub.cc.u32", t[0], src[0], PL[0])
    for i in range(1, N):
        pr.emit("subc.cc.u32", t[i], src[i], PL[i])
    bw = pr.reg("bw")
    if extra_hi is None:
        pr.emit("subc.u32", bw, 0, 0)  # 0xffffffff if src < p
    else:
        pr.emit("subc.u32", bw, extra_hi, 0)  # carry limb absorbs the borrow: all-ones only if still negative
    pb = pr.pred("pb")
    pr.emit("setp.ne.u32", pb, bw, 0)
    for i in range(N):
        pr.emit("selp.u32", dst[i], src[i], t[i], pb)


def gen_mul(square=False):
    """Montgomery product r = a*b*2^-384 mod p, a < p, b < 2^384 (normally < p), output fully reduced.  Registers: a0..a11, b0..b11 in, r0..r11 out."""
    pr = Prog()
    a = [pr.reg("a%d" % i) for i in range(N)]
    b = a if square else [pr.reg("b%d" % i) for i in range(N)]
    r = [pr.reg("r%d" % i) for i in range(N)]
    # two accumulators of 12 limbs each; X is column-0 aligned ("even role"), Y column-1 aligned
    X = [pr.reg("e%d" % i) for i in range(N)]
    Y = [pr.reg("o%d" % i) for i in range(N)]
    m = pr.reg("m")

    for i in range(N):
        bi = b[i]
        if i == 0:
            # fresh products, no carries: Y <- a_odd*b0 (cols 1..12), X <- a_even*b0 (cols 0..11)
            for j in range(0, N, 2):
                pr.emit("mul.lo.u32", Y[j], a[j + 1], bi)
                pr.emit("mul.hi.u32", Y[j + 1], a[j + 1], bi)
            for j in range(0, N, 2):
                pr.emit("mul.lo.u32", X[j], a[j], bi)
                pr.emit("mul.hi.u32", X[j + 1], a[j], bi)
        else:
            # After the previous row's shift: running sum = X(aligned 0) + (Yold >> 32) where Yold[0] == 0
            # was retired... roles were swapped at the end of the previous row, so here X is the old
            # odd-aligned accumulator (now column-0 aligned) and Y is the old even-aligned one whose
            # limb k sits at column k-1.  Fold Y[1] (column 0) into X[0]; Y[k+2] becomes the new
            # column-(k+1) limb, i.e. new Y[k], while accumulating a_odd*b_i.
            pr.emit("add.cc.u32", X[0], X[0], Y[1])
            for j in range(0, N - 2, 2):
                pr.emit("madc.lo.cc.u32", Y[j], a[j + 1], bi, Y[j + 2])
                pr.emit("madc.hi.cc.u32", Y[j + 1], a[j + 1], bi, Y[j + 3])
            pr.emit("madc.lo.cc.u32", Y[N - 2], a[N - 1], bi, 0)
            pr.emit("madc.hi.u32", Y[N - 1], a[N - 1], bi, 0)
            # X += a_even*b_i over columns 0..11, carry out into column 12 = Y[11]
            pr.emit("mad.lo.cc.u32", X[0], a[0], bi, X[0])
            pr.emit("madc.hi.cc.u32", X[1], a[0], bi, X[1])
            for j in range(2, N, 2):
                pr.emit("madc.lo.cc.u32", X[j], a[j], bi, X[j])
                pr.emit("madc.hi.cc.u32", X[j + 1], a[j], bi, X[j + 1])
            pr.emit("addc.u32", Y[N - 1], Y[N - 1], 0)
        # reduction step: m = X[0] * (-p^-1); add m*p so that column 0 becomes zero
        pr.emit("mul.lo.u32", m, X[0], M0)
        pr.emit("mad.lo.cc.u32", Y[0], m, PL[1], Y[0])
        pr.emit("madc.hi.cc.u32", Y[1], m, PL[1], Y[1])
        for j in range(2, N, 2):
            pr.emit("madc.lo.cc.u32", Y[j], m, PL[j + 1], Y[j])
            if j + 1 < N - 1:
                pr.emit("madc.hi.cc.u32", Y[j + 1], m, PL[j + 1], Y[j + 1])
            else:
                pr.emit("madc.hi.u32", Y[j + 1], m, PL[j + 1], Y[j + 1])
        pr.emit("mad.lo.cc.u32", X[0], m, PL[0], X[0])
        pr.emit("madc.hi.cc.u32", X[1], m, PL[0], X[1])
        for j in range(2, N, 2):
            pr.emit("madc.lo.cc.u32", X[j], m, PL[j], X[j])
            pr.emit("madc.hi.cc.u32", X[j + 1], m, PL[j], X[j + 1])
        pr.emit("addc.u32", Y[N - 1], Y[N - 1], 0)
        # shift by one limb == swap roles
        X, Y = Y, X
    # merge: result = X (aligned 0) + (Y >> 32), Y[0] == 0
    pr.emit("add.cc.u32", X[0], X[0], Y[1])
    for k in range(1, N - 1):
        pr.emit("addc.cc.u32", X[k], X[k], Y[k + 1])
    pr.emit("addc.u32", X[N - 1], X[N - 1], 0)
    final_reduce(pr, X, r)
    return pr


def gen_add():
    pr = Prog()
    a = [pr.reg("a%d" % i) for i in range(N)]
    b = [pr.reg("b%d" % i) for i in range(N)]
    r = [pr.reg("r%d" % i) for i in range(N)]
    s = [pr.reg("s%d" % i) for i in range(N)]
    pr.emit("add.cc.u32", s[0], a[0], b[0])
    for i in range(1, N):
        pr.emit("addc.cc.u32" if i < N - 1 else "addc.u32", s[i], a[i], b[i])
    final_reduce(pr, s, r)  # a+b < 2p < 2^384: no carry limb
    return pr


def gen_sub():
    pr = Prog()
    a = [pr.reg("a%d" % i) for i in range(N)]
    b = [pr.reg("b%d" % i) for i in range(N)]
    r = [pr.reg("r%d" % i) for i in range(N)]
    s = [pr.reg("s%d" % i) for i in range(N)]
    bw = pr.reg("bw")
    pr.emit("sub.cc.u32", s[0], a[0], b[0])
    for i in range(1, N):
        pr.emit("subc.cc.u32", s[i], a[i], b[i])
    pr.emit("subc.u32", bw, 0, 0)
    q = [pr.reg("q%d" % i) for i in range(N)]
    for i in range(N):
        pr.emit("and.b32", q[i], bw, PL[i])
    pr.emit("add.cc.u32", r[0], s[0], q[0])
    for i in range(1, N):
        pr.emit("addc.cc.u32" if i < N - 1 else "addc.u32", r[i], s[i], q[i])
    return pr


def gen_neg():
    """r = (a == 0) ? 0 : p - a"""
    pr = Prog()
    a = [pr.reg("a%d" % i) for i in range(N)]
    r = [pr.reg("r%d" % i) for i in range(N)]
    z = pr.reg("z")
    pr.emit("or.b32", z, a[0], a[1])
    for i in range(2, N):
        pr.emit("or.b32", z, z, a[i])
    pz = pr.pred("pz")
    pr.emit("setp.eq.u32", pz, z, 0)
    s = [pr.reg("s%d" % i) for i in range(N)]
    pr.emit("sub.cc.u32", s[0], PL[0], a[0])
    for i in range(1, N):
        pr.emit("subc.cc.u32" if i < N - 1 else "subc.u32", s[i], PL[i], a[i])
    for i in range(N):
        pr.emit("selp.u32", r[i], 0, s[i], pz)
    return pr


def gen_redc():
    """r = a * 2^-384 mod p (Montgomery reduction of a single-width value): mul by 1 specialised."""
    pr = Prog()
    a = [pr.reg("a%d" % i) for i in range(N)]
    r = [pr.reg("r%d" % i) for i in range(N)]
    one = [1] + [0] * (N - 1)
    # simply reuse the product code path with b = 1 would waste work; do the 12 reduction rows directly
    X = [pr.reg("e%d" % i) for i in range(N)]
    Y = [pr.reg("o%d" % i) for i in range(N)]
    m = pr.reg("m")
    for i in range(N):
        pr.emit("mov.u32", X[i], a[i])
        pr.emit("mov.u32", Y[i], 0)
    for i in range(N):
        if i > 0:
            pr.emit("add.cc.u32", X[0], X[0], Y[1])
            for j in range(0, N - 2):
                pr.emit("addc.cc.u32", Y[j], Y[j + 2], 0)
            pr.emit("addc.u32", Y[N - 2], 0, 0)
            pr.emit("mov.u32", Y[N - 1], 0)
        pr.emit("mul.lo.u32", m, X[0], M0)
        pr.emit("mad.lo.cc.u32", Y[0], m, PL[1], Y[0])
        pr.emit("madc.hi.cc.u32", Y[1], m, PL[1], Y[1])
        for j in range(2, N, 2):
            pr.emit("madc.lo.cc.u32", Y[j], m, PL[j + 1], Y[j])
            if j + 1 < N - 1:
                pr.emit("madc.hi.cc.u32", Y[j + 1], m, PL[j + 1], Y[j + 1])
            else:
                pr.emit("madc.hi.u32", Y[j + 1], m, PL[j + 1], Y[j + 1])
        pr.emit("mad.lo.cc.u32", X[0], m, PL[0], X[0])
        pr.emit("madc.hi.cc.u32", X[1], m, PL[0], X[1])
        for j in range(2, N, 2):
            pr.emit("madc.lo.cc.u32", X[j], m, PL[j], X[j])
            pr.emit("madc.hi.cc.u32", X[j + 1], m, PL[j], X[j + 1])
        pr.emit("addc.u32", Y[N - 1], Y[N - 1], 0)
        X, Y = Y, X
    pr.emit("add.cc.u32", X[0], X[0], Y[1])
    for k in range(1, N - 1):
        pr.emit("addc.cc.u32", X[k], X[k], Y[k + 1])
    pr.emit("addc.u32", X[N - 1], X[N - 1], 0)
    final_reduce(pr, X, r)
    del one
    return pr


# ---- wide (unreduced) products and their Montgomery reduction: lazy reduction for Fp2, dedicated squaring ----------------
W = 2 * N
P2L = [((P * P) >> (32 * i)) & MASK for i in range(W)]


def wide_product(pr, a, b, tag, square=False):
    """Emit T = a*b (or a*a) as 24 limbs; operands are ANY 384-bit values.  Returns the 24 register names.
    Products a[j]*b[i] whose low half lands on an even column go to accumulator E, the others to O: within one row the
    products of one parity occupy consecutive columns, so each is ONE carry chain; the carry out of a chain's top column
    lands in a cell that so far holds at most another such carry.  Squaring: cross products i < j only, doubled, plus the
    diagonal as one 24-column chain."""
    acc = {"e": [None] * (W + 1), "o": [None] * (W + 1)}
    cnt = [0]

    def fresh(kind, col):
        cnt[0] += 1
        return pr.reg("%s%s%d" % (tag, kind, col))

    for i in range(N):
        for par in (0, 1):
            js = [j for j in range(i + 1 if square else 0, N) if (i + j) % 2 == par]
            if not js:
                continue
            A = acc["e" if par == 0 else "o"]
            kind = "e" if par == 0 else "o"
            plain = all(A[i + j] is None and A[i + j + 1] is None for j in js)
            first = True
            for j in js:
                for half, col in (("lo", i + j), ("hi", i + j + 1)):
                    dst = A[col] if A[col] is not None else fresh(kind, col)
                    if plain:
                        pr.emit("mul.%s.u32" % half, dst, a[j], b[i])
                    else:
                        addend = A[col] if A[col] is not None else 0
                        pr.emit(("mad.%s.cc.u32" if first else "madc.%s.cc.u32") % half, dst, a[j], b[i], addend)
                        first = False
                    A[col] = dst
            if not plain:
                col = i + js[-1] + 2
                if col < W:
                    if A[col] is None:
                        A[col] = fresh(kind, col)
                        pr.emit("addc.u32", A[col], 0, 0)
                    else:
                        pr.emit("addc.cc.u32", A[col], A[col], 0)
                        pr.emit("assert.nocarry")
                else:
                    pr.emit("addc.cc.u32", pr.reg(tag + "dump"), 0, 0)   # carry out of column 23 must be zero
                    pr.emit("assert.zero", tag + "dump")
    E, O = acc["e"], acc["o"]
    T = [pr.reg("%st%d" % (tag, k)) for k in range(W)]
    first = True
    for k in range(W):
        x = E[k] if E[k] is not None else 0
        y = O[k] if O[k] is not None else 0
        if first:
            pr.emit("add.cc.u32", T[k], x, y)
            first = False
        else:
            pr.emit("addc.cc.u32" if k < W - 1 else "addc.u32", T[k], x, y)
    if square:
        pr.emit("add.cc.u32", T[0], T[0], T[0])
        for k in range(1, W):
            pr.emit("addc.cc.u32" if k < W - 1 else "addc.u32", T[k], T[k], T[k])
        for i in range(N):
            pr.emit("mad.lo.cc.u32" if i == 0 else "madc.lo.cc.u32", T[2 * i], a[i], a[i], T[2 * i])
            pr.emit("madc.hi.cc.u32" if i < N - 1 else "madc.hi.u32", T[2 * i + 1], a[i], a[i], T[2 * i + 1])
    return T


def redc_rows(pr, lo, tag):
    """(lo + M p) / 2^384 for the 12-limb value lo, M chosen limb by limb; result <= p, NOT reduced.  Returns 12 registers."""
    X = [pr.reg("%sx%d" % (tag, i)) for i in range(N)]
    Y = [pr.reg("%sy%d" % (tag, i)) for i in range(N)]
    m = pr.reg(tag + "m")
    for i in range(N):
        pr.emit("mov.u32", X[i], lo[i])
        pr.emit("mov.u32", Y[i], 0)
    for i in range(N):
        if i > 0:
            pr.emit("add.cc.u32", X[0], X[0], Y[1])
            for j in range(0, N - 2):
                pr.emit("addc.cc.u32", Y[j], Y[j + 2], 0)
            pr.emit("addc.u32", Y[N - 2], 0, 0)
            pr.emit("mov.u32", Y[N - 1], 0)
        pr.emit("mul.lo.u32", m, X[0], M0)
        pr.emit("mad.lo.cc.u32", Y[0], m, PL[1], Y[0])
        pr.emit("madc.hi.cc.u32", Y[1], m, PL[1], Y[1])
        for j in range(2, N, 2):
            pr.emit("madc.lo.cc.u32", Y[j], m, PL[j + 1], Y[j])
            pr.emit("madc.hi.cc.u32" if j + 1 < N - 1 else "madc.hi.u32", Y[j + 1], m, PL[j + 1], Y[j + 1])
        pr.emit("mad.lo.cc.u32", X[0], m, PL[0], X[0])
        pr.emit("madc.hi.cc.u32", X[1], m, PL[0], X[1])
        for j in range(2, N, 2):
            pr.emit("madc.lo.cc.u32", X[j], m, PL[j], X[j])
            pr.emit("madc.hi.cc.u32", X[j + 1], m, PL[j], X[j + 1])
        pr.emit("addc.u32", Y[N - 1], Y[N - 1], 0)
        X, Y = Y, X
    pr.emit("add.cc.u32", X[0], X[0], Y[1])
    for k in range(1, N - 1):
        pr.emit("addc.cc.u32", X[k], X[k], Y[k + 1])
    pr.emit("addc.u32", X[N - 1], X[N - 1], 0)
    return X


def wide_redc(pr, T, r, tag):
    """r = T 2^-384 mod p, fully reduced, for a 24-limb T < 4 p^2 (so that T_hi + (T_lo + M p)/R < 2p)."""
    X = redc_rows(pr, T[:N], tag)
    pr.emit("add.cc.u32", X[0], X[0], T[N])
    for k in range(1, N):
        pr.emit("addc.cc.u32" if k < N - 1 else "addc.u32", X[k], X[k], T[N + k])
    final_reduce(pr, X, r)


def gen_sqr_dedicated():
    """Montgomery squaring: 66 cross products doubled + 12 diagonal ones + the reduction = 234 multiply-adds instead of 300."""
    pr = Prog()
    a = [pr.reg("a%d" % i) for i in range(N)]
    r = [pr.reg("r%d" % i) for i in range(N)]
    T = wide_product(pr, a, a, "s", square=True)
    wide_redc(pr, T, r, "q")
    return pr


def gen_mulw():
    """t (24 limbs) = a * b, operands any 384-bit values.  Registers a0.., b0.. in, r0..r23 out."""
    pr = Prog()
    a = [pr.reg("a%d" % i) for i in range(N)]
    b = [pr.reg("b%d" % i) for i in range(N)]
    T = wide_product(pr, a, b, "w")
    for k in range(W):
        pr.emit("mov.u32", pr.reg("r%d" % k), T[k])
    return pr


def gen_redcw():
    """r = t 2^-384 mod p for a 24-limb t < 4 p^2.  Registers a0..a23 in, r0..r11 out."""
    pr = Prog()
    t = [pr.reg("a%d" % i) for i in range(W)]
    r = [pr.reg("r%d" % i) for i in range(N)]
    wide_redc(pr, t, r, "q")
    return pr


def gen_wide_addsub(kind):
    """24-limb helpers of the lazy Fp2 product.  'sub': r = a - b (a >= b);  'subp2': r = a - b + p^2;  'dbl': r = 2a."""
    pr = Prog()
    a = [pr.reg("a%d" % i) for i in range(W)]
    b = [pr.reg("b%d" % i) for i in range(W)] if kind != "dbl" else a
    r = [pr.reg("r%d" % i) for i in range(W)]
    if kind == "dbl":
        for k in range(W):
            pr.emit("add.cc.u32" if k == 0 else ("addc.cc.u32" if k < W - 1 else "addc.u32"), r[k], a[k], a[k])
        return pr
    src = a
    if kind == "subp2":
        u = [pr.reg("u%d" % i) for i in range(W)]
        for k in range(W):
            pr.emit("add.cc.u32" if k == 0 else ("addc.cc.u32" if k < W - 1 else "addc.u32"), u[k], a[k], P2L[k])
        src = u
    for k in range(W):
        pr.emit("sub.cc.u32" if k == 0 else ("subc.cc.u32" if k < W - 1 else "subc.u32"), r[k], src[k], b[k])
    return pr


def gen_narrow(kind):
    """12-limb helpers without reduction.  'add': r = a + b (< 2^384);  'subp': r = a - b + p (a, b < p)."""
    pr = Prog()
    a = [pr.reg("a%d" % i) for i in range(N)]
    b = [pr.reg("b%d" % i) for i in range(N)]
    r = [pr.reg("r%d" % i) for i in range(N)]
    src = a
    if kind == "subp":
        u = [pr.reg("u%d" % i) for i in range(N)]
        for k in range(N):
            pr.emit("add.cc.u32" if k == 0 else ("addc.cc.u32" if k < N - 1 else "addc.u32"), u[k], a[k], PL[k])
        for k in range(N):
            pr.emit("sub.cc.u32" if k == 0 else ("subc.cc.u32" if k < N - 1 else "subc.u32"), r[k], u[k], b[k])
    else:
        for k in range(N):
            pr.emit("add.cc.u32" if k == 0 else ("addc.cc.u32" if k < N - 1 else "addc.u32"), r[k], a[k], b[k])
    return pr


def runw(pr, a, b=None, na=N, nb=N, nr=N):
    env = {}
    for i in range(na):
        env["a%d" % i] = (a >> (32 * i)) & MASK
    if b is not None:
        for i in range(nb):
            env["b%d" % i] = (b >> (32 * i)) & MASK
    pr.run(env)
    return sum(env["r%d" % i] << (32 * i) for i in range(nr))


def check_wide():
    rnd = random.Random(381)
    rinv = pow(RMONT, -1, P)
    full = (1 << 384) - 1
    edge = [0, 1, P - 1, P, 2 * P - 2, full, full - 1, 1 << 383, MASK, MASK << 352, int("f" * 48 + "0" * 48, 16)]
    sqr, mulw, redcw = gen_sqr_dedicated(), gen_mulw(), gen_redcw()
    subw, subp2, dblw, addn, subp = (gen_wide_addsub("sub"), gen_wide_addsub("subp2"), gen_wide_addsub("dbl"), gen_narrow("add"),
                                     gen_narrow("subp"))
    vals = edge + [rnd.randrange(1 << 384) for _ in range(120)]
    for x in vals[:30]:
        for y in vals[:30]:
            assert runw(mulw, x, y, nr=W) == x * y, ("mulw", hex(x), hex(y))
    for x in vals:
        y = rnd.choice(vals)
        assert runw(mulw, x, y, nr=W) == x * y
    pv = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, 1 << 380] + [rnd.randrange(P) for _ in range(300)]
    for x in pv:
        assert runw(sqr, x) == x * x * rinv % P, ("sqr", hex(x))
    for t in [0, 1, P * P, 4 * P * P - 1, 2 * P * P, (P - 1) * (P - 1), RMONT - 1, RMONT, RMONT + 1] + [rnd.randrange(4 * P * P) for _ in range(400)]:
        assert runw(redcw, t, na=W) == t * rinv % P, ("redcw", hex(t))
    for _ in range(200):
        x, y = rnd.randrange(1 << 768), rnd.randrange(1 << 768)
        x, y = max(x, y), min(x, y)
        assert runw(subw, x, y, na=W, nb=W, nr=W) == x - y
        a, b = rnd.randrange(P * P), rnd.randrange(P * P)
        assert runw(subp2, a, b, na=W, nb=W, nr=W) == a - b + P * P
        assert runw(dblw, a, na=W, nr=W) == 2 * a
        u, v = rnd.randrange(P), rnd.randrange(P)
        assert runw(addn, u, v) == u + v
        assert runw(subp, u, v) == u - v + P
    # the lazy Fp2 product / squaring assembled from the primitives, as fp2.cuh does
    for _ in range(100):
        xa, xb, ya, yb = (rnd.choice(pv) for _ in range(4))
        t2 = runw(mulw, runw(addn, xa, xb), runw(addn, ya, yb), nr=W)
        t0, t1 = runw(mulw, xa, ya, nr=W), runw(mulw, xb, yb, nr=W)
        t2 = runw(subw, runw(subw, t2, t0, na=W, nb=W, nr=W), t1, na=W, nb=W, nr=W)
        ra = runw(redcw, runw(subp2, t0, t1, na=W, nb=W, nr=W), na=W)
        rb = runw(redcw, t2, na=W)
        assert ra == (xa * ya - xb * yb) * rinv % P and rb == (xa * yb + xb * ya) * rinv % P
        s0 = runw(redcw, runw(mulw, runw(addn, xa, xb), runw(subp, xa, xb), nr=W), na=W)
        s1 = runw(redcw, runw(dblw, runw(mulw, xa, xb, nr=W), na=W, nr=W), na=W)
        assert s0 == (xa * xa - xb * xb) * rinv % P and s1 == 2 * xa * xb * rinv % P
    for name, pr in (("fp_sqr (dedicated)", sqr), ("fp_mulw", mulw), ("fp_redcw", redcw)):
        counts = {}
        for op, _ in pr.ins:
            if op.startswith("assert"):
                continue
            k = op.split(".")[0]
            counts[k] = counts.get(k, 0) + 1
        print("emulation OK;", name, "instruction mix:", counts)


# ---------------------------------------------------------------------------------------------------
def to_limbs(x):
    return [(x >> (32 * i)) & MASK for i in range(N)]


def from_limbs(env, prefix):
    return sum(env[prefix + str(i)] << (32 * i) for i in range(N))


def run2(pr, a, b=None):
    env = {}
    for i, v in enumerate(to_limbs(a)):
        env["a%d" % i] = v
    if b is not None:
        for i, v in enumerate(to_limbs(b)):
            env["b%d" % i] = v
    pr.run(env)
    return from_limbs(env, "r")


def check():
    rnd = random.Random(12381)
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (1 << 380), (1 << 381) - 1 if (1 << 381) - 1 < P else P - 3,
            MASK, (MASK << 352) % P, int("f" * 95, 16) % P, RMONT % P, (RMONT * RMONT) % P]
    vals = edge + [rnd.randrange(P) for _ in range(300)]
    rinv = pow(RMONT, -1, P)
    mul, sqr, add, sub, neg, redc = gen_mul(), gen_mul(True), gen_add(), gen_sub(), gen_neg(), gen_redc()
    n = 0
    for x in vals[:40]:
        for y in vals[:40]:
            assert run2(mul, x, y) == x * y * rinv % P, ("mul", hex(x), hex(y))
            assert run2(add, x, y) == (x + y) % P, ("add", hex(x), hex(y))
            assert run2(sub, x, y) == (x - y) % P, ("sub", hex(x), hex(y))
            n += 1
    for x in vals:
        y = rnd.choice(vals)
        assert run2(mul, x, y) == x * y * rinv % P
        assert run2(sqr, x) == x * x * rinv % P
        assert run2(add, x, y) == (x + y) % P
        assert run2(sub, x, y) == (x - y) % P
        assert run2(neg, x) == (-x) % P
        assert run2(redc, x) == x * rinv % P
    # the SECOND operand may be any 384-bit value (first < p): used by the R=2^406 <-> 2^384 conversion
    for _ in range(200):
        x = rnd.randrange(1 << 384)
        y = rnd.randrange(P)
        assert run2(mul, y, x) == x * y * rinv % P
        assert run2(redc, x) == x * rinv % P
    for x in (RMONT - 1, RMONT - 2, P, P + 1, 2 * P - 1):
        for y in (0, 1, P - 1, rnd.randrange(P)):
            assert run2(mul, y, x) == x * y * rinv % P
    # BOTH operands below 2p (unreduced sums of two field elements: the Karatsuba operands of fp2.cuh): 4p^2/R + p < 2p
    wide = [2 * P - 1, 2 * P - 2, P, P + 1] + [rnd.randrange(2 * P) for _ in range(200)]
    for x in wide[:8]:
        for y in wide[:8]:
            assert run2(mul, x, y) == x * y * rinv % P
    for _ in range(1000):
        x, y = rnd.choice(wide), rnd.choice(wide)
        assert run2(mul, x, y) == x * y * rinv % P
    counts = {}
    for op, _ in mul.ins:
        counts[op.split(".")[0]] = counts.get(op.split(".")[0], 0) + 1
    print("emulation OK; fp_mul instruction mix:", counts, "total", len(mul.ins))


def wrapper(name, pr, n_in, la=N, lb=N, lr=N):
    """C++ wrapper text: void name(uint32_t (&r)[lr], const uint32_t (&a)[la][, const uint32_t (&b)[lb]])"""
    omap = {}
    for i in range(lr):
        omap["r%d" % i] = "%%%d" % i
    for i in range(la):
        omap["a%d" % i] = "%%%d" % (lr + i)
    if n_in == 2:
        for i in range(lb):
            omap["b%d" % i] = "%%%d" % (lr + la + i)
    # results are written only by the trailing selp/add instructions, after every input has been read?
    # not for all primitives -> route outputs through locals and copy at the end to be alias-safe.
    body = Prog()
    body.ins = list(pr.ins)
    body.regs = list(pr.regs)
    body.preds = list(pr.preds)
    rename = {"r%d" % i: "w%d" % i for i in range(lr)}
    body.regs = [rename.get(x, x) for x in body.regs]
    body.ins = [(op, tuple(rename.get(x, x) if isinstance(x, str) else x for x in a)) for op, a in body.ins]
    for i in range(lr):
        body.ins.append(("mov.u32", ("r%d" % i, "w%d" % i)))
        if "r%d" % i not in body.regs:
            body.regs.append("r%d" % i)
    lines = body.ptx(omap)
    sig = "uint32_t (&r)[%d], const uint32_t (&a)[%d]" % (lr, la) + (", const uint32_t (&b)[%d]" % lb if n_in == 2 else "")
    out = ["__device__ __forceinline__ void %s(%s)" % (name, sig), "{", "    asm("]
    for ln in lines:
        out.append('        "%s\\n\\t"' % ln)
    outs = ", ".join('"=r"(r[%d])' % i for i in range(lr))
    ins = ", ".join('"r"(a[%d])' % i for i in range(la))
    if n_in == 2:
        ins += ", " + ", ".join('"r"(b[%d])' % i for i in range(lb))
    out.append("        : " + outs)
    out.append("        : " + ins + ");")
    out.append("}")
    out.append("")
    return out


def main():
    check()
    check_wide()
    if "--check-only" in sys.argv:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "crypto12381_b200", "csrc", "fp_ptx.inc")
    text = ["// GENERATED by tools/gen_fp_ptx.py — do not edit.  sm_100a inline-PTX Fp primitives",
            "// (12 x 32-bit limbs, Montgomery R = 2^384, mad.lo.cc / madc.hi.cc carry chains; the instruction",
            "// lists in here were emulated bit-exactly against big-integer arithmetic by the generator).", ""]
    text += wrapper("fp_mul_ptx", gen_mul(), 2)
    text += wrapper("fp_sqr_ptx", gen_mul(True), 1)
    text += wrapper("fp_add_ptx", gen_add(), 2)
    text += wrapper("fp_sub_ptx", gen_sub(), 2)
    text += wrapper("fp_neg_ptx", gen_neg(), 1)
    text += wrapper("fp_redc_ptx", gen_redc(), 1)
    # wide products and lazy reduction (Fp2 products with two reductions instead of three; dedicated squaring)
    text += wrapper("fp_sqr_dedicated_ptx", gen_sqr_dedicated(), 1)
    text += wrapper("fp_mulw_ptx", gen_mulw(), 2, lr=W)
    text += wrapper("fp_redcw_ptx", gen_redcw(), 1, la=W)
    text += wrapper("fpw_sub_ptx", gen_wide_addsub("sub"), 2, la=W, lb=W, lr=W)
    text += wrapper("fpw_sub_addp2_ptx", gen_wide_addsub("subp2"), 2, la=W, lb=W, lr=W)
    text += wrapper("fpw_dbl_ptx", gen_wide_addsub("dbl"), 1, la=W, lr=W)
    text += wrapper("fp_add_noreduce_ptx", gen_narrow("add"), 2)
    text += wrapper("fp_sub_addp_ptx", gen_narrow("subp"), 2)
    with open(path, "w") as f:
        f.write("\n".join(text))
    print("wrote", path)


if __name__ == "__main__":
    main()
