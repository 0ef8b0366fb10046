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
    counts = {}
    for op, _ in mul.ins:
        counts[op.split(".")[0]] = counts.get(op.split(".")[0], 0) + 1
    print("emulation OK; fp_mul instruction mix:", counts, "total", len(mul.ins))


def wrapper(name, pr, n_in):
    """C++ wrapper text: void name(uint32_t (&r)[12], const uint32_t (&a)[12][, const uint32_t (&b)[12]])"""
    omap = {}
    for i in range(N):
        omap["r%d" % i] = "%%%d" % i
    for i in range(N):
        omap["a%d" % i] = "%%%d" % (N + i)
    if n_in == 2:
        for i in range(N):
            omap["b%d" % i] = "%%%d" % (2 * N + i)
    # results are written only by the trailing selp/add instructions, after every input has been read?
    # not for all primitives -> route outputs through locals and copy at the end to be alias-safe.
    body = Prog()
    body.ins = list(pr.ins)
    body.regs = list(pr.regs)
    body.preds = list(pr.preds)
    rename = {"r%d" % i: "w%d" % i for i in range(N)}
    body.regs = [rename.get(x, x) for x in body.regs]
    body.ins = [(op, tuple(rename.get(x, x) if isinstance(x, str) else x for x in a)) for op, a in body.ins]
    for i in range(N):
        body.ins.append(("mov.u32", ("r%d" % i, "w%d" % i)))
        if "r%d" % i not in body.regs:
            body.regs.append("r%d" % i)
    lines = body.ptx(omap)
    sig = "uint32_t (&r)[12], const uint32_t (&a)[12]" + (", const uint32_t (&b)[12]" if n_in == 2 else "")
    out = ["__device__ __forceinline__ void %s(%s)" % (name, sig), "{", "    asm("]
    for ln in lines:
        out.append('        "%s\\n\\t"' % ln)
    outs = ", ".join('"=r"(r[%d])' % i for i in range(N))
    ins = ", ".join('"r"(a[%d])' % i for i in range(N))
    if n_in == 2:
        ins += ", " + ", ".join('"r"(b[%d])' % i for i in range(N))
    out.append("        : " + outs)
    out.append("        : " + ins + ");")
    out.append("}")
    out.append("")
    return out


def main():
    check()
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
    with open(path, "w") as f:
        f.write("\n".join(text))
    print("wrote", path)


if __name__ == "__main__":
    main()
