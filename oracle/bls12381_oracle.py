"""CPU oracle for the crypto12381 hot path.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

A plain-Python (arbitrary precision int) restatement of the algorithms the reference executes through
its bridge (`/root/reference/src/miracl_core_interface.cpp`) into the vendored MIRACL-core
(`/root/reference/3rd-party/miracl-core/`, abbreviated MC/ below).  Each function cites the reference
file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may
import this module; the product (`crypto12381_b200`) never does.

Parity status: PINNED.  `tests/test_oracle_pinned.py` checks every function here against
(a) the compiled, unmodified reference (`oracle/_ref/libref12381.so`, built by `oracle/Makefile`
from the sources under /root/reference) when it is present, and (b) the golden vectors under
`tests/golden/` that were generated from that library by `tools/gen_golden.py`.

Canonical byte formats (identical to include/c12381_cuda.h):
  scalar 32 B BE | G1 affine 96 B x||y | G2 affine 192 B x.b||x.a||y.b||y.a | G1 out 49 B | G2 out 97 B
  | GT 576 B (MC/fp12_BLS12381.cpp:923-929 order).
"""
from __future__ import annotations

# ---------------------------------------------------------------------------------------------
# Constants (MC/rom_field_BLS12381.cpp:51-58, MC/rom_curve_BLS12381.cpp:77-93, 64-bit branch,
# re-assembled from the 58-bit limbs into integers)
# ---------------------------------------------------------------------------------------------
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
X_ABS = 0xD201000000010000  # CURVE_Bnx; the curve parameter is x = -X_ABS (SIGN_OF_X NEGATIVEX)
B_COEFF = 4  # CURVE_B_I
G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G2_XA = 0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8
G2_XB = 0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E
G2_YA = 0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801
G2_YB = 0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE
FRA = 0x1904D3BF02BB0667C231BEB4202C0D1F0FD603FD3CBD5F4F7B2443D784BAB9C4F67EA53D63E7813D8D0775ED92235FB8
FRB = 0x00FC3E2B36C4E03288E9E902231F9FB854A14787B6C7B36FEC0C8EC971F63C5F282D5AC14D6C7EC22CF78A126DDC4AF3
CRU = 0x5F19672FDF76CE51BA69C6076A0F77EADDB3A93BE6F89688DE17D813620A00022E01FFFFFFFEFFFE

# ---------------------------------------------------------------------------------------------
# Fp2 = Fp[i]/(i^2+1)  (MC/fp2_BLS12381.cpp; elements are tuples (a, b) = a + b*i)
# ---------------------------------------------------------------------------------------------
F2_ZERO = (0, 0)
F2_ONE = (1, 0)


def f2_add(x, y):
    return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)


def f2_sub(x, y):
    return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)


def f2_neg(x):
    return ((-x[0]) % P, (-x[1]) % P)


def f2_conj(x):  # MC/fp2_BLS12381.cpp:199
    return (x[0], (-x[1]) % P)


def f2_mul(x, y):  # MC/fp2_BLS12381.cpp:266
    return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)


def f2_sqr(x):  # MC/fp2_BLS12381.cpp:241
    return f2_mul(x, x)


def f2_imul(x, c):
    return (x[0] * c % P, x[1] * c % P)


def f2_pmul(x, s):  # multiply by an Fp element
    return (x[0] * s % P, x[1] * s % P)


def f2_mul_ip(x):  # x * (1+i), QNRI = 0  (MC/fp2_BLS12381.cpp:373-395)
    return ((x[0] - x[1]) % P, (x[0] + x[1]) % P)


def f2_inv(x):  # MC/fp2_BLS12381.cpp:334
    n = pow((x[0] * x[0] + x[1] * x[1]) % P, -1, P)
    return (x[0] * n % P, (-x[1]) * n % P)


def fp_sign(a):  # parity (MC/fp_BLS12381.cpp:912-936, BIG_ENDIAN_SIGN not defined)
    return a & 1


def f2_sign(x):  # MC/fp2_BLS12381.cpp:168-181: parity of real part, of imaginary part if real is zero
    return fp_sign(x[1]) if x[0] == 0 else fp_sign(x[0])


def fp_sqrt(a):
    """Square root mod p (p = 3 mod 4).  Returns None when `a` is a non-residue."""
    y = pow(a, (P + 1) // 4, P)
    return y if y * y % P == a % P else None


def f2_sqrt(x):
    """Square root in Fp2 (value-level restatement of MC/fp2_BLS12381.cpp:460-520)."""
    a, b = x
    if b == 0:
        s = fp_sqrt(a)
        if s is not None:
            return (s, 0)
        s = fp_sqrt((-a) % P)
        return None if s is None else (0, s)
    n = fp_sqrt((a * a + b * b) % P)
    if n is None:
        return None
    half = pow(2, -1, P)
    t = (a + n) * half % P
    s = fp_sqrt(t)
    if s is None:
        t = (a - n) * half % P
        s = fp_sqrt(t)
        if s is None:
            return None
    y = (s, b * pow(2 * s, -1, P) % P)
    return y if f2_sqr(y) == (a % P, b % P) else None


F2_FROB = (FRA, FRB)  # (1+i)^((p-1)/6)

# ---------------------------------------------------------------------------------------------
# Fp4 = Fp2[j]/(j^2-(1+i))  (MC/fp4_BLS12381.cpp; tuples (a, b) = a + b*j)
# ---------------------------------------------------------------------------------------------
F4_ZERO = (F2_ZERO, F2_ZERO)
F4_ONE = (F2_ONE, F2_ZERO)


def f4_add(x, y):
    return (f2_add(x[0], y[0]), f2_add(x[1], y[1]))


def f4_sub(x, y):
    return (f2_sub(x[0], y[0]), f2_sub(x[1], y[1]))


def f4_neg(x):
    return (f2_neg(x[0]), f2_neg(x[1]))


def f4_mul(x, y):  # MC/fp4_BLS12381.cpp:274
    return (f2_add(f2_mul(x[0], y[0]), f2_mul_ip(f2_mul(x[1], y[1]))),
            f2_add(f2_mul(x[0], y[1]), f2_mul(x[1], y[0])))


def f4_times_i(x):  # multiply by j (MC/fp4_BLS12381.cpp:343-357)
    return (f2_mul_ip(x[1]), x[0])


def f4_conj(x):  # MC/fp4_BLS12381.cpp conj: a - b*j
    return (x[0], f2_neg(x[1]))


def f4_pmul(x, s):  # by an Fp2 element
    return (f2_mul(x[0], s), f2_mul(x[1], s))


def f4_inv(x):  # MC/fp4_BLS12381.cpp:326
    t = f2_inv(f2_sub(f2_sqr(x[0]), f2_mul_ip(f2_sqr(x[1]))))
    return (f2_mul(x[0], t), f2_neg(f2_mul(x[1], t)))


def f4_frob(x, f):  # MC/fp4_BLS12381.cpp:359-364
    return (f2_conj(x[0]), f2_mul(f, f2_conj(x[1])))


# ---------------------------------------------------------------------------------------------
# Fp12 = Fp4[k]/(k^3-j)  (MC/fp12_BLS12381.cpp; tuples (a, b, c) = a + b*k + c*k^2)
# ---------------------------------------------------------------------------------------------
F12_ONE = (F4_ONE, F4_ZERO, F4_ZERO)


def f12_mul(x, y):  # value of MC/fp12_BLS12381.cpp:246-298
    a0, a1, a2 = x
    b0, b1, b2 = y
    c0 = f4_add(f4_mul(a0, b0), f4_times_i(f4_add(f4_mul(a1, b2), f4_mul(a2, b1))))
    c1 = f4_add(f4_add(f4_mul(a0, b1), f4_mul(a1, b0)), f4_times_i(f4_mul(a2, b2)))
    c2 = f4_add(f4_add(f4_mul(a0, b2), f4_mul(a2, b0)), f4_mul(a1, b1))
    return (c0, c1, c2)


def f12_sqr(x):  # MC/fp12_BLS12381.cpp:190 (value)
    return f12_mul(x, x)


def f12_conj(x):  # MC/fp12_BLS12381.cpp:117-123: (conj a, -conj b, conj c)
    return (f4_conj(x[0]), f4_neg(f4_conj(x[1])), f4_conj(x[2]))


def f12_inv(x):  # MC/fp12_BLS12381.cpp:627-665
    a, b, c = x
    f0 = f4_sub(f4_mul(a, a), f4_times_i(f4_mul(b, c)))
    f1 = f4_sub(f4_times_i(f4_mul(c, c)), f4_mul(a, b))
    f2 = f4_sub(f4_mul(b, b), f4_mul(a, c))
    f3 = f4_add(f4_add(f4_times_i(f4_mul(b, f2)), f4_mul(a, f0)), f4_times_i(f4_mul(c, f1)))
    f3 = f4_inv(f3)
    return (f4_mul(f0, f3), f4_mul(f1, f3), f4_mul(f2, f3))


def f12_frob(x, f=F2_FROB):  # MC/fp12_BLS12381.cpp:867-881
    f2 = f2_sqr(f)
    f3 = f2_mul(f2, f)
    a = f4_frob(x[0], f3)
    b = f4_pmul(f4_frob(x[1], f3), f)
    c = f4_pmul(f4_frob(x[2], f3), f2)
    return (a, b, c)


def f12_pow(x, e):
    """a^e for unitary a.  MC/fp12_BLS12381.cpp:736-777 walks the 3e/e NAF with cyclotomic squarings and
    conj as inverse; on the cyclotomic subgroup (all GT values) that equals plain exponentiation."""
    result = F12_ONE
    base = x
    while e:
        if e & 1:
            result = f12_mul(result, base)
        base = f12_sqr(base)
        e >>= 1
    return result


def f12_is_unity(x):
    return x == F12_ONE


# ---------------------------------------------------------------------------------------------
# G1: y^2 = x^3 + 4 over Fp.  Points are None (identity) or affine (x, y).
# Group law values only (the reference's projective formulas MC/ecp_BLS12381.cpp:550-588,750-812 are
# complete; outputs are compared on normalised affine encodings, so representation is irrelevant).
# ---------------------------------------------------------------------------------------------
G1_GEN = (G1_X, G1_Y)


def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % P == 0:
            return None
        lam = 3 * p[0] * p[0] * pow(2 * p[1], -1, P) % P
    else:
        lam = (q[1] - p[1]) * pow(q[0] - p[0], -1, P) % P
    x3 = (lam * lam - p[0] - q[0]) % P
    return (x3, (lam * (p[0] - x3) - p[1]) % P)


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % P)


def g1_mul(p, k):
    """k*P.  PAIR_G1mul (MC/pair_BLS12381.cpp:876-924) reduces k mod r and uses GLV; same value on G1."""
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, p)
        p = g1_add(p, p)
        k >>= 1
    return acc


def g1_msm_muln(points, scalars):
    """Restatement of ECP_muln (MC/ecp_BLS12381.cpp:1112-1148): 4-bit unsigned fixed-window Pippenger,
    16 buckets per window (bucket 0 included and ignored by the running sum), windows high to low."""
    acc = None
    if not points:
        return acc
    nb = (max(scalars).bit_length() + 3) // 4
    for i in range(nb - 1, -1, -1):
        buckets = [None] * 16
        for pt, e in zip(points, scalars):
            k = (e >> (4 * i)) & 15
            buckets[k] = g1_add(buckets[k], pt)
        run = total = None
        for j in range(15, 0, -1):
            run = g1_add(run, buckets[j])
            total = g1_add(total, run)
        for _ in range(4):
            acc = g1_add(acc, acc)
        acc = g1_add(acc, total)
    return acc


def g1_msm_live(points, scalars):
    """The LIVE DSL path (g1_point.hpp:389-401): terms paired up through double_multiply (ECP_mul2,
    value = v1*P1 + v2*P2), odd tail through multiply; partials accumulated with add."""
    acc = None
    i = 0
    while i + 1 < len(points):
        acc = g1_add(acc, g1_add(g1_mul(points[i], scalars[i]), g1_mul(points[i + 1], scalars[i + 1])))
        i += 2
    if i < len(points):
        acc = g1_add(acc, g1_mul(points[i], scalars[i]))
    return acc


def g1_on_curve(p):
    return p is None or (p[1] * p[1] - p[0] ** 3 - B_COEFF) % P == 0


# ---------------------------------------------------------------------------------------------
# G2: y^2 = x^3 + 4(1+i) over Fp2 (M-type twist, MC/ecp2_BLS12381.cpp:270-296)
# ---------------------------------------------------------------------------------------------
G2_GEN = ((G2_XA, G2_XB), (G2_YA, G2_YB))
G2_B = f2_mul_ip((B_COEFF, 0))


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if f2_add(p[1], q[1]) == F2_ZERO:
            return None
        lam = f2_mul(f2_imul(f2_sqr(p[0]), 3), f2_inv(f2_imul(p[1], 2)))
    else:
        lam = f2_mul(f2_sub(q[1], p[1]), f2_inv(f2_sub(q[0], p[0])))
    x3 = f2_sub(f2_sub(f2_sqr(lam), p[0]), q[0])
    return (x3, f2_sub(f2_mul(lam, f2_sub(p[0], x3)), p[1]))


def g2_neg(p):
    return None if p is None else (p[0], f2_neg(p[1]))


def g2_mul(p, k):
    """k*Q.  PAIR_G2mul (MC/pair_BLS12381.cpp:927-983) reduces mod r and uses the GS split; same value."""
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, p)
        p = g2_add(p, p)
        k >>= 1
    return acc


def g2_msm(points, scalars):
    """G2 product as the reference evaluates it: eager per-term multiply + add loop (g2_point.hpp:202-236)."""
    acc = None
    for pt, e in zip(points, scalars):
        acc = g2_add(acc, g2_mul(pt, e))
    return acc


def g2_on_curve(p):
    return p is None or f2_sub(f2_sqr(p[1]), f2_add(f2_mul(f2_sqr(p[0]), p[0]), G2_B)) == F2_ZERO


# projective G2 arithmetic EXACTLY as MIRACL does it inside the Miller loop (the un-exponentiated
# Miller value depends on the projective representative, so the formulas are restated verbatim)
def _ecp2_dbl(A):  # MC/ecp2_BLS12381.cpp:358-409 (M_TYPE)
    x, y, z = A
    iy = y
    t0 = f2_sqr(y)
    t1 = f2_mul(iy, z)
    t2 = f2_sqr(z)
    z = f2_imul(t0, 8)
    t2 = f2_mul_ip(f2_imul(t2, 3 * B_COEFF))
    x3 = f2_mul(t2, z)
    y3 = f2_add(t0, t2)
    z = f2_mul(z, t1)
    t2 = f2_imul(t2, 3)
    t0 = f2_sub(t0, t2)
    y3 = f2_mul(y3, t0)
    y = f2_add(y3, x3)
    t1 = f2_mul(x, iy)
    x = f2_imul(f2_mul(t0, t1), 2)
    return (x, y, z)


def _ecp2_add(A, Bp):  # MC/ecp2_BLS12381.cpp:413-502 (M_TYPE)
    b3 = 3 * B_COEFF
    x1, y1, z1 = A
    x2, y2, z2 = Bp
    t0 = f2_mul(x1, x2)
    t1 = f2_mul(y1, y2)
    t2 = f2_mul(z1, z2)
    t3 = f2_sub(f2_mul(f2_add(x1, y1), f2_add(x2, y2)), f2_add(t0, t1))
    t4 = f2_sub(f2_mul(f2_add(y1, z1), f2_add(y2, z2)), f2_add(t1, t2))
    y3 = f2_sub(f2_mul(f2_add(x1, z1), f2_add(x2, z2)), f2_add(t0, t2))
    t0 = f2_imul(t0, 3)
    t2 = f2_mul_ip(f2_imul(t2, b3))
    z3 = f2_add(t1, t2)
    t1 = f2_sub(t1, t2)
    y3 = f2_mul_ip(f2_imul(y3, b3))
    x3 = f2_mul(y3, t4)
    t2 = f2_mul(t3, t1)
    xo = f2_sub(t2, x3)
    y3 = f2_mul(y3, t0)
    t1 = f2_mul(t1, z3)
    yo = f2_add(y3, t1)
    t0 = f2_mul(t0, t3)
    z3 = f2_mul(z3, t4)
    zo = f2_add(z3, t0)
    return (xo, yo, zo)


# ---------------------------------------------------------------------------------------------
# Pairing (MC/pair_BLS12381.cpp)
# ---------------------------------------------------------------------------------------------
def _pair_double(A):  # MC/pair_BLS12381.cpp:40-78 -> (A', AA, BB, CC)
    x, y, z = A
    CC = x
    YY = y
    BB = z
    AA = f2_mul(YY, BB)
    CC = f2_sqr(CC)
    YY = f2_sqr(YY)
    BB = f2_sqr(BB)
    AA = f2_mul_ip(f2_neg(f2_add(AA, AA)))  # -2YZ * (1+i)
    BB = f2_mul_ip(f2_imul(BB, 3 * B_COEFF))  # 3b Z^2 * (1+i)   (M_TYPE)
    CC = f2_imul(CC, 3)  # 3X^2
    BB = f2_sub(BB, YY)
    return _ecp2_dbl(A), AA, BB, CC


def _pair_add(A, Bq):  # MC/pair_BLS12381.cpp:81-116; Bq affine (x, y) with implicit z = 1
    x1, y1, z1 = A
    x2, y2 = Bq
    T1 = f2_mul(z1, y2)
    BB = f2_mul(z1, x2)
    AA = f2_sub(x1, BB)
    CC = f2_sub(y1, T1)
    T1 = AA
    AA = f2_mul_ip(AA)  # M_TYPE
    T1 = f2_mul(T1, y2)
    BB = f2_sub(f2_mul(CC, x2), T1)
    CC = f2_neg(CC)
    return _ecp2_add(A, (x2, y2, F2_ONE)), AA, BB, CC


def _line_to_f12(AA, BB, CC, Qx, Qy):  # MC/pair_BLS12381.cpp:119-144 (M_TYPE): a=[AA*Qy, BB], b=0, c=[0, CC*Qx]
    return ((f2_pmul(AA, Qy), BB), F4_ZERO, (F2_ZERO, f2_pmul(CC, Qx)))


def _ate_bits():  # MC/pair_BLS12381.cpp:147-169: n = |x|, n3 = 3n, loop i = nbits(n3)-2 .. 1
    n = X_ABS
    n3 = 3 * n
    return [((n3 >> i) & 1) - ((n >> i) & 1) for i in range(n3.bit_length() - 2, 0, -1)]


ATE_DIGITS = _ate_bits()


def miller_loop(pairs):
    """Product of Miller loops over `pairs` = [(P in G1 affine, Q in G2 affine), ...] with shared squarings.

    For one pair this is PAIR_ate (MC/pair_BLS12381.cpp:425-505), for two PAIR_double_ate (:508-626);
    for k pairs it is the same loop body repeated per pair, which as a VALUE equals the product of the
    single Miller values (what the DSL builds with multiply(fp12&, fp12&), liner_pair.hpp:219-230).
    Pairs whose G1 point is the identity contribute 1 (:449, :532-541); a G2 identity falls through the
    complete formulas and contributes a value that final exponentiation maps to 1.  Not exponentiated.
    """
    live = [(p, q) for (p, q) in pairs if p is not None]
    r = F12_ONE
    if not live:
        return r
    st = []
    for (p, q) in live:
        if q is None:
            st.append([(F2_ZERO, F2_ONE, F2_ZERO), None, p])
        else:
            st.append([(q[0], q[1], F2_ONE), q, p])
    for bt in ATE_DIGITS:
        r = f12_sqr(r)
        for s in st:
            A, AA, BB, CC = _pair_double(s[0])
            s[0] = A
            r = f12_mul(r, _line_to_f12(AA, BB, CC, s[2][0], s[2][1]))
        if bt:
            for s in st:
                if s[1] is None:
                    # ECP2_affine of the identity leaves (0,1,0); PAIR_add with it is what MIRACL executes
                    qq = (F2_ZERO, F2_ONE) if bt == 1 else (F2_ZERO, f2_neg(F2_ONE))
                    A, AA, BB, CC = _pair_add_proj(s[0], (qq[0], qq[1], F2_ZERO))
                else:
                    qq = s[1] if bt == 1 else (s[1][0], f2_neg(s[1][1]))
                    A, AA, BB, CC = _pair_add(s[0], qq)
                s[0] = A
                r = f12_mul(r, _line_to_f12(AA, BB, CC, s[2][0], s[2][1]))
    return f12_conj(r)  # x < 0 (:485-487)


def _pair_add_proj(A, Bp):
    """PAIR_add when B is the (non-affine) identity (0:1:0): same statements, B->x, B->y read as stored."""
    x1, y1, z1 = A
    x2, y2, _ = Bp
    T1 = f2_mul(z1, y2)
    BB = f2_mul(z1, x2)
    AA = f2_sub(x1, BB)
    CC = f2_sub(y1, T1)
    T1 = AA
    AA = f2_mul_ip(AA)
    T1 = f2_mul(T1, y2)
    BB = f2_sub(f2_mul(CC, x2), T1)
    CC = f2_neg(CC)
    return _ecp2_add(A, Bp), AA, BB, CC


def final_exp(f):
    """PAIR_fexp (MC/pair_BLS12381.cpp:629-755): easy part, then the eprint 2020/875 hard part times f^3,
    i.e. overall exponent 3*(p^12-1)/r."""
    t0 = f12_inv(f)
    r = f12_mul(f12_conj(f), t0)
    t0 = r
    r = f12_mul(f12_frob(f12_frob(r)), t0)

    def pow_x(a):  # a^x with x negative: pow by |x| then conjugate
        return f12_conj(f12_pow(a, X_ABS))

    y1 = f12_mul(f12_sqr(r), r)
    r = f12_mul(pow_x(r), f12_conj(r))
    r = f12_mul(pow_x(r), f12_conj(r))
    r = f12_mul(pow_x(r), f12_frob(r))
    # y0 = r^(x^2): MIRACL applies FP12_pow by |x| twice WITHOUT conjugating (:741-742); x^2 = |x|^2
    y0 = f12_pow(f12_pow(r, X_ABS), X_ABS)
    y0 = f12_mul(y0, f12_frob(f12_frob(r)))
    r = f12_mul(y0, f12_conj(r))
    return f12_mul(r, y1)


def pairing(p, q):
    return final_exp(miller_loop([(p, q)]))


def pairing_product(pairs):
    return final_exp(miller_loop(pairs))


# ---------------------------------------------------------------------------------------------
# Wire formats
# ---------------------------------------------------------------------------------------------
def fp_to_bytes(a):
    return int(a % P).to_bytes(48, "big")


def scalar_to_bytes(k):
    return int(k).to_bytes(32, "big")


def scalar_from_bytes(b):
    return int.from_bytes(b, "big")


def g1_to_affine_bytes(p):
    return bytes(96) if p is None else fp_to_bytes(p[0]) + fp_to_bytes(p[1])


def g1_from_affine_bytes(b):
    if b == bytes(96):
        return None
    return (int.from_bytes(b[:48], "big"), int.from_bytes(b[48:], "big"))


def g1_compress(p):  # ECP_toOctet compressed (MC/ecp_BLS12381.cpp:445-491); identity per g1_point.hpp:113-117
    if p is None:
        return bytes(49)
    return bytes([0x02 | fp_sign(p[1])]) + fp_to_bytes(p[0])


def g1_decompress(b):  # ECP_fromOctet / ECP_setx (MC/ecp_BLS12381.cpp:495-545,302-323)
    if b == bytes(49):
        return None
    x = int.from_bytes(b[1:49], "big")
    if b[0] not in (2, 3) or x >= P:
        raise ValueError("bad G1 encoding")
    y = fp_sqrt((x ** 3 + B_COEFF) % P)
    if y is None:
        raise ValueError("x not on curve")
    if fp_sign(y) != (b[0] & 1):
        y = (-y) % P
    return (x, y)


def f2_to_bytes(x):  # FP2_toBytes: b (imaginary) first, then a (MC/fp2_BLS12381.cpp:83-87)
    return fp_to_bytes(x[1]) + fp_to_bytes(x[0])


def f2_from_bytes(b):
    return (int.from_bytes(b[48:96], "big"), int.from_bytes(b[:48], "big"))


def g2_to_affine_bytes(p):
    return bytes(192) if p is None else f2_to_bytes(p[0]) + f2_to_bytes(p[1])


def g2_from_affine_bytes(b):
    if b == bytes(192):
        return None
    return (f2_from_bytes(b[:96]), f2_from_bytes(b[96:]))


def g2_compress(p):  # ECP2_toOctet compressed (MC/ecp2_BLS12381.cpp:184-222); identity per g2_point.hpp:97-101
    if p is None:
        return bytes(97)
    return bytes([0x02 | f2_sign(p[1])]) + f2_to_bytes(p[0])


def g2_decompress(b):  # ECP2_fromOctet / ECP2_setx (MC/ecp2_BLS12381.cpp:225-266,322-344)
    if b == bytes(97):
        return None
    x = f2_from_bytes(b[1:97])
    if b[0] not in (2, 3):
        raise ValueError("bad G2 encoding")
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), G2_B))
    if y is None:
        raise ValueError("x not on curve")
    if f2_sign(y) != (b[0] & 1):
        y = f2_neg(y)
    return (x, y)


def f4_to_bytes(x):  # FP4_toBytes: b then a (MC/fp4_BLS12381.cpp:61-65)
    return f2_to_bytes(x[1]) + f2_to_bytes(x[0])


def f4_from_bytes(b):
    return (f2_from_bytes(b[96:192]), f2_from_bytes(b[:96]))


def gt_to_bytes(f):  # FP12_toOctet: c, b, a (MC/fp12_BLS12381.cpp:923-929)
    return f4_to_bytes(f[2]) + f4_to_bytes(f[1]) + f4_to_bytes(f[0])


def gt_from_bytes(b):
    return (f4_from_bytes(b[384:576]), f4_from_bytes(b[192:384]), f4_from_bytes(b[:192]))


# ---------------------------------------------------------------------------------------------
# Hash to G1: G1Point::from_hash (include/crypto12381/g1_point.hpp:219-234)
#   SHA3-512 digest -> big2 -> fixed_time_mod p -> ECP_map2point (MC/ecp_BLS12381.cpp:1276,1493-1627: simplified SWU on
#   the isogenous curve E': y^2 = x^3 + A'x + B', Z = 11, then the 11-isogeny to E, projective) -> ECP_cfp (:1252-1273,
#   times CURVE_Cof = 1 - x).  Value-level restatement: the reference's constant-time selections and its shared
#   inversion/square-root exponentiation compute exactly the RFC 9380 map with sgn0 = parity (FP_sign).
# ---------------------------------------------------------------------------------------------
from .iso11_g1 import ISO_A, ISO_B, SSWU_Z, H_EFF, ISO_XNUM, ISO_XDEN, ISO_YNUM, ISO_YDEN  # noqa: E402


def sswu_iso_curve(u):
    """Point (x, y) on E' for the field element u (MC/ecp_BLS12381.cpp:1509-1566)."""
    u %= P
    zu2 = SSWU_Z * u * u % P
    tv1 = (zu2 * zu2 + zu2) % P
    if tv1 == 0:
        # u = 0 or Z u^2 = -1: the reference has no exceptional branch; its shared denominator A'(Z^2u^4 + Zu^2) is zero,
        # FP_inv(0) = 0 zeroes the projective Z and the result is the identity (checked against the compiled reference)
        return None
    x1 = (-ISO_B) * pow(ISO_A, -1, P) % P * (1 + pow(tv1, -1, P)) % P
    gx1 = (x1 * x1 * x1 + ISO_A * x1 + ISO_B) % P
    y = fp_sqrt(gx1)
    x = x1
    if y is None:
        x = zu2 * x1 % P
        y = fp_sqrt((x * x * x + ISO_A * x + ISO_B) % P)
        assert y is not None
    if fp_sign(y) != fp_sign(u):
        y = (-y) % P
    return (x, y)


def _horner(coeffs, x, monic=False):
    acc = 1 if monic else 0
    for c in coeffs:
        acc = (acc * x + c) % P
    return acc


def iso11_map(pt):
    """The 11-isogeny E' -> E (MC/ecp_BLS12381.cpp:1568-1627), affine value; None if a denominator vanishes."""
    if pt is None:
        return None
    x, y = pt
    xn, xd = _horner(ISO_XNUM, x), _horner(ISO_XDEN, x, monic=True)
    yn, yd = _horner(ISO_YNUM, x), _horner(ISO_YDEN, x, monic=True)
    if xd == 0 or yd == 0:
        return None
    return (xn * pow(xd, -1, P) % P, y * yn % P * pow(yd, -1, P) % P)


def g1_mul_raw(p, k):
    """k*P without reducing k mod r (the point need not have order r)."""
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, p)
        p = g1_add(p, p)
        k >>= 1
    return acc


def map_to_g1(u):
    """map_to_point + multiply_cofactor (src/miracl_core_interface.cpp:154-162)."""
    return g1_mul_raw(iso11_map(sswu_iso_curve(u)), H_EFF)


def hash_to_g1(msg: bytes):
    """G1Point::from_hash over the bytes a hash_state absorbed (g1_point.hpp:219-234)."""
    import hashlib
    return map_to_g1(int.from_bytes(hashlib.sha3_512(msg).digest(), "big") % P)
