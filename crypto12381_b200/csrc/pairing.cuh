// Fp4 / Fp12 tower, optimal-ate Miller loop and final exponentiation for BLS12-381, host+device bodies.
//
// Tower (MIRACL's 2-2-3): Fp4 = Fp2[j]/(j^2 - (1+i)), Fp12 = Fp4[k]/(k^3 - j); values are kept in
// exactly this basis so that the 576-byte GT wire format (fp12_BLS12381.cpp:923-929) is a plain dump.
// Replaces FP4_*/FP12_* (3rd-party/miracl-core/fp4_BLS12381.cpp:243-364, fp12_BLS12381.cpp:117-298,627-881),
// PAIR_ate / PAIR_double_ate / PAIR_fexp (pair_BLS12381.cpp:425-755) behind the bridge entries
// pair_ate / pair_double_ate / pair_final_exponentiation / multiply(fp12&) / pow(fp12&)
// (src/miracl_core_interface.cpp:251-289).
//
// The Miller loop follows MIRACL's line functions statement by statement (PAIR_double :40-78, PAIR_add :81-116,
// PAIR_line :119-144 with the M-type sparse form) on the complete projective G2 formulas, so even the
// un-exponentiated Miller value is bit-identical; the final exponentiation uses the same exponent
// 3 (p^12 - 1) / r (eprint 2020/875 hard part times f^3).
#pragma once
#include "scalar_mul.cuh"

// -DC12_PAIR_LOCKSTEP (device, thread-per-instance kernels): a block-wide barrier at the top of every Miller-loop iteration and
// of every exponentiation step.  Every thread runs the same instruction stream (the loop bits are public constants), so the
// barrier only keeps the warps of a block at the same place in the ~100 KB of straight-line code, where they share
// instruction-cache lines instead of evicting each other's (ncu: no_instruction stalls, sm__icc hit rate; profiles/).
// Kernels built with it must bring EVERY thread of a block through the bodies (no early return).
#if defined(C12_PAIR_LOCKSTEP) && defined(__CUDA_ARCH__)
#define C12_BLOCK_ALIGN() __syncthreads()
#else
#define C12_BLOCK_ALIGN() ((void)0)
#endif

namespace c12 {

struct Fp4 {
    Fp2 a, b;  // a + b j
};
struct Fp12 {
    Fp4 a, b, c;  // a + b k + c k^2
};

// ---- Fp4 ------------------------------------------------------------------------------------------------
C12_HD Fp4 fp4_zero() { return Fp4{fp2_zero(), fp2_zero()}; }
C12_HD Fp4 fp4_one() { return Fp4{fp2_one(), fp2_zero()}; }
C12_HD Fp4 add(const Fp4& x, const Fp4& y) { return Fp4{add(x.a, y.a), add(x.b, y.b)}; }
C12_HD Fp4 sub(const Fp4& x, const Fp4& y) { return Fp4{sub(x.a, y.a), sub(x.b, y.b)}; }
C12_HD Fp4 neg(const Fp4& x) { return Fp4{neg(x.a), neg(x.b)}; }
C12_HD Fp4 dbl(const Fp4& x) { return Fp4{dbl(x.a), dbl(x.b)}; }
C12_HD Fp4 conj(const Fp4& x) { return Fp4{x.a, neg(x.b)}; }   // FP4_conj
C12_HD Fp4 nconj(const Fp4& x) { return Fp4{neg(x.a), x.b}; }  // FP4_nconj = -conj
C12_HD bool eq(const Fp4& x, const Fp4& y) { return eq(x.a, y.a) && eq(x.b, y.b); }
C12_HD bool is_zero(const Fp4& x) { return is_zero(x.a) && is_zero(x.b); }
C12_HD Fp4 times_j(const Fp4& x) { return Fp4{mul_ip(x.b), x.a}; }  // FP4_times_i (fp4_BLS12381.cpp:343-357)

C12_HD Fp4 mul(const Fp4& x, const Fp4& y)  // Karatsuba, 3 Fp2 products (FP4_mul, fp4_BLS12381.cpp:274)
{
    Fp2 t0 = mul(x.a, y.a);
    Fp2 t1 = mul(x.b, y.b);
    Fp2 t2 = mul(add(x.a, x.b), add(y.a, y.b));
    return Fp4{add(t0, mul_ip(t1)), sub(sub(t2, t0), t1)};
}
C12_HD Fp4 sqr(const Fp4& x)  // 2 Fp2 products (FP4_sqr, fp4_BLS12381.cpp:243)
{
    Fp2 t0 = mul(x.a, x.b);
    Fp2 t1 = mul(add(x.a, x.b), add(x.a, mul_ip(x.b)));  // a^2 + xi b^2 + ab + xi ab
    return Fp4{sub(sub(t1, t0), mul_ip(t0)), dbl(t0)};
}
C12_HD Fp4 mul_fp2(const Fp4& x, const Fp2& s) { return Fp4{mul(x.a, s), mul(x.b, s)}; }  // FP4_pmul
// x * (s j): product with a "high-half only" Fp4, 2 Fp2 products
C12_HD Fp4 mul_fp2_j(const Fp4& x, const Fp2& s) { return Fp4{mul_ip(mul(x.b, s)), mul(x.a, s)}; }
C12_HD Fp4 inv(const Fp4& x)  // FP4_inv (fp4_BLS12381.cpp:326)
{
    Fp2 t = inv(sub(sqr(x.a), mul_ip(sqr(x.b))));
    return Fp4{mul(x.a, t), neg(mul(x.b, t))};
}
C12_HD Fp4 frob(const Fp4& x, const Fp2& f3)  // FP4_frob (fp4_BLS12381.cpp:359-364)
{
    return Fp4{conj(x.a), mul(f3, conj(x.b))};
}

// ---- Fp12 -----------------------------------------------------------------------------------------------
C12_HD Fp12 fp12_one() { return Fp12{fp4_one(), fp4_zero(), fp4_zero()}; }
C12_HD bool eq(const Fp12& x, const Fp12& y) { return eq(x.a, y.a) && eq(x.b, y.b) && eq(x.c, y.c); }
C12_HD Fp12 conj(const Fp12& x) { return Fp12{conj(x.a), nconj(x.b), conj(x.c)}; }  // FP12_conj (fp12:117-123)

// Karatsuba over Fp4: 6 Fp4 products = 54 Fp products (FP12_mul, fp12_BLS12381.cpp:246-298)
C12_HD_NOINLINE Fp12 mul(const Fp12& x, const Fp12& y)
{
    Fp4 v0 = mul(x.a, y.a);
    Fp4 v1 = mul(x.b, y.b);
    Fp4 v2 = mul(x.c, y.c);
    Fp4 t0 = sub(sub(mul(add(x.b, x.c), add(y.b, y.c)), v1), v2);
    Fp4 t1 = sub(sub(mul(add(x.a, x.b), add(y.a, y.b)), v0), v1);
    Fp4 t2 = sub(sub(mul(add(x.a, x.c), add(y.a, y.c)), v0), v2);
    return Fp12{add(v0, times_j(t0)), add(t1, times_j(v2)), add(t2, v1)};
}

// Chung–Hasan SQR2: 2 Fp4 products + 3 Fp4 squarings = 36 Fp products (FP12_sqr, fp12_BLS12381.cpp:190)
C12_HD_NOINLINE Fp12 sqr(const Fp12& x)
{
    Fp4 s0 = sqr(x.a);
    Fp4 s1 = dbl(mul(x.a, x.b));
    Fp4 s2 = sqr(add(sub(x.a, x.b), x.c));
    Fp4 s3 = dbl(mul(x.b, x.c));
    Fp4 s4 = sqr(x.c);
    return Fp12{add(s0, times_j(s3)), add(s1, times_j(s4)), sub(sub(add(add(s1, s2), s3), s0), s4)};
}

// Granger–Scott squaring, valid on the cyclotomic subgroup: 3 Fp4 squarings = 18 Fp products
// (FP12_usqr, fp12_BLS12381.cpp:147-187)
C12_HD_NOINLINE Fp12 usqr(const Fp12& x)
{
    Fp4 A = sqr(x.a);
    Fp4 B = times_j(sqr(x.c));
    Fp4 C = sqr(x.b);
    Fp12 r;
    r.a = add(add(dbl(A), A), dbl(nconj(x.a)));
    r.b = add(add(dbl(B), B), dbl(conj(x.b)));
    r.c = add(add(dbl(C), C), dbl(nconj(x.c)));
    return r;
}

// f * line, line = l0 + l1 j + (l2 j) k^2 (PAIR_line's M-type sparse form, pair_BLS12381.cpp:119-144).
// 15 Fp2 products.  Same value as FP12_ssmul / FP12_smul (fp12_BLS12381.cpp:304-620).
C12_HD_NOINLINE Fp12 mul_line(const Fp12& f, const Fp2& l0, const Fp2& l1, const Fp2& l2)
{
    Fp4 la = Fp4{l0, l1};
    Fp12 r;
    r.a = add(mul(f.a, la), times_j(mul_fp2_j(f.b, l2)));
    r.b = add(mul(f.b, la), times_j(mul_fp2_j(f.c, l2)));
    r.c = add(mul_fp2_j(f.a, l2), mul(f.c, la));
    return r;
}

C12_HD_NOINLINE Fp12 inv(const Fp12& x)  // FP12_inv (fp12_BLS12381.cpp:627-665)
{
    Fp4 f0 = sub(sqr(x.a), times_j(mul(x.b, x.c)));
    Fp4 f1 = sub(times_j(sqr(x.c)), mul(x.a, x.b));
    Fp4 f2 = sub(sqr(x.b), mul(x.a, x.c));
    Fp4 f3 = add(add(times_j(mul(x.b, f2)), mul(x.a, f0)), times_j(mul(x.c, f1)));
    f3 = inv(f3);
    return Fp12{mul(f0, f3), mul(f1, f3), mul(f2, f3)};
}

// x^p (FP12_frob, fp12_BLS12381.cpp:867-881) with f = (1+i)^((p-1)/6): f, f^2, f^3 precomputed
C12_HD_NOINLINE Fp12 frob(const Fp12& x)
{
    Fp2 f1 = frob_c1_m(), f2 = frob_c2_m(), f3 = frob_c3_m();
    return Fp12{frob(x.a, f3), mul_fp2(frob(x.b, f3), f1), mul_fp2(frob(x.c, f3), f2)};
}

// x^|x_curve| for unitary x: square-and-multiply over the 64-bit |x| (6 set bits) with Granger–Scott
// squarings.  Value of FP12_pow(., |x|) (fp12_BLS12381.cpp:736-777, which walks the 3x/x NAF instead).
C12_HD_NOINLINE Fp12 pow_x_abs(const Fp12& x)
{
    const uint64_t e = C12_X_ABS;
    Fp12 r = x;
#pragma unroll 1
    for (int i = 62; i >= 0; --i) {
        C12_BLOCK_ALIGN();
        r = usqr(r);
        if ((e >> i) & 1ull) r = mul(r, x);
    }
    return r;
}

// PAIR_fexp (pair_BLS12381.cpp:629-755)
C12_HD_NOINLINE Fp12 final_exp(const Fp12& f)
{
    // easy part: f^((p^6-1)(p^2+1))
    Fp12 t0 = inv(f);
    Fp12 r = mul(conj(f), t0);
    t0 = r;
    r = mul(frob(frob(r)), t0);
    // hard part (eprint 2020/875), times r^3
    Fp12 y1 = mul(usqr(r), r);
    r = mul(conj(pow_x_abs(r)), conj(r));   // r^(x-1), x < 0
    r = mul(conj(pow_x_abs(r)), conj(r));   // ^(x-1) again
    r = mul(conj(pow_x_abs(r)), frob(r));   // ^(x+p)
    Fp12 y0 = pow_x_abs(pow_x_abs(r));      // ^(x^2)  (|x|^2 = x^2: no conjugations, :741-742)
    y0 = mul(y0, frob(frob(r)));
    r = mul(y0, conj(r));                   // ^(x^2 + p^2 - 1)
    return mul(r, y1);
}

// a^k for unitary a and a 256-bit scalar (FP12_pow, fp12_BLS12381.cpp:736-777; GTPoint::operator^)
C12_HD_NOINLINE Fp12 gt_pow(const Fp12& a, const Scalar256& k)
{
    Fp12 r = fp12_one();
    bool started = false;
#pragma unroll 1
    for (int i = 255; i >= 0; --i) {
        if (started) r = usqr(r);
        if ((k.v[i >> 5] >> (i & 31)) & 1u) {
            r = started ? mul(r, a) : a;
            started = true;
        }
    }
    return r;
}

// a^k for a in GT (order r) through the Galbraith-Scott split: value of PAIR_GTpow with USE_GS_GT
// (pair_BLS12381.cpp:985-1026; unbridged in crypto12381).  On GT the Frobenius is the power p = x = -z (mod r),
// so with k = sum k_i z^i (gls_split: |k_i| < 2^63) a^k = prod g_i^|k_i|, g_i = frob^i(a), conjugated (= inverted: a is unitary)
// when i is odd or k_i is negative - one or the other.  Joint square-and-multiply over the four 63-bit exponents with the 15
// subset products: 62 Granger-Scott squarings + <= 63 + 11 products against 254 + ~128 for the plain ladder.
C12_HD_NOINLINE Fp12 gt_pow_gs(const Fp12& a, const Scalar256& k)
{
    ScalarParts parts;
    gls_split(k, parts);
    Fp12 tab[16];
    tab[0] = fp12_one();
    Fp12 g = a;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        if (i) g = frob(g);
        const bool inv = ((i & 1) != 0) != (parts.neg[i] != 0);
        const Fp12 gi = inv ? conj(g) : g;
        const int top = 1 << i;
        tab[top] = gi;
#pragma unroll 1
        for (int m = 1; m < top; ++m) tab[top + m] = mul(tab[m], gi);
    }
    Fp12 r = fp12_one();
    bool started = false;
#pragma unroll 1
    for (int bit = 63; bit >= 0; --bit) {
        if (started) r = usqr(r);
        int m = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) m |= (int)((parts.mag[i][bit >> 5] >> (bit & 31)) & 1u) << i;
        if (m) {
            r = started ? mul(r, tab[m]) : tab[m];
            started = true;
        }
    }
    return r;
}

// ---- Miller loop ----------------------------------------------------------------------------------------
struct LineCoeffs {
    Fp2 aa, bb, cc;
};

// PAIR_double (pair_BLS12381.cpp:40-78): line through A,A then A <- 2A (ECP2_dbl, ecp2_BLS12381.cpp:358-409).
// The reference computes Y^2, YZ and Z^2 once for the line and once more inside the doubling; the values are the same
// field elements, so they are computed once here.
C12_HD LineCoeffs pair_double(Proj<Fp2>& A)
{
    LineCoeffs l;
    Fp2 yy = sqr(A.y);
    Fp2 yz = mul(A.y, A.z);
    Fp2 t2 = FieldOps<Fp2>::mul_b3(sqr(A.z));   // 3b' Z^2 = 12 (1+i) Z^2
    l.aa = mul_ip(neg(dbl(yz)));                // -2YZ (1+i)
    l.bb = sub(t2, yy);                         // 3b' Z^2 - Y^2
    l.cc = mul3(sqr(A.x));                      // 3X^2
    // RCB15 Algorithm 9 on the shared products (proj_dbl in ec.cuh)
    Fp2 xy = mul(A.x, A.y);
    Fp2 z3 = mul8(yy);
    Fp2 x3 = mul(t2, z3);
    Fp2 y3 = add(yy, t2);
    Fp2 t0 = sub(yy, mul3(t2));
    A.z = mul(z3, yz);
    A.y = add(mul(y3, t0), x3);
    A.x = dbl(mul(t0, xy));
    return l;
}

// PAIR_add (pair_BLS12381.cpp:81-116): line through A,B then A <- A + B (ECP2_add); B = (bx : by : bz)
C12_HD LineCoeffs pair_add(Proj<Fp2>& A, const Proj<Fp2>& B)
{
    LineCoeffs l;
    Fp2 t1 = mul(A.z, B.y);
    Fp2 t2 = mul(A.z, B.x);
    Fp2 x1 = sub(A.x, t2);
    Fp2 y1 = sub(A.y, t1);
    l.aa = mul_ip(x1);
    l.bb = sub(mul(y1, B.x), mul(x1, B.y));
    l.cc = neg(y1);
    A = proj_add(A, B);
    return l;
}

#define C12_MAX_PAIRS 8

// Product of k <= C12_MAX_PAIRS Miller loops with shared squarings; un-exponentiated, conjugated (x < 0).
// Value-identical to PAIR_ate (k = 1), PAIR_double_ate (k = 2) and to products of those combined with FP12_mul.
// Pairs whose G1 point is the identity contribute 1 (pair_BLS12381.cpp:449, 532-541).
C12_HD_NOINLINE Fp12 miller_loop(const Affine<Fp>* P, const Affine<Fp2>* Q, uint32_t k)
{
    Proj<Fp2> A[C12_MAX_PAIRS], B[C12_MAX_PAIRS];
    bool live[C12_MAX_PAIRS];
#pragma unroll 1
    for (uint32_t j = 0; j < k; ++j) {
        live[j] = !affine_is_inf(P[j]);
#if defined(C12_EXP_LAZY_ADDS)
        live[j] = true;     // timing experiment (fp.cuh): the garbage values must not shorten the loop
#endif
        B[j] = proj_from_affine(Q[j]);  // identity stays (0:1:0), as ECP2_affine leaves it
        A[j] = B[j];
    }
    Fp12 r = fp12_one();
    // digit_i = bit_i(3|x|) - bit_i(|x|) for i = 64 .. 1 (pair_BLS12381.cpp:147-169,466-483); bit i-1 of the masks
    const uint64_t pos = 0x1201000000010000ull, negm = 0x4000000000000000ull;
#pragma unroll 1
    for (int i = 64; i >= 1; --i) {
        C12_BLOCK_ALIGN();
        r = sqr(r);
#pragma unroll 1
        for (uint32_t j = 0; j < k; ++j) {
            C12_BLOCK_ALIGN();      // k is uniform over the block; the live test below is not
            if (!live[j]) continue;
            LineCoeffs l = pair_double(A[j]);
            r = mul_line(r, mul_fp(l.aa, P[j].y), l.bb, mul_fp(l.cc, P[j].x));
        }
        int bt = (int)((pos >> (i - 1)) & 1ull) - (int)((negm >> (i - 1)) & 1ull);
        if (bt != 0) {
#pragma unroll 1
            for (uint32_t j = 0; j < k; ++j) {
                C12_BLOCK_ALIGN();
                if (!live[j]) continue;
                Proj<Fp2> T = B[j];
                if (bt < 0) T.y = neg(T.y);
                LineCoeffs l = pair_add(A[j], T);
                r = mul_line(r, mul_fp(l.aa, P[j].y), l.bb, mul_fp(l.cc, P[j].x));
            }
        }
    }
    return conj(r);
}

// ---- GT wire format (FP12_toOctet, fp12_BLS12381.cpp:923-929: c, b, a; each Fp4 b then a; each Fp2 b then a)
C12_HD void fp2_to_bytes(uint8_t* o, const Fp2& x)
{
    fp_to_be48(o, fp_from_mont(x.b));
    fp_to_be48(o + 48, fp_from_mont(x.a));
}
C12_HD Fp2 fp2_from_bytes(const uint8_t* b)
{
    Fp2 x;
    x.b = fp_to_mont(fp_from_be48(b));
    x.a = fp_to_mont(fp_from_be48(b + 48));
    return x;
}
C12_HD void fp4_to_bytes(uint8_t* o, const Fp4& x)
{
    fp2_to_bytes(o, x.b);
    fp2_to_bytes(o + 96, x.a);
}
C12_HD Fp4 fp4_from_bytes(const uint8_t* b) { return Fp4{fp2_from_bytes(b + 96), fp2_from_bytes(b)}; }
C12_HD void fp12_to_bytes(uint8_t* o, const Fp12& x)
{
    fp4_to_bytes(o, x.c);
    fp4_to_bytes(o + 192, x.b);
    fp4_to_bytes(o + 384, x.a);
}
C12_HD Fp12 fp12_from_bytes(const uint8_t* b)
{
    return Fp12{fp4_from_bytes(b + 384), fp4_from_bytes(b + 192), fp4_from_bytes(b)};
}

// ---- per-instance bodies --------------------------------------------------------------------------------
// mode 0: raw Miller product; mode 1: final-exponentiated GT value
C12_HD bool pairing_product_body(const uint8_t* g1, const uint8_t* g2, uint32_t k, int mode, uint8_t* out576)
{
    Affine<Fp> P[C12_MAX_PAIRS];
    Affine<Fp2> Q[C12_MAX_PAIRS];
    bool ok = k <= C12_MAX_PAIRS;
    if (!ok) k = 0;
#pragma unroll 1
    for (uint32_t j = 0; j < k; ++j) {
        ok = g1_from_bytes96(P[j], g1 + 96 * j) && ok;
        ok = g2_from_bytes192(Q[j], g2 + 192 * j) && ok;
    }
    Fp12 f = miller_loop(P, Q, k);
    if (mode == 1) f = final_exp(f);
    fp12_to_bytes(out576, f);
    return ok;
}
C12_HD void final_exp_body(const uint8_t* in576, uint8_t* out576)
{
    fp12_to_bytes(out576, final_exp(fp12_from_bytes(in576)));
}
C12_HD void gt_mul_body(const uint8_t* a, const uint8_t* b, uint8_t* out576)
{
    fp12_to_bytes(out576, mul(fp12_from_bytes(a), fp12_from_bytes(b)));
}
C12_HD void gt_pow_body(const uint8_t* a, const uint8_t* s32, uint8_t* out576)
{
    fp12_to_bytes(out576, gt_pow(fp12_from_bytes(a), scalar_from_be32(s32)));
}
C12_HD void gt_pow_gs_body(const uint8_t* a, const uint8_t* s32, uint8_t* out576)
{
    fp12_to_bytes(out576, gt_pow_gs(fp12_from_bytes(a), scalar_from_be32(s32)));
}

} // namespace c12
