// Fp2 = Fp[i]/(i^2+1) on top of fp.cuh.  Replaces MIRACL's FP2_* for the device path
// (reference: 3rd-party/miracl-core/fp2_BLS12381.cpp:199 conj, :241 sqr, :266 mul, :334 inv, :373 mul_ip).
// Also the overload set (add/sub/mul/sqr/...) the curve templates in ec.cuh are written against, for both
// Fp (G1) and Fp2 (G2).
#pragma once
#include "fp.cuh"

namespace c12 {

// ---- uniform field interface over Fp --------------------------------------------------------------------
C12_HD Fp add(const Fp& a, const Fp& b) { return fp_add(a, b); }
C12_HD Fp sub(const Fp& a, const Fp& b) { return fp_sub(a, b); }
C12_HD Fp mul(const Fp& a, const Fp& b) { return fp_mul(a, b); }
C12_HD Fp sqr(const Fp& a) { return fp_sqr(a); }
C12_HD Fp neg(const Fp& a) { return fp_neg(a); }
C12_HD Fp dbl(const Fp& a) { return fp_dbl(a); }
C12_HD bool is_zero(const Fp& a) { return fp_is_zero(a); }
C12_HD bool eq(const Fp& a, const Fp& b) { return fp_eq(a, b); }
C12_HD Fp select(bool c, const Fp& a, const Fp& b) { return fp_select(c, a, b); }
C12_HD Fp inv(const Fp& a) { return fp_inv(a); }

// ---- Fp2 ------------------------------------------------------------------------------------------------
C12_HD Fp2 fp2_zero() { return Fp2{fp_zero(), fp_zero()}; }
C12_HD Fp2 fp2_one() { return Fp2{fp_one(), fp_zero()}; }
// Additive operations stay inline: as by-value calls they cost ~70 register moves each, no smaller than the inlined
// carry chains, and the pairing kernels measured 3-5 % slower (profiles/r01n).
C12_HD Fp2 add(const Fp2& x, const Fp2& y) { return Fp2{fp_add(x.a, y.a), fp_add(x.b, y.b)}; }
C12_HD Fp2 sub(const Fp2& x, const Fp2& y) { return Fp2{fp_sub(x.a, y.a), fp_sub(x.b, y.b)}; }
C12_HD Fp2 neg(const Fp2& x) { return Fp2{fp_neg(x.a), fp_neg(x.b)}; }
C12_HD Fp2 dbl(const Fp2& x) { return add(x, x); }
C12_HD Fp2 conj(const Fp2& x) { return Fp2{x.a, fp_neg(x.b)}; }
C12_HD bool is_zero(const Fp2& x) { return fp_is_zero(x.a) && fp_is_zero(x.b); }
C12_HD bool eq(const Fp2& x, const Fp2& y) { return fp_eq(x.a, y.a) && fp_eq(x.b, y.b); }
C12_HD Fp2 select(bool c, const Fp2& x, const Fp2& y)
{
    return Fp2{fp_select(c, x.a, y.a), fp_select(c, x.b, y.b)};
}

// Karatsuba: 3 Fp products.  A real call on the device: the towers above it (Fp4/Fp12, G2 curve formulas)
// would otherwise inline ~2,000 SASS instructions per use and thrash the instruction cache.  Operands and result
// travel BY VALUE (in registers): by-reference parameters of a non-inlined function would force every caller to
// park both operands in local memory first.
// -DC12_FP2_INLINE_MULS: the three (two) independent Montgomery products are inlined into the Fp2 product so that ptxas
// can interleave their carry chains (A/B knob for the latency-bound kernels, profiles/).
#if defined(C12_FP2_INLINE_MULS)
#define C12_FP2_MUL fp_mul_inl
#else
#define C12_FP2_MUL fp_mul
#endif
// -DC12_FP2_LAZY (device): lazy reduction - the three products stay 768 bits wide, are combined there (a0 b0 - a1 b1 + p^2 and
// (a0+a1)(b0+b1) - a0 b0 - a1 b1, both < 2 p^2 since p < 2^381) and only the two results are Montgomery-reduced:
// 3 x 144 + 2 x 156 = 744 multiply-adds instead of 900.  The squaring keeps its two products but skips the reductions of the
// sums: (a+b)(a-b+p) and 2ab, both < 4 p^2.  Results are fully reduced, so every caller sees the same values.
#if defined(C12_FP2_LAZY) && defined(__CUDA_ARCH__)
__device__ __forceinline__ Fp2 fp2_mul_body(const Fp2& x, const Fp2& y)
{
    uint32_t sa[12], sb[12], t0[24], t1[24], t2[24];
    fp_add_noreduce_ptx(sa, x.a.v, x.b.v);
    fp_add_noreduce_ptx(sb, y.a.v, y.b.v);
    fp_mulw_ptx(t2, sa, sb);
    fp_mulw_ptx(t0, x.a.v, y.a.v);
    fpw_sub_ptx(t2, t2, t0);
    fp_mulw_ptx(t1, x.b.v, y.b.v);
    fpw_sub_ptx(t2, t2, t1);
    fpw_sub_addp2_ptx(t0, t0, t1);
    Fp2 r;
    fp_redcw_ptx(r.a.v, t0);
    fp_redcw_ptx(r.b.v, t2);
    return r;
}
__device__ __forceinline__ Fp2 fp2_sqr_body(const Fp2& x)
{
    uint32_t s[12], d[12], t0[24], t1[24];
    fp_add_noreduce_ptx(s, x.a.v, x.b.v);
    fp_sub_addp_ptx(d, x.a.v, x.b.v);
    fp_mulw_ptx(t0, s, d);
    fp_mulw_ptx(t1, x.a.v, x.b.v);
    fpw_dbl_ptx(t1, t1);
    Fp2 r;
    fp_redcw_ptx(r.a.v, t0);
    fp_redcw_ptx(r.b.v, t1);
    return r;
}
#else
// The Karatsuba operands a + b (and a - b + p in the squaring) are NOT reduced on the device: the Montgomery product takes
// operands below 2p (4p^2 / R + p < 2p since p < R / 8; emulated in tools/gen_fp_ptx.py) and reduces its result fully, so the
// values are the same and 26 instructions per sum are saved.  -DC12_FP2_REDUCED_SUMS restores the reduced sums (A/B).
#if defined(__CUDA_ARCH__) && !defined(C12_FP2_REDUCED_SUMS)
__device__ __forceinline__ Fp fp_add_wide(const Fp& a, const Fp& b)
{
    Fp r;
    fp_add_noreduce_ptx(r.v, a.v, b.v);
    return r;
}
__device__ __forceinline__ Fp fp_sub_wide(const Fp& a, const Fp& b)   // a - b + p, in [1, 2p)
{
    Fp r;
    fp_sub_addp_ptx(r.v, a.v, b.v);
    return r;
}
#else
C12_HD Fp fp_add_wide(const Fp& a, const Fp& b) { return fp_add(a, b); }
C12_HD Fp fp_sub_wide(const Fp& a, const Fp& b) { return fp_sub(a, b); }
#endif
C12_HD Fp2 fp2_mul_body(const Fp2& x, const Fp2& y)
{
    Fp t0 = C12_FP2_MUL(x.a, y.a);
    Fp t1 = C12_FP2_MUL(x.b, y.b);
    Fp t2 = C12_FP2_MUL(fp_add_wide(x.a, x.b), fp_add_wide(y.a, y.b));
    return Fp2{fp_sub(t0, t1), fp_sub(fp_sub(t2, t0), t1)};
}
// (a+b)(a-b) + 2ab i : 2 Fp products
C12_HD Fp2 fp2_sqr_body(const Fp2& x)
{
    Fp t0 = C12_FP2_MUL(fp_add_wide(x.a, x.b), fp_sub_wide(x.a, x.b));
    Fp t1 = C12_FP2_MUL(x.a, x.b);
    return Fp2{t0, fp_dbl(t1)};
}
#endif
#if defined(__CUDA_ARCH__)
__device__ __noinline__ Fp2 fp2_mul_call(Fp2 x, Fp2 y) { return fp2_mul_body(x, y); }
__device__ __noinline__ Fp2 fp2_sqr_call(Fp2 x) { return fp2_sqr_body(x); }
C12_HD Fp2 mul(const Fp2& x, const Fp2& y) { return fp2_mul_call(x, y); }
C12_HD Fp2 sqr(const Fp2& x) { return fp2_sqr_call(x); }
#else
C12_HD Fp2 mul(const Fp2& x, const Fp2& y) { return fp2_mul_body(x, y); }
C12_HD Fp2 sqr(const Fp2& x) { return fp2_sqr_body(x); }
#endif

// products of the bucket-accumulation hot loop (ec.cuh xyzz_madd): straight-line code over Fp, the regular call over Fp2
C12_HD Fp mul_hot(const Fp& a, const Fp& b) { return fp_mul_inl(a, b); }
C12_HD Fp sqr_hot(const Fp& a) { return fp_sqr_inl(a); }
C12_HD Fp2 mul_hot(const Fp2& x, const Fp2& y) { return mul(x, y); }
C12_HD Fp2 sqr_hot(const Fp2& x) { return sqr(x); }

C12_HD Fp2 mul_fp(const Fp2& x, const Fp& s) { return Fp2{fp_mul(x.a, s), fp_mul(x.b, s)}; }

// x * (1+i): the non-residue of the tower (QNRI = 0)
C12_HD Fp2 mul_ip(const Fp2& x) { return Fp2{fp_sub(x.a, x.b), fp_add(x.a, x.b)}; }

C12_HD Fp2 inv(const Fp2& x)
{
    Fp n = fp_inv(fp_add(fp_sqr(x.a), fp_sqr(x.b)));
    return Fp2{fp_mul(x.a, n), fp_neg(fp_mul(x.b, n))};
}

// ---- square roots (point decompression; replaces FP_sqrt / FP2_sqrt, fp_BLS12381.cpp:842-876, fp2_BLS12381.cpp:460-520)
// Any root will do: the callers fix the sign from the encoding's sign bit.
C12_HD bool fp_sqrt(Fp& r, const Fp& a)
{
    r = fp_sqrt_candidate(a);
    return fp_eq(fp_sqr(r), a);
}
// sqrt(a + b i): with s = sqrt(a^2 + b^2), x = sqrt((a +- s) / 2), y = b / (2 x)
C12_HD bool fp2_sqrt(Fp2& r, const Fp2& z)
{
    if (is_zero(z)) {
        r = fp2_zero();
        return true;
    }
    Fp t;
    if (fp_is_zero(z.b)) {          // real: -1 is a non-residue, so exactly one of a, -a has a root in Fp
        if (fp_sqrt(t, z.a)) {
            r = Fp2{t, fp_zero()};
            return true;
        }
        if (!fp_sqrt(t, fp_neg(z.a))) return false;
        r = Fp2{fp_zero(), t};
        return true;
    }
    Fp s;
    if (!fp_sqrt(s, fp_add(fp_sqr(z.a), fp_sqr(z.b)))) return false;   // the norm of a square is a square
    Fp x;
    if (!fp_sqrt(x, fp_mul(fp_add(z.a, s), fp_half_m())) && !fp_sqrt(x, fp_mul(fp_sub(z.a, s), fp_half_m()))) return false;
    Fp y = fp_mul(z.b, fp_inv(fp_dbl(x)));
    r = Fp2{x, y};
    return eq(sqr(r), z);
}

C12_HD Fp2 mul3(const Fp2& x) { return Fp2{fp_mul3(x.a), fp_mul3(x.b)}; }
C12_HD Fp2 mul4(const Fp2& x) { return Fp2{fp_mul4(x.a), fp_mul4(x.b)}; }
C12_HD Fp2 mul8(const Fp2& x) { return Fp2{fp_mul8(x.a), fp_mul8(x.b)}; }
C12_HD Fp2 mul12(const Fp2& x) { return Fp2{fp_mul12(x.a), fp_mul12(x.b)}; }
C12_HD Fp mul3(const Fp& x) { return fp_mul3(x); }
C12_HD Fp mul4(const Fp& x) { return fp_mul4(x); }
C12_HD Fp mul8(const Fp& x) { return fp_mul8(x); }
C12_HD Fp mul12(const Fp& x) { return fp_mul12(x); }

// FP2_sign (3rd-party/miracl-core/fp2_BLS12381.cpp:168-181): parity of the real part, of the imaginary part
// when the real part is zero
C12_HD int fp2_sign(const Fp2& x) { return fp_is_zero(x.a) ? fp_sign(x.b) : fp_sign(x.a); }

// field traits used by ec.cuh
template <class F> struct FieldOps;
template <> struct FieldOps<Fp> {
    static C12_HD Fp zero() { return fp_zero(); }
    static C12_HD Fp one() { return fp_one(); }
    // 3b with b = 4  (G1: y^2 = x^3 + 4)
    static C12_HD Fp mul_b3(const Fp& x) { return fp_mul12(x); }
};
template <> struct FieldOps<Fp2> {
    static C12_HD Fp2 zero() { return fp2_zero(); }
    static C12_HD Fp2 one() { return fp2_one(); }
    // 3b' with b' = 4(1+i)  (M-type twist, 3rd-party/miracl-core/ecp2_BLS12381.cpp:270-296)
    static C12_HD Fp2 mul_b3(const Fp2& x) { return mul_ip(mul12(x)); }
};

} // namespace c12
