// Fp: the BLS12-381 base field on 12 x 32-bit limbs, Montgomery form with R = 2^384.
//
// Replaces, for the device path, MIRACL's BIG_mul/BIG_sqr/BIG_monty/FP_* on 7 x 58-bit limbs
// (reference: 3rd-party/miracl-core/big_B384_58.cpp:570,720,836; fp_BLS12381.cpp:223-252,396-415,466-588).
// Values are always fully reduced (in [0, p)), so equality is limb equality and there is no XES excess
// bookkeeping.
//
// On the device every primitive is one inline-PTX block generated (and emulated bit-exactly) by
// tools/gen_fp_ptx.py.  When this header is compiled by a host compiler (tests/hostmirror only — the
// product never runs field arithmetic on the CPU) the same functions are provided in portable C++ so the
// curve/tower/pairing templates above them can be checked against the oracle without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define C12_HD __host__ __device__ __forceinline__
#define C12_D __device__ __forceinline__
#define C12_HD_NOINLINE __host__ __device__ __noinline__ inline
#else
#define C12_HD inline
#define C12_D inline
#define C12_HD_NOINLINE inline
#endif

namespace c12 {

struct Fp {
    uint32_t v[12];
};

// Fp2 = Fp[i]/(i^2+1), a + b*i  (reference: 3rd-party/miracl-core/fp2_BLS12381.h:33-37)
struct Fp2 {
    Fp a;
    Fp b;
};

#if defined(__CUDA_ARCH__)
#include "fp_ptx.inc"
#endif
// -DC12_FP_SQR_DEDICATED: squarings through the dedicated routine (66 doubled cross products + 12 diagonal ones + a separate
// reduction: 234 multiply-adds) instead of the interleaved product with both operands equal (300); A/B knob, profiles/.
#if defined(C12_FP_SQR_DEDICATED)
#define C12_FP_SQR_PTX fp_sqr_dedicated_ptx
#else
#define C12_FP_SQR_PTX fp_sqr_ptx
#endif

#define C12_P_LIMBS                                                                                               \
    {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u,      \
     0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau}
#define C12_M0 0xfffcfffdu

#if !defined(__CUDA_ARCH__)
namespace host {
static const uint32_t PL[12] = C12_P_LIMBS;
inline void cond_sub_p(uint32_t (&r)[12], const uint32_t (&t)[13])
{
    uint32_t s[12];
    int64_t borrow = 0;
    for (int i = 0; i < 12; ++i) {
        int64_t d = (int64_t)t[i] - PL[i] + borrow;
        s[i] = (uint32_t)d;
        borrow = d >> 32;
    }
    bool ge = ((int64_t)t[12] + borrow) >= 0;
    for (int i = 0; i < 12; ++i) r[i] = ge ? s[i] : t[i];
}
#if defined(C12_COUNT_FP_MUL)
inline unsigned long long& addsub_counter()     // host mirror only: field additions / subtractions / negations executed
{
    static unsigned long long n = 0;
    return n;
}
inline unsigned long long& mont_mul_counter()   // host mirror only: Montgomery products executed (the algorithmic work unit of DESIGN.md)
{
    static unsigned long long n = 0;
    return n;
}
#endif
inline void mont_mul(uint32_t (&r)[12], const uint32_t (&a)[12], const uint32_t (&b)[12])
{
#if defined(C12_COUNT_FP_MUL)
    ++mont_mul_counter();
#endif
    uint32_t t[14] = {0};
    for (int i = 0; i < 12; ++i) {
        uint64_t c = 0;
        for (int j = 0; j < 12; ++j) {
            uint64_t x = (uint64_t)a[j] * b[i] + t[j] + c;
            t[j] = (uint32_t)x;
            c = x >> 32;
        }
        uint64_t x = (uint64_t)t[12] + c;
        t[12] = (uint32_t)x;
        t[13] = (uint32_t)(x >> 32);
        uint32_t m = t[0] * C12_M0;
        c = ((uint64_t)m * PL[0] + t[0]) >> 32;
        for (int j = 1; j < 12; ++j) {
            uint64_t y = (uint64_t)m * PL[j] + t[j] + c;
            t[j - 1] = (uint32_t)y;
            c = y >> 32;
        }
        x = (uint64_t)t[12] + c;
        t[11] = (uint32_t)x;
        t[12] = t[13] + (uint32_t)(x >> 32);
        t[13] = 0;
    }
    uint32_t tt[13];
    for (int i = 0; i < 13; ++i) tt[i] = t[i];
    cond_sub_p(r, tt);
}
} // namespace host
#endif

// On the device the Montgomery product / squaring are REAL calls by default (operands and result travel in
// registers: the ABI passes these 48-byte structs by value without touching local memory).  That divides the SASS
// footprint of every kernel by ~10 at the price of ~36 register moves per product; measured (profiles/r01c): the
// instruction-cache misses of fully inlined tower / curve code cost more than the calls (pairings 1.19 M/s against
// 0.91 M/s, MSM tail 5.2 ms against 6.0 ms).  The one loop where straight-line code wins - the bucket accumulation's
// XYZZ addition - uses fp_mul_inl / fp_sqr_inl explicitly.  -DC12_FP_INLINE_ALL restores inlining everywhere (A/B).
#if defined(__CUDA_ARCH__) && !defined(C12_FP_INLINE_ALL)
#define C12_FP_CALL 1
__device__ __noinline__ Fp fp_mul_call(Fp a, Fp b)
{
    Fp r;
    fp_mul_ptx(r.v, a.v, b.v);
    return r;
}
__device__ __noinline__ Fp fp_sqr_call(Fp a)
{
    Fp r;
    C12_FP_SQR_PTX(r.v, a.v);
    return r;
}
#endif

// always-inline product / squaring (hot loops only)
C12_HD Fp fp_mul_inl(const Fp& a, const Fp& b)
{
    Fp r;
#if defined(__CUDA_ARCH__)
    fp_mul_ptx(r.v, a.v, b.v);
#else
    host::mont_mul(r.v, a.v, b.v);
#endif
    return r;
}
C12_HD Fp fp_sqr_inl(const Fp& a)
{
    Fp r;
#if defined(__CUDA_ARCH__)
    C12_FP_SQR_PTX(r.v, a.v);
#else
    host::mont_mul(r.v, a.v, a.v);
#endif
    return r;
}

C12_HD Fp fp_mul(const Fp& a, const Fp& b)
{
#if defined(__CUDA_ARCH__) && defined(C12_FP_CALL)
    return fp_mul_call(a, b);
#else
    return fp_mul_inl(a, b);
#endif
}

C12_HD Fp fp_sqr(const Fp& a)
{
#if defined(__CUDA_ARCH__) && defined(C12_FP_CALL)
    return fp_sqr_call(a);
#else
    return fp_sqr_inl(a);
#endif
}

C12_HD Fp fp_add(const Fp& a, const Fp& b)
{
    Fp r;
#if defined(C12_COUNT_FP_MUL) && !defined(__CUDA_ARCH__)
    ++host::addsub_counter();
#endif
#if defined(__CUDA_ARCH__) && defined(C12_EXP_LAZY_ADDS) && defined(C12_PAIRING_TU)
    fp_add_noreduce_ptx(r.v, a.v, b.v);      // TIMING EXPERIMENT ONLY (wrong values): what the tower would cost with bounds-tracked lazy additions
#elif defined(__CUDA_ARCH__)
    fp_add_ptx(r.v, a.v, b.v);
#else
    uint32_t t[13];
    uint64_t c = 0;
    for (int i = 0; i < 12; ++i) {
        uint64_t x = (uint64_t)a.v[i] + b.v[i] + c;
        t[i] = (uint32_t)x;
        c = x >> 32;
    }
    t[12] = (uint32_t)c;
    host::cond_sub_p(r.v, t);
#endif
    return r;
}

C12_HD Fp fp_sub(const Fp& a, const Fp& b)
{
    Fp r;
#if defined(C12_COUNT_FP_MUL) && !defined(__CUDA_ARCH__)
    ++host::addsub_counter();
#endif
#if defined(__CUDA_ARCH__) && defined(C12_EXP_LAZY_ADDS) && defined(C12_PAIRING_TU)
    fp_sub_addp_ptx(r.v, a.v, b.v);          // TIMING EXPERIMENT ONLY
#elif defined(__CUDA_ARCH__)
    fp_sub_ptx(r.v, a.v, b.v);
#else
    int64_t borrow = 0;
    for (int i = 0; i < 12; ++i) {
        int64_t d = (int64_t)a.v[i] - b.v[i] + borrow;
        r.v[i] = (uint32_t)d;
        borrow = d >> 32;
    }
    if (borrow) {
        uint64_t c = 0;
        for (int i = 0; i < 12; ++i) {
            uint64_t x = (uint64_t)r.v[i] + host::PL[i] + c;
            r.v[i] = (uint32_t)x;
            c = x >> 32;
        }
    }
#endif
    return r;
}

C12_HD Fp fp_neg(const Fp& a)
{
    Fp r;
#if defined(C12_COUNT_FP_MUL) && !defined(__CUDA_ARCH__)
    ++host::addsub_counter();
#endif
#if defined(__CUDA_ARCH__)
    fp_neg_ptx(r.v, a.v);
#else
    uint32_t z = 0;
    for (int i = 0; i < 12; ++i) z |= a.v[i];
    int64_t borrow = 0;
    for (int i = 0; i < 12; ++i) {
        int64_t d = (int64_t)host::PL[i] - a.v[i] + borrow;
        r.v[i] = z ? (uint32_t)d : 0u;
        borrow = d >> 32;
    }
#endif
    return r;
}

// a * 2^-384 mod p : leaves Montgomery form
C12_HD Fp fp_redc(const Fp& a)
{
    Fp r;
#if defined(__CUDA_ARCH__)
    fp_redc_ptx(r.v, a.v);
#else
    Fp one = {{1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};
    host::mont_mul(r.v, a.v, one.v);
#endif
    return r;
}

C12_HD Fp fp_zero()
{
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; ++i) r.v[i] = 0;
    return r;
}

C12_HD bool fp_is_zero(const Fp& a)
{
    uint32_t z = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) z |= a.v[i];
    return z == 0;
}

C12_HD bool fp_eq(const Fp& a, const Fp& b)
{
    uint32_t z = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) z |= a.v[i] ^ b.v[i];
    return z == 0;
}

C12_HD Fp fp_dbl(const Fp& a) { return fp_add(a, a); }

C12_HD Fp fp_select(bool c, const Fp& a, const Fp& b)
{
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; ++i) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
}

// small-constant multiples by addition chains (3, 4, 8, 12 are what the curve formulas need)
C12_HD Fp fp_mul3(const Fp& a) { return fp_add(fp_dbl(a), a); }
C12_HD Fp fp_mul4(const Fp& a) { return fp_dbl(fp_dbl(a)); }
C12_HD Fp fp_mul8(const Fp& a) { return fp_dbl(fp_mul4(a)); }
C12_HD Fp fp_mul12(const Fp& a) { return fp_mul4(fp_mul3(a)); }

#include "constants_gen.cuh"

C12_HD Fp fp_one() { return fp_one_m(); }
C12_HD Fp fp_to_mont(const Fp& a) { return fp_mul(fp_r2(), a); }  // a may be any 384-bit value (2nd operand)
C12_HD Fp fp_from_mont(const Fp& a) { return fp_redc(a); }

// ---- inversion ---------------------------------------------------------------------------------------------
// Bernstein-Yang "safegcd" division steps (eprint 2019/266) on 13 signed 30-bit limbs, in batches of 30 steps: a batch walks
// the LOW WORDS of f and g only (30 x ~14 single-word instructions) and yields a 2 x 2 transition matrix that is then
// applied to the full f, g (exactly divisible by 2^30) and, modulo p, to the cofactors d, e - 8 x 13 signed multiply-adds.
// Replaces FP_inv (3rd-party/miracl-core/fp_BLS12381.cpp:817, Fermat a^(p-2) through FP_progen: ~480 Montgomery products).
// About 27 batches = ~20 k instructions, and - unlike a binary Euclid with its data-dependent inner loops - every lane of a
// warp runs the same instruction stream (the batch body is branch-free; lanes only differ in the batch at which g reaches 0,
// by one or two), which is what matters where 32 lanes invert 32 different elements (k_ba_inv) and where ONE thread
// normalises a final result (the serial tail of an MSM).  The loop runs until g = 0, so it relies on no iteration bound; the
// proven one is 1,101 steps = 37 batches for inputs below 2^381 (Thm. 11.2).  Not constant time.  inv(0) = 0.
// Inversions are still amortised by Montgomery's trick wherever a batch exists.
namespace detail {
struct S30 {
    int32_t v[13];      // value = sum v[i] 2^(30 i)
};
struct Trans30 {
    int32_t u, v, q, r;
};
#define C12_P30_SIGNED {0x3fffaaab, 0x27fbffff, 0x153ffffb, 0x2affffac, 0x30f6241e, 0x034a83da, 0x112bf673, 0x12e13ce1, 0x2cd76477, 0x1ed90d2e, 0x29a4b1ba, 0x3a8e5ff9, 0x001a0111}
constexpr uint32_t P_INV30 = 0x00030003u;   // p^-1 mod 2^30
constexpr int32_t M30 = 0x3fffffff;

// 30 division steps on the low words.  zeta = -(delta + 1/2); returns the new zeta, t maps (f, g) to (f', g') 2^30.
C12_HD int32_t safegcd_divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, Trans30& t)
{
    uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 5
    for (int i = 0; i < 30; ++i) {
        uint32_t c1 = (uint32_t)(zeta >> 31);         // all ones when delta > 0 ...
        const uint32_t c2 = 0u - (g & 1u);            // ... and g is odd: swap-and-subtract, else add f to g when g is odd
        const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
        g += x & c2;
        q += y & c2;
        r += z & c2;
        c1 &= c2;
        zeta = (int32_t)(((uint32_t)zeta ^ c1) - 1u);
        f += g & c1;
        u += q & c1;
        v += r & c1;
        g >>= 1;
        u <<= 1;
        v <<= 1;
    }
    t.u = (int32_t)u;
    t.v = (int32_t)v;
    t.q = (int32_t)q;
    t.r = (int32_t)r;
    return zeta;
}

// (d, e) <- t (d, e) / 2^30 mod p, both kept in (-2p, p)
C12_HD void safegcd_update_de(S30& d, S30& e, const Trans30& t)
{
    const int32_t pl[13] = C12_P30_SIGNED;
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    const int32_t sd = d.v[12] >> 31, se = e.v[12] >> 31;
    int32_t md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);
    int64_t cd = u * d.v[0] + v * e.v[0], ce = q * d.v[0] + r * e.v[0];
    md -= (int32_t)((P_INV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
    me -= (int32_t)((P_INV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
    cd += (int64_t)pl[0] * md;
    ce += (int64_t)pl[0] * me;
    cd >>= 30;      // the low 30 bits are zero by the choice of md, me
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < 13; ++i) {
        cd += u * d.v[i] + v * e.v[i] + (int64_t)pl[i] * md;
        ce += q * d.v[i] + r * e.v[i] + (int64_t)pl[i] * me;
        d.v[i - 1] = (int32_t)cd & M30;
        e.v[i - 1] = (int32_t)ce & M30;
        cd >>= 30;
        ce >>= 30;
    }
    d.v[12] = (int32_t)cd;
    e.v[12] = (int32_t)ce;
}

// (f, g) <- t (f, g) / 2^30 (exact)
C12_HD void safegcd_update_fg(S30& f, S30& g, const Trans30& t)
{
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    int64_t cf = u * f.v[0] + v * g.v[0], cg = q * f.v[0] + r * g.v[0];
    cf >>= 30;
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < 13; ++i) {
        cf += u * f.v[i] + v * g.v[i];
        cg += q * f.v[i] + r * g.v[i];
        f.v[i - 1] = (int32_t)cf & M30;
        g.v[i - 1] = (int32_t)cg & M30;
        cf >>= 30;
        cg >>= 30;
    }
    f.v[12] = (int32_t)cf;
    g.v[12] = (int32_t)cg;
}

// d in (-2p, p), negated when `sign` is negative -> [0, p), limbs in [0, 2^30)
C12_HD void safegcd_normalize(S30& d, int32_t sign)
{
    const int32_t pl[13] = C12_P30_SIGNED;
    const int32_t neg = sign >> 31;
    int32_t add = d.v[12] >> 31, carry = 0;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        int32_t x = d.v[i] + (pl[i] & add);
        x = (x ^ neg) - neg + carry;
        carry = i < 12 ? x >> 30 : 0;
        d.v[i] = i < 12 ? x & M30 : x;
    }
    add = d.v[12] >> 31;        // now in (-p, p)
    carry = 0;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        int32_t x = d.v[i] + (pl[i] & add) + carry;
        carry = i < 12 ? x >> 30 : 0;
        d.v[i] = i < 12 ? x & M30 : x;
    }
}
} // namespace detail

C12_HD_NOINLINE Fp fp_inv(const Fp& a)
{
    using namespace detail;
    const int32_t pl[13] = C12_P30_SIGNED;
    S30 d, e, f, g;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        d.v[i] = 0;
        e.v[i] = i == 0 ? 1 : 0;
        f.v[i] = pl[i];
        // 12 x 32 -> 13 x 30: a R as a plain integer
        const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
        uint32_t x = a.v[w] >> sh;
        if (sh > 2 && w + 1 < 12) x |= a.v[w + 1] << (32 - sh);
        g.v[i] = (int32_t)(x & (uint32_t)M30);
    }
    int32_t zeta = -1;
#pragma unroll 1
    for (int batch = 0; batch < 48; ++batch) {
        int32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 13; ++i) nz |= g.v[i];
        if (nz == 0) break;
        Trans30 t;
        zeta = safegcd_divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        safegcd_update_de(d, e, t);
        safegcd_update_fg(f, g, t);
    }
    // f = +-gcd = +-1 (or +-p when a = 0, where d = 0 anyway): d = +-(a R)^-1
    safegcd_normalize(d, f.v[12]);
    Fp x;
#pragma unroll
    for (int j = 0; j < 12; ++j) {          // 13 x 30 -> 12 x 32
        const int bit = 32 * j, i0 = bit / 30, o = bit - 30 * i0;
        uint32_t w = (uint32_t)d.v[i0] >> o;
        if (i0 + 1 < 13) w |= (uint32_t)d.v[i0 + 1] << (30 - o);
        if (60 - o < 32 && i0 + 2 < 13) w |= (uint32_t)d.v[i0 + 2] << (60 - o);
        x.v[j] = w;
    }
    // (a R)^-1; times R^3 / R  ->  a^-1 R
    const Fp r3 = {{0xd94ca1e0u, 0xed48ac6bu, 0x03a7adf8u, 0x315f831eu, 0x615e29ddu, 0x9a53352au, 0x921e1761u, 0x34c04e5eu, 0x65724728u,
                    0x2512d435u, 0x91755d4du, 0x0aa63460u}};
    return fp_mul(x, r3);
}

// a^((p+1)/4): the square root when a is a residue (p = 3 mod 4).  Replaces FP_sqrt
// (3rd-party/miracl-core/fp_BLS12381.cpp:842-876); caller checks r^2 == a.
C12_HD_NOINLINE Fp fp_sqrt_candidate(const Fp& a)
{
    // (p+1)/4
    const uint32_t e[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                            0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};
    Fp tab[16];
    tab[0] = fp_one();
    tab[1] = a;
#pragma unroll 1
    for (int i = 2; i < 16; ++i) tab[i] = fp_mul(tab[i - 1], a);
    Fp r = fp_one();
#pragma unroll 1
    for (int i = 95; i >= 0; --i) {
        if (i != 95) {
            r = fp_sqr(r);
            r = fp_sqr(r);
            r = fp_sqr(r);
            r = fp_sqr(r);
        }
        uint32_t d = (e[i >> 3] >> ((i & 7) * 4)) & 15u;
        if (d) r = fp_mul(r, tab[d]);
    }
    return r;
}

// 48-byte big-endian <-> plain limbs (not Montgomery).  Wire format of FP/BIG: BIG_toBytes/BIG_fromBytes
// (3rd-party/miracl-core/big_B384_58.cpp:171,185).
C12_HD Fp fp_from_be48(const uint8_t* b)
{
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const uint8_t* q = b + 44 - 4 * i;
        r.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    return r;
}

C12_HD void fp_to_be48(uint8_t* b, const Fp& a)
{
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        uint8_t* q = b + 44 - 4 * i;
        q[0] = (uint8_t)(a.v[i] >> 24);
        q[1] = (uint8_t)(a.v[i] >> 16);
        q[2] = (uint8_t)(a.v[i] >> 8);
        q[3] = (uint8_t)(a.v[i]);
    }
}

// plain value < p ?
C12_HD bool fp_is_canonical(const Fp& a)
{
    const uint32_t pl[12] = C12_P_LIMBS;
    bool lt = false, decided = false;
#pragma unroll
    for (int i = 11; i >= 0; --i) {
        if (!decided && a.v[i] != pl[i]) {
            lt = a.v[i] < pl[i];
            decided = true;
        }
    }
    return lt;
}

// parity of the canonical (non-Montgomery) value: FP_sign (3rd-party/miracl-core/fp_BLS12381.cpp:912-936)
C12_HD int fp_sign(const Fp& a_mont) { return (int)(fp_from_mont(a_mont).v[0] & 1u); }

} // namespace c12
