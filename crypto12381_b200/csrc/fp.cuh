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
#if defined(__CUDA_ARCH__)
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
#if defined(__CUDA_ARCH__)
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
// Binary extended Euclid on the plain 384-bit integers (Replaces FP_inv, 3rd-party/miracl-core/fp_BLS12381.cpp:817,
// which is Fermat a^(p-2) through FP_progen).  About 2*381 shift/subtract steps of 12-limb add/sub instead of
// ~480 Montgomery products: an order of magnitude fewer instructions, which matters most where ONE thread
// normalises a final result (the serial tail of an MSM).  Not constant time — all inputs here are public.
// inv(0) = 0.  Inversions are still amortised by Montgomery's trick wherever a batch exists.
namespace detail {
C12_HD bool limbs_is_one(const uint32_t (&a)[12])
{
    uint32_t z = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 12; ++i) z |= a[i];
    return z == 0;
}
C12_HD bool limbs_ge(const uint32_t (&a)[12], const uint32_t (&b)[12])
{
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        uint64_t d = (uint64_t)a[i] - b[i] - borrow;
        borrow = (d >> 32) & 1u;
    }
    return borrow == 0;
}
C12_HD void limbs_sub(uint32_t (&a)[12], const uint32_t (&b)[12])  // a -= b, a >= b
{
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        uint64_t d = (uint64_t)a[i] - b[i] - borrow;
        a[i] = (uint32_t)d;
        borrow = (d >> 32) & 1u;
    }
}
// x = x / 2 mod p for x < p: (x even ? x : x + p) >> 1
C12_HD void limbs_half_mod_p(uint32_t (&x)[12])
{
    const uint32_t pl[12] = C12_P_LIMBS;
    const uint32_t mask = 0u - (x[0] & 1u);
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        uint64_t t = (uint64_t)x[i] + (pl[i] & mask) + c;
        x[i] = (uint32_t)t;
        c = t >> 32;
    }
#pragma unroll
    for (int i = 0; i < 11; ++i) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[11] = (x[11] >> 1) | ((uint32_t)c << 31);
}
C12_HD void limbs_shr1(uint32_t (&x)[12])
{
#pragma unroll
    for (int i = 0; i < 11; ++i) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[11] >>= 1;
}
} // namespace detail

C12_HD_NOINLINE Fp fp_inv(const Fp& a)
{
    if (fp_is_zero(a)) return fp_zero();
    const uint32_t pl[12] = C12_P_LIMBS;
    Fp u = a, v, x1 = fp_zero(), x2 = fp_zero();   // u = a R (as a plain integer), v = p
#pragma unroll
    for (int i = 0; i < 12; ++i) v.v[i] = pl[i];
    x1.v[0] = 1;
#pragma unroll 1
    while (!detail::limbs_is_one(u.v) && !detail::limbs_is_one(v.v)) {
#pragma unroll 1
        while (!(u.v[0] & 1u)) {
            detail::limbs_shr1(u.v);
            detail::limbs_half_mod_p(x1.v);
        }
#pragma unroll 1
        while (!(v.v[0] & 1u)) {
            detail::limbs_shr1(v.v);
            detail::limbs_half_mod_p(x2.v);
        }
        if (detail::limbs_ge(u.v, v.v)) {
            detail::limbs_sub(u.v, v.v);
            x1 = fp_sub(x1, x2);
        } else {
            detail::limbs_sub(v.v, u.v);
            x2 = fp_sub(x2, x1);
        }
    }
    // (a R)^-1; times R^3 / R  ->  a^-1 R
    const Fp r3 = {{0xd94ca1e0u, 0xed48ac6bu, 0x03a7adf8u, 0x315f831eu, 0x615e29ddu, 0x9a53352au, 0x921e1761u, 0x34c04e5eu, 0x65724728u,
                    0x2512d435u, 0x91755d4du, 0x0aa63460u}};
    return fp_mul(detail::limbs_is_one(u.v) ? x1 : x2, r3);
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
