// Fp on 13 x 30-bit UNSATURATED limbs: the experimental field layer of VERDICT r01 task 4 / DESIGN.md §7.
//
// Why: the shipped Montgomery product (fp_ptx.inc) is 288 carry-chained IMAD.WIDE.X, which the fmaheavy pipe issues at
// ~8.6 T/s against ~13.6 T/s for plain IMAD.WIDE (c12381_probe kinds 1 / 2).  With 30-bit limbs a 32 x 32 -> 64 product has
// 4 bits of headroom, so 13 of them add up in a 64-bit column with NO carry: the product and the reduction become
// 169 + 169 independent IMAD.WIDE (plus 13 IMAD for the quotient digits), and the carries are resolved afterwards by shifts
// and adds on the ALU pipe, which idles in the carry-chain version.  Montgomery radix R' = 2^390.
//
// Conventions: limb k holds bits [30 k, 30 k + 30) of the integer; "normalised" = every limb < 2^30; fp30_mul accepts limbs up
// to 2^30 + 2^26 (13 products of that size still fit 64 bits) and values up to ~25 p, and returns a normalised value < 1.1 p
// (a b / 2^390 + p with p / 2^390 < 2^-9).  Host+device; tests/hostmirror runs the same code on the CPU.
#pragma once
#include "fp.cuh"

namespace c12 {

struct Fp30 {
    uint32_t v[13];
};

constexpr uint32_t FP30_MASK = 0x3fffffffu;
#ifndef FP30_MID_LO
#define FP30_MID_LO 7
#define FP30_MID_HI 17
#endif
#define C12_P30_LIMBS {0x3fffaaabu, 0x27fbffffu, 0x153ffffbu, 0x2affffacu, 0x30f6241eu, 0x034a83dau, 0x112bf673u, 0x12e13ce1u, 0x2cd76477u, 0x1ed90d2eu, 0x29a4b1bau, 0x3a8e5ff9u, 0x001a0111u}
constexpr uint32_t FP30_M0 = 0x3ffcfffdu;     // -p^-1 mod 2^30

// the same integer, repacked: 12 x 32 -> 13 x 30 (normalised)
C12_HD Fp30 fp30_from_fp(const Fp& a)
{
    Fp30 r;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
        uint32_t x = a.v[w] >> sh;
        if (sh > 2 && w + 1 < 12) x |= a.v[w + 1] << (32 - sh);
        r.v[i] = x & FP30_MASK;
    }
    return r;
}

// 13 x 30 (normalised, value < 2^384) -> 12 x 32, same integer
C12_HD Fp fp30_pack(const Fp30& a)
{
    Fp r;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const int bit = 32 * j, i0 = bit / 30, o = bit - 30 * i0;
        uint32_t x = a.v[i0] >> o;
        if (i0 + 1 < 13) x |= a.v[i0 + 1] << (30 - o);
        if (60 - o < 32 && i0 + 2 < 13) x |= a.v[i0 + 2] << (60 - o);
        r.v[j] = x;
    }
    return r;
}

// value < 2 p -> the canonical residue, packed
C12_HD Fp fp30_to_fp_canonical(const Fp30& a) { return fp_add(fp30_pack(a), fp_zero()); }

// 2^378 mod p as plain limbs: fp_mul(t, .) = t 2^-6, which turns (x y 2^-384) into (x y 2^-390) in cross checks
C12_HD Fp fp30_check_const()
{
    return Fp{{0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u, 0x04000000u}};
}

// c += a * b on a 64-bit column.  On the device this is spelled in PTX: written in C, nvcc widens the constant modulus limbs to
// 64 bits and follows every IMAD.WIDE by an addition of a zero high half (about a hundred wasted IADD3 per product).
C12_HD void fp30_mac(uint64_t& c, uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    uint64_t t;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a), "r"(b));     // opaque to nvcc; ptxas fuses it with the addition below
    c += t;
#else
    c += (uint64_t)a * b;
#endif
}

// the modulus limbs come from constant memory on the device (an IMAD.WIDE takes a constant-bank operand directly): as literals
// nvcc widens them to 64 bits and follows every multiply-add by an addition of the (zero) high half
#if defined(__CUDACC__)
__constant__ uint32_t FP30_P_CONST[13] = C12_P30_LIMBS;
#endif

// a b 2^-390 mod p.  Operand scanning into 26 64-bit columns; one carry-save pass between the product and the reduction so that
// a column never holds more than 13 full-size products; the 13 reduction rows retire one column each; a final carry pass.
C12_HD Fp30 fp30_mul(const Fp30& a, const Fp30& b)
{
#if defined(__CUDA_ARCH__)
    const uint32_t* p = FP30_P_CONST;
#else
    const uint32_t p[13] = C12_P30_LIMBS;
#endif
    uint64_t c[26];
#pragma unroll
    for (int k = 0; k < 26; ++k) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 13; ++i)
#pragma unroll
        for (int j = 0; j < 13; ++j) fp30_mac(c[i + j], a.v[i], b.v[j]);
    // carry-save between the two halves, only where it is needed: column k receives min(k + 1, 25 - k) products here and as
    // many from the reduction rows; 16 products of (2^30)^2 fit 64 bits, so only columns 7 .. 17 could overflow.  Each of
    // them keeps its low 30 bits and hands its high part to the next column (all from the OLD values).
    uint64_t hp = 0;
#pragma unroll
    for (int k = FP30_MID_LO; k <= FP30_MID_HI + 1; ++k) {
        const uint64_t h = k <= FP30_MID_HI ? c[k] >> 30 : 0;
        c[k] = (k <= FP30_MID_HI ? (c[k] & FP30_MASK) : c[k]) + hp;
        hp = h;
    }
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const uint32_t m = ((uint32_t)c[i] * FP30_M0) & FP30_MASK;
#pragma unroll
        for (int j = 0; j < 13; ++j) fp30_mac(c[i + j], m, p[j]);
        c[i + 1] += c[i] >> 30;
    }
    Fp30 r;
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 13; ++k) {
        const uint64_t t = c[13 + k] + carry;
        r.v[k] = k < 12 ? (uint32_t)t & FP30_MASK : (uint32_t)t;
        carry = t >> 30;
    }
    return r;
}

} // namespace c12
