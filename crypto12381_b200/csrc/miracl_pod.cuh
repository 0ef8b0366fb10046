// Conversions between the reference's in-memory PODs and this library's field elements; host+device
// bodies so tests/hostmirror can check them against structs produced by the compiled reference.
//   big = int64[7], 58-bit digits (possibly un-normalised); fp = { big g; int32 xes }: Montgomery residue with
//   R = 2^406, value < xes * p.   (include/crypto12381/miracl_core_interface.hpp:32-33,76-79;
//   3rd-party/miracl-core/fp_BLS12381.cpp:223-252)
#pragma once
#include "msm_core.cuh"

namespace c12 {

constexpr int POD_FP = 64, POD_BIG = 56, POD_P1 = 192, POD_P2 = 384, POD_FP12 = 776;

// int64[7] digits base 2^58 (signed, un-normalised allowed) -> 13 x u32 words of the non-negative value (< 2^416)
C12_HD bool big_to_words(const long long* g, uint32_t (&w)[13])
{
    unsigned long long d[7];
    long long carry = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        long long t = g[i] + carry;
        if (i < 6) {
            d[i] = (unsigned long long)t & ((1ull << 58) - 1ull);
            carry = t >> 58;
        } else {
            d[i] = (unsigned long long)t;
            carry = t < 0 ? -1 : 0;
        }
    }
#pragma unroll
    for (int i = 0; i < 13; ++i) w[i] = 0;
    bool ok = carry == 0 && (d[6] >> 58) == 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const int o = 58 * i, word = o >> 5, sh = o & 31;
        unsigned long long lo = d[i] << sh;
        unsigned long long hi = sh ? (d[i] >> (64 - sh)) : 0ull;
        w[word] |= (uint32_t)lo;
        if (word + 1 < 13) w[word + 1] |= (uint32_t)(lo >> 32);
        if (word + 2 < 13) w[word + 2] |= (uint32_t)hi;
    }
    return ok;
}

C12_HD void words_to_big(const uint32_t (&w)[12], long long* g)
{
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const int o = 58 * i, word = o >> 5, sh = o & 31;
        unsigned long long x = (unsigned long long)w[word] >> sh;
        if (word + 1 < 12) x |= (unsigned long long)w[word + 1] << (32 - sh);
        if (word + 2 < 12 && sh) x |= (unsigned long long)w[word + 2] << (64 - sh);
        g[i] = (long long)(x & ((1ull << 58) - 1ull));
    }
}

// reference fp (64 B) -> our Montgomery residue
C12_HD Fp fp_from_pod(const uint8_t* pod, bool& ok)
{
    uint32_t w[13];
    if (!big_to_words(reinterpret_cast<const long long*>(pod), w)) ok = false;
    Fp lo, hi = fp_zero();
#pragma unroll
    for (int i = 0; i < 12; ++i) lo.v[i] = w[i];
    hi.v[0] = w[12];
    return fp_add(fp_mul(fp_k_362(), lo), fp_mul(fp_k_746(), hi));
}

C12_HD void fp_to_pod(uint8_t* pod, const Fp& a)
{
    Fp m = fp_mul(a, fp_k_406());
    words_to_big(m.v, reinterpret_cast<long long*>(pod));
    *reinterpret_cast<int*>(pod + 56) = 1;   // XES: fully reduced
    *reinterpret_cast<int*>(pod + 60) = 0;   // padding
}


C12_HD Fp2 fp2_from_pod(const uint8_t* pod, bool& ok) { return Fp2{fp_from_pod(pod, ok), fp_from_pod(pod + POD_FP, ok)}; }
C12_HD void fp2_to_pod(uint8_t* pod, const Fp2& a)
{
    fp_to_pod(pod, a.a);
    fp_to_pod(pod + POD_FP, a.b);
}

template <class F> struct Pod;
template <> struct Pod<Fp> {
    static constexpr int POINT = POD_P1, COORD = POD_FP;
    static C12_HD Fp load(const uint8_t* p, bool& ok) { return fp_from_pod(p, ok); }
    static C12_HD void store(uint8_t* p, const Fp& a) { fp_to_pod(p, a); }
};
template <> struct Pod<Fp2> {
    static constexpr int POINT = POD_P2, COORD = 2 * POD_FP;
    static C12_HD Fp2 load(const uint8_t* p, bool& ok) { return fp2_from_pod(p, ok); }
    static C12_HD void store(uint8_t* p, const Fp2& a) { fp2_to_pod(p, a); }
};

// point1 / point2 (projective (X:Y:Z), lazy residues) -> wire-format affine bytes (ECP_affine semantics)
template <class F> C12_HD bool pod_point_to_wire(const uint8_t* pod, uint8_t* wire)
{
    bool ok = true;
    Proj<F> P{Pod<F>::load(pod, ok), Pod<F>::load(pod + Pod<F>::COORD, ok), Pod<F>::load(pod + 2 * Pod<F>::COORD, ok)};
    Wire<F>::serialize(wire, proj_to_affine(P));
    return ok;
}

// wire-format affine bytes -> point1 / point2 with Z = 1 (identity: (0 : 1 : 0), ECP_inf / ECP2_inf)
template <class F> C12_HD bool wire_to_pod_point(const uint8_t* wire, uint8_t* pod)
{
    Affine<F> a;
    bool ok = Wire<F>::parse(a, wire);
    Proj<F> P = proj_from_affine(a);
    Pod<F>::store(pod, P.x);
    Pod<F>::store(pod + Pod<F>::COORD, P.y);
    Pod<F>::store(pod + 2 * Pod<F>::COORD, P.z);
    return ok;
}

// w (< 2^416) mod r by shift-and-subtract; only values >= r pay for it (the DSL hands over scalars in [0, r))
C12_HD void words_mod_r(uint32_t (&w)[13])
{
    const uint32_t r[8] = C12_R_LIMBS;
    bool small = (w[8] | w[9] | w[10] | w[11] | w[12]) == 0;
    if (small) {
        bool lt = false, decided = false;
        for (int i = 7; i >= 0; --i)
            if (!decided && w[i] != r[i]) {
                lt = w[i] < r[i];
                decided = true;
            }
        if (lt) return;
    }
#pragma unroll 1
    for (int sh = 160; sh >= 0; --sh) {            // r < 2^255: r << 160 still fits the 13 words
        uint32_t t[13];
        const int ws = sh >> 5, bs = sh & 31;
        for (int i = 0; i < 13; ++i) {
            const int j = i - ws;
            uint32_t lo = (j >= 0 && j < 8) ? r[j] : 0u, below = (j - 1 >= 0 && j - 1 < 8) ? r[j - 1] : 0u;
            t[i] = bs ? ((lo << bs) | (below >> (32 - bs))) : lo;
        }
        bool ge = true;
        for (int i = 12; i >= 0; --i)
            if (w[i] != t[i]) {
                ge = w[i] > t[i];
                break;
            }
        if (ge) {
            uint64_t borrow = 0;
            for (int i = 0; i < 13; ++i) {
                uint64_t d = (uint64_t)w[i] - t[i] - borrow;
                w[i] = (uint32_t)d;
                borrow = (d >> 32) & 1u;
            }
        }
    }
}

// big -> 32-byte big-endian scalar REDUCED mod r, as PAIR_G1mul / PAIR_G2mul reduce theirs (pair_BLS12381.cpp:878-881,929-932;
// for ECP_muln / ECP_mul2, which do not, the group element is the same); false only for a negative value
C12_HD bool pod_big_to_scalar(const uint8_t* big, uint8_t* out32)
{
    uint32_t w[13];
    bool ok = big_to_words(reinterpret_cast<const long long*>(big), w);
    words_mod_r(w);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint8_t* q = out32 + 28 - 4 * j;
        q[0] = (uint8_t)(w[j] >> 24);
        q[1] = (uint8_t)(w[j] >> 16);
        q[2] = (uint8_t)(w[j] >> 8);
        q[3] = (uint8_t)w[j];
    }
    return ok;
}

// fp12 <-> 576-byte wire value, one of the 12 coefficients per call: the wire order (c, b, a; each Fp4 b, a; each
// Fp2 b, a; fp12_BLS12381.cpp:923-929) is the struct's 12 fp members in exactly reverse order.
C12_HD bool pod_fp12_coeff_to_wire(const uint8_t* pod, uint32_t j, uint8_t* wire)
{
    bool ok = true;
    Fp a = fp_from_pod(pod + (size_t)POD_FP * (11 - j), ok);
    fp_to_be48(wire + 48ull * j, fp_from_mont(a));
    return ok;
}
C12_HD void wire_to_pod_fp12_coeff(const uint8_t* wire, uint32_t j, uint8_t* pod)
{
    Fp a = fp_to_mont(fp_from_be48(wire + 48ull * j));
    fp_to_pod(pod + (size_t)POD_FP * (11 - j), a);
    if (j == 0) {
        *reinterpret_cast<int*>(pod + 768) = 5;  // FP_DENSE (core.h:57-62)
        *reinterpret_cast<int*>(pod + 772) = 0;
    }
}

} // namespace c12
