// Batched scalar multiplication bodies (one (P, k) per thread), templated on the coordinate field.
// Replaces `multiply(point1&, const big&)` -> PAIR_G1mul and `multiply(point2&, const big&)` -> PAIR_G2mul
// (reference: src/miracl_core_interface.cpp:174-177,202-205; 3rd-party/miracl-core/pair_BLS12381.cpp:876-983)
// and the fixed-base `select` path g^x (include/crypto12381/g1_point.hpp:355-369, g2_point.hpp:129-143).
// The value k*P is independent of the evaluation strategy (MIRACL: GLV/GS + joint windows), outputs are
// compared on normalised encodings.
#pragma once
#include "msm_core.cuh"

namespace c12 {

// r (group order), little-endian limbs
C12_HD bool scalar_is_canonical(const Scalar256& s)
{
    const uint32_t r[8] = C12_R_LIMBS;
    bool lt = false, decided = false;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (!decided && s.v[i] != r[i]) {
            lt = s.v[i] < r[i];
            decided = true;
        }
    }
    return lt;
}

// k * P, 4-bit fixed windows over 256 bits on the complete projective formulas
template <class F> C12_HD Proj<F> proj_scalar_mul(const Affine<F>& p, const Scalar256& k)
{
    if (affine_is_inf(p)) return proj_inf<F>();
    Proj<F> tab[16];
    tab[0] = proj_inf<F>();
    tab[1] = proj_from_affine(p);
    tab[2] = proj_dbl(tab[1]);
#pragma unroll 1
    for (int i = 3; i < 16; ++i) tab[i] = proj_add_affine_nz(tab[i - 1], p);
    Proj<F> acc = proj_inf<F>();
    bool started = false;
#pragma unroll 1
    for (int i = 63; i >= 0; --i) {
        if (started) {
            acc = proj_dbl(acc);
            acc = proj_dbl(acc);
            acc = proj_dbl(acc);
            acc = proj_dbl(acc);
        }
        uint32_t d = (k.v[i >> 3] >> ((i & 7) * 4)) & 15u;
        if (d) {
            acc = started ? proj_add(acc, tab[d]) : tab[d];
            started = true;
        }
    }
    return acc;
}

// k * P for P in the r-torsion subgroup through the scalar split of the MSM (msm_split: G1 two GLV halves, G2 four GLS
// quarters) and a joint double-and-add over the 2^parts - 1 subset sums of the endomorphism images - what the reference
// does with glv() + ECP_mul2 and gs() + ECP2_mul4 (pair_BLS12381.cpp:759-983).  128 (G1) / 64 (G2) doublings instead of 256.
template <class F> C12_HD Proj<F> proj_scalar_mul_split(const Affine<F>& p, const Scalar256& k)
{
    if (affine_is_inf(p)) return proj_inf<F>();
    constexpr uint32_t PARTS = MsmTraits<F>::PARTS, NT = 1u << PARTS, BITS = 256 / PARTS;
    ScalarParts sp;
    msm_split(k, PARTS, sp);
    Affine<F> base[PARTS];
    base[0] = p;
    for (uint32_t q = 1; q < PARTS; ++q) base[q] = MsmTraits<F>::endo(q, p);
    for (uint32_t q = 0; q < PARTS; ++q)
        if (sp.neg[q]) base[q].y = neg(base[q].y);
    Proj<F> tab[NT];
    tab[0] = proj_inf<F>();
#pragma unroll 1
    for (uint32_t m = 1; m < NT; ++m) {
        uint32_t j = 0;
        while (!((m >> j) & 1u)) ++j;
        const uint32_t rest = m ^ (1u << j);
        tab[m] = rest ? proj_add_affine_nz(tab[rest], base[j]) : proj_from_affine(base[j]);
    }
    Proj<F> acc = proj_inf<F>();
    bool started = false;
#pragma unroll 1
    for (int bit = (int)BITS - 1; bit >= 0; --bit) {
        if (started) acc = proj_dbl(acc);
        uint32_t m = 0;
        for (uint32_t q = 0; q < PARTS; ++q) m |= ((sp.mag[q][bit >> 5] >> (bit & 31)) & 1u) << q;
        if (m) {
            acc = started ? proj_add(acc, tab[m]) : tab[m];
            started = true;
        }
    }
    return acc;
}

// affine_out = false: Wire<F>::COMPRESSED bytes; true: Wire<F>::AFFINE bytes
template <class F> C12_HD bool scalar_mul_body(const uint8_t* point_bytes, const uint8_t* scalar_be32, uint8_t* out, bool affine_out = false)
{
    Affine<F> p;
    bool ok = Wire<F>::parse(p, point_bytes);
    Scalar256 k = scalar_from_be32(scalar_be32);
    Proj<F> r = proj_scalar_mul_split(p, k);
    if (affine_out)
        Wire<F>::serialize(out, proj_to_affine(r));
    else
        Wire<F>::compress(out, proj_to_affine(r));
    return ok;
}

// ---- subgroup membership (PAIR_G1member / PAIR_G2member, pair_BLS12381.cpp:1034-1065,1068-1130) --------
// The endomorphism test of Scott (eprint 2021/1130) the reference uses: a point of the curve is in the r-torsion subgroup
// iff the endomorphism acts on it as its eigenvalue:  G1: (beta X, -Y) == [z^2]P,  G2: -psi(Q) == [z]Q,  z = |x|.
// As in the reference the identity is NOT a member, and for G1 a point with [z]P == P (low order) is rejected.
C12_HD Scalar256 scalar_z_abs()
{
    Scalar256 z;
    for (int i = 0; i < 8; ++i) z.v[i] = 0;
    z.v[0] = (uint32_t)(C12_X_ABS & 0xffffffffull);
    z.v[1] = (uint32_t)(C12_X_ABS >> 32);
    return z;
}
C12_HD bool subgroup_member(const Affine<Fp>& P)
{
    if (affine_is_inf(P)) return false;
    const Scalar256 z = scalar_z_abs();
    Proj<Fp> T = proj_scalar_mul(P, z);
    if (proj_eq(proj_from_affine(P), T)) return false;
    T = proj_scalar_mul(proj_to_affine(T), z);                       // [z^2]P
    return proj_eq(proj_from_affine(MsmTraits<Fp>::endo(1, P)), T);
}
C12_HD bool subgroup_member(const Affine<Fp2>& Q)
{
    if (affine_is_inf(Q)) return false;
    Proj<Fp2> T = proj_scalar_mul(Q, scalar_z_abs());                // [z]Q
    return proj_eq(proj_from_affine(MsmTraits<Fp2>::endo(1, Q)), T); // endo(1, Q) = -psi(Q) = [z]Q on G2
}

template <class F> C12_HD Affine<F> generator();
template <> C12_HD Affine<Fp> generator<Fp>() { return Affine<Fp>{g1_gen_x_m(), g1_gen_y_m()}; }
template <> C12_HD Affine<Fp2> generator<Fp2>() { return Affine<Fp2>{g2_gen_x_m(), g2_gen_y_m()}; }

// g^x with g the default generator (ECP_generator / ECP2_generator); affine output bytes
template <class F> C12_HD void fixed_base_body(const uint8_t* scalar_be32, uint8_t* out_affine)
{
    Scalar256 k = scalar_from_be32(scalar_be32);
    Proj<F> r = proj_scalar_mul(generator<F>(), k);
    Wire<F>::serialize(out_affine, proj_to_affine(r));
}

} // namespace c12
