// Short-Weierstrass a = 0 curve arithmetic, templated on the coordinate field F (Fp for G1, Fp2 for G2).
//
// Three representations:
//   Affine<F>  (x, y); the identity is encoded as (0, 0) (never on the curve: b != 0)
//   Proj<F>    homogeneous (X:Y:Z), identity (0:1:0), with the COMPLETE Renes–Costello–Batina formulas — the
//              same ones MIRACL uses (reference: 3rd-party/miracl-core/ecp_BLS12381.cpp:550-588 dbl, :750-812
//              add; ecp2_BLS12381.cpp:358-409, :413-502).  No special cases: used wherever operands can
//              collide (bucket reduction, window combination, partial-sum merges, scalar multiplication).
//   XYZZ<F>    (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ, identity ZZ = 0: cheapest mixed addition
//              (8M + 2S) for the bucket-accumulation hot loop; equal / opposite operands are detected
//              and routed to the doubling / identity branches.
#pragma once
#include "fp2.cuh"

namespace c12 {

// 16-byte alignment: a point gathered from HBM (k_accumulate reads one 96 / 192-byte affine point per addition through the sorted
// indices) moves as 128-bit loads - 6 LDG.E.128 per G1 point instead of 24 LDG.E - and leaves as 128-bit stores
// (-DC12_NO_ALIGN16 restores the packed-u32 layout for the A/B; the sizes are multiples of 16 either way).
#if defined(C12_NO_ALIGN16)
#define C12_POINT_ALIGN
#else
#define C12_POINT_ALIGN alignas(16)
#endif
template <class F> struct C12_POINT_ALIGN Affine {
    F x, y;
};
template <class F> struct C12_POINT_ALIGN Proj {
    F x, y, z;
};
template <class F> struct C12_POINT_ALIGN XYZZ {
    F x, y, zz, zzz;
};

template <class F> C12_HD bool affine_is_inf(const Affine<F>& p) { return is_zero(p.x) && is_zero(p.y); }
template <class F> C12_HD Affine<F> affine_inf() { return Affine<F>{FieldOps<F>::zero(), FieldOps<F>::zero()}; }
template <class F> C12_HD Affine<F> affine_neg(const Affine<F>& p) { return Affine<F>{p.x, neg(p.y)}; }

template <class F> C12_HD Proj<F> proj_inf()
{
    return Proj<F>{FieldOps<F>::zero(), FieldOps<F>::one(), FieldOps<F>::zero()};
}
template <class F> C12_HD bool proj_is_inf(const Proj<F>& p) { return is_zero(p.z); }
template <class F> C12_HD Proj<F> proj_neg(const Proj<F>& p) { return Proj<F>{p.x, neg(p.y), p.z}; }
template <class F> C12_HD Proj<F> proj_from_affine(const Affine<F>& p)
{
    if (affine_is_inf(p)) return proj_inf<F>();
    return Proj<F>{p.x, p.y, FieldOps<F>::one()};
}

// RCB15 Algorithm 7 (a = 0): 12M + 2 mul_b3
template <class F> C12_HD Proj<F> proj_add(const Proj<F>& p, const Proj<F>& q)
{
    F t0 = mul(p.x, q.x);
    F t1 = mul(p.y, q.y);
    F t2 = mul(p.z, q.z);
    F t3 = sub(mul(add(p.x, p.y), add(q.x, q.y)), add(t0, t1));
    F t4 = sub(mul(add(p.y, p.z), add(q.y, q.z)), add(t1, t2));
    F y3 = sub(mul(add(p.x, p.z), add(q.x, q.z)), add(t0, t2));
    t0 = mul3(t0);
    t2 = FieldOps<F>::mul_b3(t2);
    F z3 = add(t1, t2);
    t1 = sub(t1, t2);
    y3 = FieldOps<F>::mul_b3(y3);
    F x3 = mul(y3, t4);
    t2 = mul(t3, t1);
    Proj<F> r;
    r.x = sub(t2, x3);
    y3 = mul(y3, t0);
    t1 = mul(t1, z3);
    r.y = add(y3, t1);
    t0 = mul(t0, t3);
    z3 = mul(z3, t4);
    r.z = add(z3, t0);
    return r;
}

// RCB15 Algorithm 8 (a = 0), q affine and NOT the identity: 11M + 2 mul_b3
template <class F> C12_HD Proj<F> proj_add_affine_nz(const Proj<F>& p, const Affine<F>& q)
{
    F t0 = mul(p.x, q.x);
    F t1 = mul(p.y, q.y);
    F t3 = sub(mul(add(p.x, p.y), add(q.x, q.y)), add(t0, t1));
    F t4 = add(mul(q.y, p.z), p.y);
    F y3 = add(mul(q.x, p.z), p.x);
    t0 = mul3(t0);
    F t2 = FieldOps<F>::mul_b3(p.z);
    F z3 = add(t1, t2);
    t1 = sub(t1, t2);
    y3 = FieldOps<F>::mul_b3(y3);
    F x3 = mul(y3, t4);
    t2 = mul(t3, t1);
    Proj<F> r;
    r.x = sub(t2, x3);
    y3 = mul(y3, t0);
    t1 = mul(t1, z3);
    r.y = add(y3, t1);
    t0 = mul(t0, t3);
    z3 = mul(z3, t4);
    r.z = add(z3, t0);
    return r;
}

template <class F> C12_HD Proj<F> proj_add_affine(const Proj<F>& p, const Affine<F>& q)
{
    if (affine_is_inf(q)) return p;
    return proj_add_affine_nz(p, q);
}

// RCB15 Algorithm 9 (a = 0): 6M + 2S + 1 mul_b3
template <class F> C12_HD Proj<F> proj_dbl(const Proj<F>& p)
{
    F t0 = sqr(p.y);
    F t1 = mul(p.y, p.z);
    F t2 = sqr(p.z);
    F z3 = mul8(t0);
    t2 = FieldOps<F>::mul_b3(t2);
    F x3 = mul(t2, z3);
    F y3 = add(t0, t2);
    Proj<F> r;
    r.z = mul(z3, t1);
    t2 = mul3(t2);
    t0 = sub(t0, t2);
    y3 = mul(y3, t0);
    r.y = add(y3, x3);
    t1 = mul(p.x, p.y);
    r.x = dbl(mul(t0, t1));
    return r;
}

// (X:Y:Z) -> affine with ONE inversion; identity -> (0,0)
template <class F> C12_HD Affine<F> proj_to_affine(const Proj<F>& p)
{
    if (proj_is_inf(p)) return affine_inf<F>();
    F zi = inv(p.z);
    return Affine<F>{mul(p.x, zi), mul(p.y, zi)};
}

// projective equality without inversion (ECP_equals, ecp_BLS12381.cpp:105-131)
template <class F> C12_HD bool proj_eq(const Proj<F>& p, const Proj<F>& q)
{
    return eq(mul(p.x, q.z), mul(q.x, p.z)) && eq(mul(p.y, q.z), mul(q.y, p.z));
}

// ---- XYZZ ------------------------------------------------------------------------------------------------
template <class F> C12_HD XYZZ<F> xyzz_inf()
{
    F z = FieldOps<F>::zero();
    return XYZZ<F>{z, z, z, z};
}
template <class F> C12_HD bool xyzz_is_inf(const XYZZ<F>& p) { return is_zero(p.zz); }

// affine doubling into XYZZ (mdbl-2008-s-1), q != identity, y != 0 always on these curves (odd order)
template <class F> C12_HD XYZZ<F> xyzz_dbl_affine(const Affine<F>& q)
{
    F u = dbl(q.y);
    F v = sqr(u);
    F w = mul(u, v);
    F s = mul(q.x, v);
    F m = mul3(sqr(q.x));
    XYZZ<F> r;
    r.x = sub(sqr(m), dbl(s));
    r.y = sub(mul(m, sub(s, r.x)), mul(w, q.y));
    r.zz = v;
    r.zzz = w;
    return r;
}

// acc += q (madd-2008-s), q affine and not the identity.  8M + 2S on the common path.
template <class F> C12_HD void xyzz_madd(XYZZ<F>& acc, const Affine<F>& q)
{
    if (xyzz_is_inf(acc)) {
        acc.x = q.x;
        acc.y = q.y;
        acc.zz = FieldOps<F>::one();
        acc.zzz = FieldOps<F>::one();
        return;
    }
    F u2 = mul_hot(q.x, acc.zz);
    F s2 = mul_hot(q.y, acc.zzz);
    F pp = sub(u2, acc.x);
    F rr = sub(s2, acc.y);
    if (is_zero(pp)) {
        if (is_zero(rr))
            acc = xyzz_dbl_affine(q);  // same point
        else
            acc = xyzz_inf<F>();       // opposite points
        return;
    }
    F p2 = sqr_hot(pp);
    F p3 = mul_hot(pp, p2);
    F qq = mul_hot(acc.x, p2);
    F x3 = sub(sub(sqr_hot(rr), p3), dbl(qq));
    acc.y = sub(mul_hot(rr, sub(qq, x3)), mul_hot(acc.y, p3));
    acc.x = x3;
    acc.zz = mul_hot(acc.zz, p2);
    acc.zzz = mul_hot(acc.zzz, p3);
}

// XYZZ -> homogeneous projective: (X*ZZZ : Y*ZZ : ZZ*ZZZ)
template <class F> C12_HD Proj<F> xyzz_to_proj(const XYZZ<F>& p)
{
    if (xyzz_is_inf(p)) return proj_inf<F>();
    return Proj<F>{mul(p.x, p.zzz), mul(p.y, p.zz), mul(p.zz, p.zzz)};
}

// k * P for a small non-negative integer k (bucket-reduction weights), double-and-add on complete formulas
template <class F> C12_HD Proj<F> proj_mul_small(const Proj<F>& p, uint32_t k)
{
    if (k == 0) return proj_inf<F>();
    int top = 31;
    while (!((k >> top) & 1u)) --top;
    Proj<F> r = p;
#pragma unroll 1
    for (int i = top - 1; i >= 0; --i) {
        r = proj_dbl(r);
        if ((k >> i) & 1u) r = proj_add(r, p);
    }
    return r;
}

} // namespace c12
