// Lane-cooperative pairing products: ONE k-pair product (k <= 4 per pass) per group of SIX lanes, state in shared memory.
//
// Why: a thread-per-instance Miller loop carries ~5 KB of live state (f, k G2 accumulators, temporaries).  At the
// 256-512 threads an SM needs to keep its integer pipe busy that is 1.3-2.5 MB per SM - it lives in L2, and the kernel
// is bound by local-memory traffic (measured: 27 % of the integer-multiply peak, and *lower* at higher occupancy,
// profiles/r01g).  Here an instance's state (2,880 B) is shared by six lanes, so 60-75 instances = 11-15 warps per SM
// fit in the 227 KB of shared memory and nothing spills off chip.
//
// Representation: Fp12 = Fp2[w]/(w^6 - (1+i)).  With MIRACL's tower a + b k + c k^2 over Fp4 = Fp2[j]/(j^2 - (1+i)),
// k^3 = j: w = k, and (a, b, c) = ((f0, f3), (f1, f4), (f2, f5)) - the same six Fp2 coefficients, reordered.
// Lane r of a group owns coefficient f_r.
//   sparse line product (dominant): line = l0 + l1 w^3 + l2 w^5 (PAIR_line's M-type form: a = [l0, l1], c = [0, l2]),
//       f_r' = l0 f_r + xi^[r<3] l1 f_(r+3 mod 6) + xi^[r<5] l2 f_(r+1 mod 6)        3 products per lane, no exchange
//   squaring / product: lane r computes the r-th Fp4 product of the Karatsuba / Chung-Hasan schedule (3 Fp2 products),
//       publishes it, then every lane assembles its own coefficient                   3 products per lane + exchange
//   cyclotomic squaring: 6 Fp2 products, one per lane
//   G2 doubling / addition with line evaluation: lane j < k owns pair j (MIRACL's formulas, pair_BLS12381.cpp:40-144;
//       products shared between the line and the doubling have identical values, so they are computed once)
// Values are those of the thread-per-instance bodies in pairing.cuh (field arithmetic is exact; only the evaluation
// order differs), which tests/hostmirror checks against the reference's golden vectors.
#pragma once
#include "pairing.cuh"

namespace c12 {
namespace pc {

constexpr int GROUP = 6;            // lanes per instance
constexpr int INST_PER_WARP = 5;    // lanes 30, 31 idle
constexpr int CHUNK = 4;            // pairs per Miller pass (one lane per pair for the point arithmetic)

struct __align__(16) InstSmem {
    Fp2 f[6];       // the accumulator, coefficient of w^i
    Fp2 x[12];      // exchange area (Fp4 partial products) / line coefficients (4 lines x 3) / Fp12 slots
    Fp2 t[12];      // G2 accumulators T_j = (X, Y, Z), j < 4 / two Fp12 slots during the final exponentiation
};
static_assert(sizeof(InstSmem) == 2880, "shared-memory budget");

struct Lane {
    int r;          // role 0..5
    bool on;        // this lane belongs to a real instance
    InstSmem* s;
};

__device__ __forceinline__ Fp2 ld(const Fp2* p)
{
    Fp2 v;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&v);
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = s[i];
    return v;
}
__device__ __forceinline__ void st(Fp2* p, const Fp2& v, bool on)
{
    if (!on) return;
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = s[i];
}
__device__ __forceinline__ Fp4 ld4(const Fp2* base, int i) { return Fp4{ld(base + i), ld(base + i + 3)}; }   // (f_i, f_(i+3))

// xi^e * v for e in {0, 1}
__device__ __forceinline__ Fp2 mul_ip_if(bool e, const Fp2& v) { return select(e, mul_ip(v), v); }

// ---- f *= line (coefficients at ln[0..2]) ----------------------------------------------------------------------------
__device__ __noinline__ void sparse_mul(Lane L, const Fp2* ln, bool keep)
{
    const int r = L.r;
    Fp2 acc = mul(ld(ln), ld(L.s->f + r));
    acc = add(acc, mul_ip_if(r < 3, mul(ld(ln + 1), ld(L.s->f + (r + 3) % 6))));
    acc = add(acc, mul_ip_if(r < 5, mul(ld(ln + 2), ld(L.s->f + (r + 1) % 6))));
    __syncwarp();
    st(L.s->f + r, acc, L.on && keep);
    __syncwarp();
}

// ---- dst = a * b, general (dst may alias a or b; uses x[] as the exchange area, so a, b, dst must not live there) ---------
// Karatsuba over Fp4 (FP12_mul, fp12_BLS12381.cpp:246-298): lane r computes the r-th Fp4 product
//   v0 = a0 b0, v1 = a1 b1, v2 = a2 b2, m0 = (a1+a2)(b1+b2), m1 = (a0+a1)(b0+b1), m2 = (a0+a2)(b0+b2)
// Operands are re-read from shared memory where they are needed instead of being held in registers: the live set stays
// at two operands plus one partial sum, which is what lets three blocks (12 warps) share an SM's register file.
struct Opnd {           // an Fp4 operand in shared memory: half h = p[h * stride] (+ q[h * stride] when q != nullptr)
    const Fp2* p;
    const Fp2* q;
    int stride;
};
__device__ __forceinline__ Fp2 fetch(const Opnd& o, int h)
{
    Fp2 v = ld(o.p + h * o.stride);
    if (o.q) v = add(v, ld(o.q + h * o.stride));
    return v;
}
// the Fp4 product U V (times two when `twice`) -> x[2 slot], x[2 slot + 1]
__device__ __forceinline__ void fp4_product_to_x(Lane L, int slot, const Opnd& U, const Opnd& V, bool twice, bool store)
{
    Fp2 t0 = mul(fetch(U, 0), fetch(V, 0));
    Fp2 t1 = mul(fetch(U, 1), fetch(V, 1));
    Fp2 lo = add(t0, mul_ip(t1));
    Fp2 s = add(t0, t1);
    if (twice) lo = dbl(lo);
    st(L.s->x + 2 * slot, lo, L.on && store);
    Fp2 ua = add(fetch(U, 0), fetch(U, 1));
    Fp2 ub = add(fetch(V, 0), fetch(V, 1));
    Fp2 hi = sub(mul(ua, ub), s);
    if (twice) hi = dbl(hi);
    st(L.s->x + 2 * slot + 1, hi, L.on && store);
}
__device__ __forceinline__ Fp2 assemble_product(const Fp2* x, int r)
{
    // x[2 q], x[2 q + 1] = the two halves of Fp4 product q (order v0 v1 v2 m0 m1 m2)
    // t0 = m0 - v1 - v2, t1 = m1 - v0 - v1, t2 = m2 - v0 - v2;  a' = v0 + j t0, b' = t1 + j v2, c' = t2 + v1
    // j (y0, y1) = (xi y1, y0).  Coefficients: f0 = a'.lo, f3 = a'.hi, f1 = b'.lo, f4 = b'.hi, f2 = c'.lo, f5 = c'.hi
    const int h = r >= 3 ? 1 : 0;   // which half of the Fp4 this lane owns
    const int q = r % 3;            // a', b', c'
    if (q == 0) {
        // a' = v0 + j (m0 - v1 - v2): lo = v0.lo + xi (m0 - v1 - v2).hi ; hi = v0.hi + (m0 - v1 - v2).lo
        Fp2 t = sub(sub(ld(x + 6 + (1 - h)), ld(x + 2 + (1 - h))), ld(x + 4 + (1 - h)));
        return add(ld(x + h), mul_ip_if(h == 0, t));
    }
    if (q == 1) {
        // b' = (m1 - v0 - v1) + j v2: lo = t1.lo + xi v2.hi ; hi = t1.hi + v2.lo
        Fp2 t = sub(sub(ld(x + 8 + h), ld(x + h)), ld(x + 2 + h));
        return add(t, mul_ip_if(h == 0, ld(x + 4 + (1 - h))));
    }
    // c' = (m2 - v0 - v2) + v1
    Fp2 t = sub(sub(ld(x + 10 + h), ld(x + h)), ld(x + 4 + h));
    return add(t, ld(x + 2 + h));
}
__device__ __noinline__ void full_mul(Lane L, Fp2* dst, const Fp2* a, const Fp2* b)
{
    // r: 0,1,2 -> a_r; 3 -> a1+a2; 4 -> a0+a1; 5 -> a0+a2
    const int r = L.r;
    const int i = r < 3 ? r : (r == 3 ? 1 : 0);
    const int j = r < 3 ? -1 : (r == 4 ? 1 : 2);
    const Opnd U = {a + i, j < 0 ? nullptr : a + j, 3}, V = {b + i, j < 0 ? nullptr : b + j, 3};
    fp4_product_to_x(L, r, U, V, false, true);
    __syncwarp();
    Fp2 c = assemble_product(L.s->x, r);
    __syncwarp();
    st(dst + r, c, L.on);
    __syncwarp();
}

// ---- dst = a^2, general (Chung-Hasan SQR2, FP12_sqr, fp12_BLS12381.cpp:190): five Fp4 products, lane 5 idles ----------
//   s0 = a0^2, s1 = 2 a0 a1, s2 = (a0 - a1 + a2)^2, s3 = 2 a1 a2, s4 = a2^2
//   a' = s0 + j s3, b' = s1 + j s4, c' = s1 + s2 + s3 - s0 - s4
// Lane 2 first parks d = a0 - a1 + a2 in x[10], x[11] (free until the partial products land in x[0..9]).
__device__ __noinline__ void full_sqr(Lane L, Fp2* dst, const Fp2* a)
{
    const int r = L.r;
    if (r == 2) {
        st(L.s->x + 10, add(sub(ld(a), ld(a + 1)), ld(a + 2)), L.on);
        st(L.s->x + 11, add(sub(ld(a + 3), ld(a + 4)), ld(a + 5)), L.on);
    }
    __syncwarp();
    // operands: coefficients of `a` three apart (lanes 0, 1, 3, 4), or the parked d, halves one apart (lane 2)
    const int iu = r == 3 ? 1 : (r >= 4 ? 2 : 0);       // 0: a0, 1: a0, 3: a1, 4: a2
    const int iv = r == 0 ? 0 : (r == 1 ? 1 : 2);       // 0: a0, 1: a1, 3: a2, 4: a2
    const Opnd U = {r == 2 ? L.s->x + 10 : a + iu, nullptr, r == 2 ? 1 : 3};
    const Opnd V = {r == 2 ? L.s->x + 10 : a + iv, nullptr, r == 2 ? 1 : 3};
    fp4_product_to_x(L, r, U, V, r == 1 || r == 3, r < 5);
    __syncwarp();
    const Fp2* x = L.s->x;      // x[2q], x[2q+1] = halves of s_q
    const int h = r >= 3 ? 1 : 0, q = r % 3;
    Fp2 c;
    if (q == 0)
        c = add(ld(x + h), mul_ip_if(h == 0, ld(x + 6 + (1 - h))));                 // s0 + j s3
    else if (q == 1)
        c = add(ld(x + 2 + h), mul_ip_if(h == 0, ld(x + 8 + (1 - h))));             // s1 + j s4
    else
        c = sub(sub(add(add(ld(x + 2 + h), ld(x + 4 + h)), ld(x + 6 + h)), ld(x + h)), ld(x + 8 + h));
    __syncwarp();
    st(dst + r, c, L.on);
    __syncwarp();
}

// ---- s = s^2 on the cyclotomic subgroup (Granger-Scott, FP12_usqr, fp12_BLS12381.cpp:147-187), in place ------------
//   A = a0^2, B = j a2^2, C = a1^2;  a' = 3A + 2 nconj(a0), b' = 3B + 2 conj(a1), c' = 3C + 2 nconj(a2)
//   Fp4 square (y0, y1)^2 = (P' - P - xi P, 2 P) with P = y0 y1, P' = (y0 + y1)(y0 + xi y1): one product per lane.
__device__ __noinline__ void cyclo_sqr(Lane L, Fp2* s)
{
    const int r = L.r;
    // lanes (0,1) square a0 = (f0, f3); lanes (2,3) square a2 = (f2, f5); lanes (4,5) square a1 = (f1, f4)
    const int src = r < 2 ? 0 : (r < 4 ? 2 : 1);
    Fp2 y0 = ld(s + src), y1 = ld(s + src + 3);
    const bool odd = r & 1;
    Fp2 u = select(odd, add(y0, y1), y0);
    Fp2 v = select(odd, add(y0, mul_ip(y1)), y1);
    Fp2 p = mul(u, v);
    Fp2 own = ld(s + r);
    __syncwarp();
    st(L.s->x + r, p, L.on);    // x[0],x[1] = P, P' of a0^2;  x[2],x[3] of a2^2;  x[4],x[5] of a1^2
    __syncwarp();
    const Fp2* x = L.s->x;
    // squares: A = a0^2 = (x1 - x0 - xi x0, 2 x0); S2 = a2^2 = (x3 - x2 - xi x2, 2 x2) -> B = j S2 = (xi 2 x2, S2.lo);
    //          C = a1^2 = (x5 - x4 - xi x4, 2 x4)
    // f0' = 3 A.lo - 2 f0     f3' = 3 A.hi + 2 f3        (nconj(a0) = (-f0, f3))
    // f1' = 3 B.lo + 2 f1     f4' = 3 B.hi - 2 f4        (conj(a1)  = (f1, -f4))
    // f2' = 3 C.lo - 2 f2     f5' = 3 C.hi + 2 f5        (nconj(a2) = (-f2, f5))
    const int pbase = (r == 0 || r == 3) ? 0 : ((r == 1 || r == 4) ? 2 : 4);
    Fp2 P = ld(x + pbase), Pp = ld(x + pbase + 1);
    Fp2 lo = sub(sub(Pp, P), mul_ip(P));    // low half of the square
    Fp2 hi = dbl(P);                        // high half
    Fp2 term;                               // the half of A / B / C this lane needs
    bool plus;
    if (r == 0 || r == 2) { term = lo; plus = false; }            // A.lo, C.lo  (minus 2 f)
    else if (r == 3 || r == 5) { term = hi; plus = true; }        // A.hi, C.hi  (plus 2 f)
    else if (r == 1) { term = mul_ip(hi); plus = true; }          // B.lo = xi * S2.hi
    else { term = lo; plus = false; }                             // r == 4: B.hi = S2.lo
    Fp2 t3 = add(dbl(term), term);
    Fp2 o2 = dbl(own);
    Fp2 c = plus ? add(t3, o2) : sub(t3, o2);
    __syncwarp();
    st(s + r, c, L.on);
    __syncwarp();
}

// conj (p^6 Frobenius): w -> -w
__device__ void conj_inplace(Lane L, Fp2* s)
{
    if (L.r & 1) st(s + L.r, neg(ld(s + L.r)), L.on);
    __syncwarp();
}
__device__ void copy12(Lane L, Fp2* dst, const Fp2* src)
{
    st(dst + L.r, ld(src + L.r), L.on);
    __syncwarp();
}
// x^p (FP12_frob, fp12_BLS12381.cpp:867-881): f_r -> conj(f_r) * gamma_r, gamma = (1, c1, c2, c3, c1 c3, c2 c3)
__device__ __noinline__ void frob_inplace(Lane L, Fp2* s)
{
    const int r = L.r;
    Fp2 c1 = frob_c1_m(), c2 = frob_c2_m(), c3 = frob_c3_m();
    Fp2 g = (r == 1 || r == 4) ? c1 : c2;
    Fp2 v = conj(ld(s + r));
    Fp2 y = mul(v, g);                 // lanes 0, 3 discard it
    y = select(r == 0 || r == 3, v, y);
    Fp2 z = mul(y, c3);
    y = select(r >= 3, z, y);
    st(s + r, y, L.on);
    __syncwarp();
}

// dst = 1 / src: lane 0 of the group runs the scalar FP12_inv body once per instance (one of ~2,000 steps)
__device__ __noinline__ void inv12(Lane L, Fp2* dst, const Fp2* src)
{
    if (L.r == 0 && L.on) {
        Fp12 v = Fp12{Fp4{ld(src + 0), ld(src + 3)}, Fp4{ld(src + 1), ld(src + 4)}, Fp4{ld(src + 2), ld(src + 5)}};
        v = inv(v);
        st(dst + 0, v.a.a, L.on);
        st(dst + 3, v.a.b, L.on);
        st(dst + 1, v.b.a, L.on);
        st(dst + 4, v.b.b, L.on);
        st(dst + 2, v.c.a, L.on);
        st(dst + 5, v.c.b, L.on);
    }
    __syncwarp();
}

// acc = base^|x| for unitary base (pow_x_abs in pairing.cuh); acc, base distinct slots outside x[]
__device__ void pow_x(Lane L, Fp2* acc, const Fp2* base)
{
    const uint64_t e = C12_X_ABS;
    copy12(L, acc, base);
#pragma unroll 1
    for (int i = 62; i >= 0; --i) {
        C12_BLOCK_ALIGN();
        cyclo_sqr(L, acc);
        if ((e >> i) & 1ull) full_mul(L, acc, acc, base);
    }
}

// ---- final exponentiation of s->f in place (PAIR_fexp, pair_BLS12381.cpp:629-755; final_exp in pairing.cuh) ---------
// slots: F = s->f, S1 = s->t[0..5], S2 = s->t[6..11]; g = this instance's six-coefficient global scratch
__device__ void final_exp_coop(Lane L, Fp2* g)
{
    Fp2* F = L.s->f;
    Fp2* S1 = L.s->t;
    Fp2* S2 = L.s->t + 6;
    // easy part
    inv12(L, S1, F);
    conj_inplace(L, F);
    full_mul(L, F, F, S1);          // r = conj(f) / f
    copy12(L, S1, F);
    frob_inplace(L, F);
    frob_inplace(L, F);
    full_mul(L, F, F, S1);          // r = r^(p^2) r
    // hard part: y1 = r^3 (kept in global scratch)
    copy12(L, S1, F);
    cyclo_sqr(L, S1);
    full_mul(L, S1, S1, F);
    if (L.on) g[L.r] = S1[L.r];
    __syncwarp();
    // r = r^(x-1) twice  (x < 0: r^x = conj(r^|x|))
#pragma unroll 1
    for (int rep = 0; rep < 2; ++rep) {
        pow_x(L, S1, F);
        conj_inplace(L, S1);
        conj_inplace(L, F);
        full_mul(L, F, S1, F);
    }
    // r = r^(x+p)
    pow_x(L, S1, F);
    conj_inplace(L, S1);
    frob_inplace(L, F);
    full_mul(L, F, S1, F);
    // y0 = r^(x^2) * r^(p^2) ; r = y0 * conj(r)
    pow_x(L, S1, F);
    pow_x(L, S2, S1);               // S2 = r^(x^2)
    copy12(L, S1, F);
    frob_inplace(L, S1);
    frob_inplace(L, S1);
    full_mul(L, S2, S2, S1);
    conj_inplace(L, F);
    full_mul(L, F, S2, F);
    // times y1
    st(S1 + L.r, L.on ? g[L.r] : ld(S1 + L.r), L.on);
    __syncwarp();
    full_mul(L, F, F, S1);
}

// ---- G2 point steps with line evaluation: lane j < k owns pair j ---------------------------------------------------------
// PAIR_double + PAIR_line (pair_BLS12381.cpp:40-78,119-144): line through T,T evaluated at P = (px, py), then T <- 2T.
__device__ __noinline__ void point_double_line(Fp2* T, Fp2* ln, const Fp& px, const Fp& py, bool on)
{
    Fp2 Y = ld(T + 1), Z = ld(T + 2);
    Fp2 yz = mul(Y, Z);
    st(ln, mul_fp(mul_ip(neg(dbl(yz))), py), on);          // l0 = -2YZ (1+i) * Py
    Fp2 t2 = FieldOps<Fp2>::mul_b3(sqr(Z));                 // 3b' Z^2
    Fp2 yy = sqr(Y);
    st(ln + 1, sub(t2, yy), on);                            // l1 = 3b' Z^2 - Y^2
    Fp2 X = ld(T);
    st(ln + 2, mul_fp(mul3(sqr(X)), px), on);               // l2 = 3X^2 * Px
    Fp2 xy = mul(X, Y);
    Fp2 z3 = mul8(yy);
    st(T + 2, mul(z3, yz), on);                             // Z3
    Fp2 x3 = mul(t2, z3);
    Fp2 y3 = add(yy, t2);
    Fp2 t0 = sub(yy, mul3(t2));
    st(T + 1, add(mul(y3, t0), x3), on);                    // Y3
    st(T, dbl(mul(t0, xy)), on);                            // X3
}

// PAIR_add + PAIR_line (pair_BLS12381.cpp:81-144): line through T and B = +-Q, T <- T + B.  B is what ECP2_affine leaves:
// (x, y, 1), or (0 : +-1 : 0) for the identity - the reference runs the same formulas on it, so do we.
__device__ __noinline__ void point_add_line(Fp2* T, Fp2* ln, const Affine<Fp2>& Q, bool negate, const Fp& px, const Fp& py, bool on)
{
    Proj<Fp2> A = Proj<Fp2>{ld(T), ld(T + 1), ld(T + 2)};
    const bool qinf = affine_is_inf(Q);
    Fp2 bx = Q.x, by = qinf ? fp2_one() : Q.y;
    if (negate) by = neg(by);
    Fp2 x1 = sub(A.x, mul(A.z, bx));
    Fp2 y1 = sub(A.y, mul(A.z, by));
    st(ln, mul_fp(mul_ip(x1), py), on);
    st(ln + 1, sub(mul(y1, bx), mul(x1, by)), on);
    st(ln + 2, mul_fp(neg(y1), px), on);
    Proj<Fp2> R = qinf ? proj_add(A, Proj<Fp2>{bx, by, fp2_zero()}) : proj_add_affine_nz(A, Affine<Fp2>{bx, by});
    st(T, R.x, on);
    st(T + 1, R.y, on);
    st(T + 2, R.z, on);
}

// Parsed inputs of one pair, kept in global scratch between uses (288 B): P (Montgomery affine), Q (Montgomery affine)
struct PairIn {
    Affine<Fp> P;
    Affine<Fp2> Q;
};

// Miller product of pairs [0, kk) of this instance into s->f (un-exponentiated, conjugated).  Lanes j < kk own pair j.
__device__ void miller_coop(Lane L, const PairIn* pin, int kk)
{
    const int j = L.r;
    const bool mine = j < kk && j < CHUNK;
    InstSmem* s = L.s;
    Fp px = fp_zero(), py = fp_zero();
    bool live = false;
    if (mine && L.on) {
        px = pin[j].P.x;
        py = pin[j].P.y;
        live = !(fp_is_zero(px) && fp_is_zero(py));
        Proj<Fp2> T0 = proj_from_affine(pin[j].Q);
        st(s->t + 3 * j, T0.x, true);
        st(s->t + 3 * j + 1, T0.y, true);
        st(s->t + 3 * j + 2, T0.z, true);
    }
    st(s->f + L.r, L.r == 0 ? fp2_one() : fp2_zero(), L.on);
    __syncwarp();
    // which pairs of this instance are live (uniform across its six lanes)
    const unsigned base = (threadIdx.x & 31) / GROUP * GROUP;
    unsigned live_mask = (__ballot_sync(0xffffffffu, live) >> base) & 0xfu;
    const uint64_t pos = 0x1201000000010000ull, negm = 0x4000000000000000ull;
#pragma unroll 1
    for (int i = 64; i >= 1; --i) {
        C12_BLOCK_ALIGN();
        if (i != 64) full_sqr(L, s->f, s->f);     // f = 1 before the first step
        if (mine) point_double_line(s->t + 3 * j, s->x + 3 * j, px, py, L.on);
        __syncwarp();
#pragma unroll 1
        for (int q = 0; q < kk; ++q) sparse_mul(L, s->x + 3 * q, (live_mask >> q) & 1u);
        int bt = (int)((pos >> (i - 1)) & 1ull) - (int)((negm >> (i - 1)) & 1ull);
        if (bt != 0) {
            if (mine) {
                Affine<Fp2> Q = L.on ? pin[j].Q : affine_inf<Fp2>();
                point_add_line(s->t + 3 * j, s->x + 3 * j, Q, bt < 0, px, py, L.on);
            }
            __syncwarp();
#pragma unroll 1
            for (int q = 0; q < kk; ++q) sparse_mul(L, s->x + 3 * q, (live_mask >> q) & 1u);
        }
    }
    conj_inplace(L, s->f);
}

// wire <-> lane-owned coefficient.  FP12_toOctet order (fp12_BLS12381.cpp:923-929): c, b, a; each Fp4 hi then lo; so the
// 96-byte slot of coefficient f_r (r = q + 3 h: Fp4 index q, half h) sits at byte 192 (2 - q) + 96 (1 - h).
__device__ __forceinline__ int wire_offset(int r) { return 192 * (2 - r % 3) + 96 * (1 - r / 3); }

} // namespace pc
} // namespace c12
