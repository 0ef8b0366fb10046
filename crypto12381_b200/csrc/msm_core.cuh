// Multi-scalar multiplication: signed-window Pippenger bodies, templated on the coordinate field
// (Fp -> G1, Fp2 -> G2).  Replaces the reference's MSM seam `sum_of_products` -> ECP_muln (4-bit unsigned
// windows, complete adds in input order; src/miracl_core_interface.cpp:134-137,
// 3rd-party/miracl-core/ecp_BLS12381.cpp:1112-1148) and the live DSL loop of ECP_mul2 calls
// (include/crypto12381/g1_point.hpp:371-404); for G2 the per-term PAIR_G2mul + ECP2_add loop
// (g2_point.hpp:202-236).  Outputs are compared on normalised affine encodings, so any evaluation order of
// the same group sum is bit-exact.
//
// Every function here is a per-thread BODY (host+device) so the indexing and the algorithm can be executed
// serially by tests/hostmirror on the CPU; the __global__ wrappers live in msm_impl.cuh.
//
// Plan for n caller terms with window width c:
//   scalar split : k -> `parts` signed pieces of 256 / parts bits (G1: 2 GLV halves, G2: 4 GLS quarters; msm_split), piece q
//                  of term i is pipeline term q n + i and its point the endomorphism image of P_i (MsmTraits<F>::endo)
//   digit recode : piece = sum_w d_w 2^(c w), d_w in [-2^(c-1)+1, 2^(c-1)], W = ceil(256 / parts / c) windows
//   key          : window-local |d_w| - 1, stored at [w*n' + idx]; value = term index | sign << 31; zero digits get
//                  key = 2^(c-1) (owns no bucket)
//   (segmented radix sort by key, one segment per window; bucket b = w*2^(c-1) + key owns sorted[start[b] .. end[b]))
//   accumulate   : XYZZ mixed additions of the affine terms, one thread per chunk of a bucket's list, partials folded
//   reduce       : per window sum_k k * B_k by multi-level segmented running sums
//   combine      : Horner over windows, c doublings per step
#pragma once
#include "ec.cuh"

namespace c12 {

constexpr uint32_t MSM_MAX_PLANES = 16;      // bit planes of the segment index: at most 2^15 buckets per window / 1 per segment
constexpr uint32_t MSM_WPART_SLOTS = MSM_MAX_PLANES + 1;

struct MsmPlan {
    uint32_t n;           // terms fed to the bucket pipeline = parts * n_in
    uint32_t n_in;        // caller's terms
    uint32_t parts;       // 1: none; 2: GLV halves of < 2^127 (G1); 4: GLS quarters of < 2^63 (G2).  Windows cover 256 / parts bits
    uint32_t c;           // window bits
    uint32_t windows;     // W
    uint32_t half;        // 2^(c-1) buckets per window
    uint32_t total;       // W * half
    uint32_t seg_len;     // buckets per level-0 reduction thread (a power of two)
    uint32_t segs;        // segments per window = ceil(half / seg_len)
    uint32_t plane_bits;  // bits of a segment index: the segment totals are summed once per bit plane (see "bucket reduction")
    // bucket lists longer than `chunk` entries are accumulated by several threads (one per chunk) and folded afterwards:
    // bounds the serial chain of a thread under skewed scalars and evens out the load at large n
    uint32_t chunk;       // entries per accumulation thread, a power of two in [32, 256]
    uint32_t vmax;        // upper bound on the number of chunks (virtual buckets): total + n W / chunk
    // upload groups (msm_list_plan): the terms are cut into `groups` runs of n_group consecutive terms, and every
    // (group, window) pair is a sort segment with buckets of its own - "virtual windows" group * real_windows + w - so the
    // bucket lists of a group need only that group's points (the host entry uploads the points group by group while the
    // earlier groups are already being added up).  The groups' buckets are merged after the accumulation (k_fold).
    uint32_t groups, n_group, real_windows;
};

// Window width for n pipeline terms of `bits`-bit pieces (256: unsplit scalars < r; 128: GLV halves; 64: GLS quarters),
// c in [4, 16] so window-local keys always fit two 8-bit radix passes.  W c >= bits and magnitudes < 2^(bits-1) guarantee
// the signed recoding never carries out of the top window.  The choice minimises a TIME model fitted to the measured
// phases (profiles/r01x_sweep_probe.txt), in microseconds:
//   accumulation  max( W n * 0.39 ns  [integer pipe, full occupancy],  chain * 6.5 us  [one thread's dependent additions] )
//                 with chain = entries of the fullest buckets, at most one chunk (see MsmPlan::chunk)
//   tail          ~1.05 ms of latency-bound reduction / Horner, growing with the bucket count (1.38 ms at c = 16)
// The top window sees only  t = (bits - 1) - c (W - 1)  bits of the magnitudes: widths with t < c / 2 would send every
// term of that window to a handful of buckets and are skipped.
// Measured correction (profiles/r03c_window_sweep_*.txt, r03d): from 2^17 (GLV halves) / 2^16 (GLS quarters) pipeline terms on the
// widest window wins outright over the split scalars - its lists are a handful of entries, so the accumulation is no longer one thread's chain of 32+ dependent
// additions (G1 n = 2^16: 1.96 -> 1.74 ms; G2 n = 2^16: 6.40 -> 3.63 ms).
inline uint32_t msm_choose_window(uint64_t n, uint32_t bits = 256)
{
    if (bits < 256 && n >= (bits <= 64 ? (1ull << 16) : (1ull << 17))) return 16;
    uint32_t best = 4;
    double best_cost = 1e300;
    for (uint32_t c = 4; c <= 16; ++c) {
        const uint32_t W = (bits + c - 1) / c;
        const uint32_t t = (bits - 1) - c * (W - 1);
        if (2 * t < c && c != 4) continue;
        const double load = (double)n / (double)(1u << ((t < c - 1 ? t : c - 1)));     // fullest buckets
        double chunk = 32;
        while (chunk < 256 && chunk < 2.0 * (double)n / (double)(1u << (c - 1))) chunk *= 2;
        const double chain = load < chunk ? load : chunk;
        const double thr = (double)W * (double)n * 0.39e-3, lat = chain * 6.5;
        const double tail = 1050.0 + 330.0 * (double)(1u << (c - 1)) / 32768.0 + (W > 16 ? 10.0 * (W - 16) : 0.0);
        const double cost = (thr > lat ? thr : lat) + tail;
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}

// fills seg_len / segs / plane_bits for a segment length (a power of two <= half)
inline void msm_plan_levels(MsmPlan& pl, uint32_t seg_len)
{
    pl.seg_len = seg_len;
    pl.segs = (pl.half + seg_len - 1) / seg_len;
    pl.plane_bits = 0;
    while ((1u << pl.plane_bits) < pl.segs) ++pl.plane_bits;
}

// n_in caller terms; every scalar is split into `parts` (1, 2 or 4) signed pieces of 256 / parts bits
inline MsmPlan msm_make_plan(uint32_t n_in, uint32_t c, uint32_t parts = 1, uint32_t seg_wave = 0)
{
    MsmPlan pl;
    pl.n_in = n_in;
    pl.parts = parts;
    pl.n = parts * n_in;
    pl.c = c;
    pl.windows = (256u / parts + c - 1) / c;
    pl.half = 1u << (c - 1);
    pl.total = pl.windows * pl.half;
    // level-0 segment length: a power of two in [4, 64] that keeps the level-0 grid within one wave (~48 Ki threads)
#ifndef C12_MSM_SEG_WAVE
#define C12_MSM_SEG_WAVE 49152u
#endif
    if (!seg_wave) seg_wave = C12_MSM_SEG_WAVE;
    uint32_t seg = 2;
    while (seg < 64 && pl.total / seg > seg_wave) seg <<= 1;
    if (seg > pl.half) seg = pl.half;
    msm_plan_levels(pl, seg);
    const uint64_t N = (uint64_t)pl.n * pl.windows;
    uint32_t chunk = 32;
    while (chunk < 256 && (uint64_t)chunk * pl.total < 2 * N) chunk <<= 1;      // about twice the mean bucket load
    pl.chunk = chunk;
    pl.vmax = pl.total + (uint32_t)(N / chunk) + 1;
    pl.groups = 1;
    pl.n_group = n_in;
    pl.real_windows = pl.windows;
    return pl;
}

// the plan of the LIST side of the pipeline (recode, sort, bucket bounds, halving rounds, accumulation) for `groups` upload
// groups: windows / total / n describe the virtual windows and the entries per sort segment; the reduction keeps the plain plan
inline MsmPlan msm_list_plan(const MsmPlan& pl, uint32_t groups)
{
    MsmPlan lp = pl;
    if (groups <= 1) return lp;
    lp.groups = groups;
    lp.n_group = (pl.n_in + groups - 1) / groups;
    lp.n = pl.parts * lp.n_group;
    lp.windows = pl.windows * groups;
    lp.total = lp.windows * pl.half;
    const uint64_t N = (uint64_t)lp.n * lp.windows;
    uint32_t chunk = 32;
    while (chunk < 256 && (uint64_t)chunk * lp.total < 2 * N) chunk <<= 1;
    lp.chunk = chunk;
    lp.vmax = lp.total + (uint32_t)(N / chunk) + 1;
    return lp;
}

C12_HD uint32_t msm_invalid_key(const MsmPlan& pl) { return pl.half; }

// scalars arrive as 32-byte big-endian; limbs little-endian
struct Scalar256 {
    uint32_t v[8];
};

C12_HD Scalar256 scalar_from_be32(const uint8_t* b)
{
    Scalar256 s;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint8_t* q = b + 28 - 4 * i;
        s.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    return s;
}

// c bits starting at bit position `pos` (pos + c may run past 256: zero-extended)
C12_HD uint32_t scalar_bits(const Scalar256& s, uint32_t pos, uint32_t c)
{
    uint32_t limb = pos >> 5, off = pos & 31u;
    if (limb >= 8) return 0;
    uint64_t lo = s.v[limb];
    uint64_t hi = (limb + 1 < 8) ? s.v[limb + 1] : 0;
    uint64_t x = (lo | (hi << 32)) >> off;
    return (uint32_t)(x & ((1ull << c) - 1ull));
}

// same on an L-limb magnitude
template <int L> C12_HD uint32_t limbs_bits(const uint32_t (&v)[L], uint32_t pos, uint32_t c)
{
    uint32_t limb = pos >> 5, off = pos & 31u;
    if (limb >= (uint32_t)L) return 0;
    uint64_t lo = v[limb];
    uint64_t hi = (limb + 1 < (uint32_t)L) ? v[limb + 1] : 0;
    uint64_t x = (lo | (hi << 32)) >> off;
    return (uint32_t)(x & ((1ull << c) - 1ull));
}

// ---- GLV split for G1 (replaces MIRACL's glv(), 3rd-party/miracl-core/pair_BLS12381.cpp:759-810) -----------------
// mu = x^2 (128 bits) acts on G1 as  mu * (X, Y) = (beta X, -Y)  with beta = CRu (rom_field_BLS12381.cpp:58), and
// mu^2 = mu - 1 (mod r) because r = x^4 - x^2 + 1.  k = k0 + k1 mu is rebalanced to |k0|, |k1| <= mu/2 + 1 < 2^127:
//   k0 > mu/2:  k0 -= mu, k1 += 1;      k1 > mu/2:  k1 -= mu - 1, k0 -= 1      (second step uses mu^2 = mu - 1)
// so that eight 16-bit signed windows cover each half with no carry out of the top window.
struct GlvHalves {
    uint32_t a0[4], a1[4];   // magnitudes
    uint32_t neg0, neg1;     // signs
};

namespace detail {
C12_HD bool u128_gt(const uint32_t (&a)[4], const uint32_t (&b)[4])
{
    for (int i = 3; i >= 0; --i)
        if (a[i] != b[i]) return a[i] > b[i];
    return false;
}
C12_HD void u128_sub(uint32_t (&r)[4], const uint32_t (&a)[4], const uint32_t (&b)[4])  // a >= b
{
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t d = (uint64_t)a[i] - b[i] - borrow;
        r[i] = (uint32_t)d;
        borrow = (d >> 32) & 1u;
    }
}
C12_HD void u128_inc(uint32_t (&a)[4])
{
    for (int i = 0; i < 4; ++i)
        if (++a[i] != 0) return;
}
C12_HD void u128_dec(uint32_t (&a)[4])  // a > 0
{
    for (int i = 0; i < 4; ++i)
        if (a[i]-- != 0) return;
}
C12_HD bool u128_is_zero(const uint32_t (&a)[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
} // namespace detail

// k = q mu + rem, 0 <= rem < mu = x^2 (q < mu because r < mu^2)
C12_HD void divmod_mu(const Scalar256& k, uint32_t (&q_out)[4], uint32_t (&rem_out)[4])
{
    const uint32_t mu[4] = C12_X2_LIMBS, rc[5] = C12_X2_RECIP_LIMBS;
    // q = floor(k * floor(2^256 / mu) / 2^256)  (<= floor(k / mu), short by at most 2)
    uint32_t prod[13];
    for (int i = 0; i < 13; ++i) prod[i] = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t carry = 0;
        for (int j = 0; j < 5; ++j) {
            uint64_t t = (uint64_t)k.v[i] * rc[j] + prod[i + j] + carry;
            prod[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        prod[i + 5] = (uint32_t)carry;
    }
    uint32_t q[5];
    for (int i = 0; i < 5; ++i) q[i] = prod[8 + i];
    // rem = k - q * mu  (fits 5 limbs: < 3 mu)
    uint32_t t[9];
    for (int i = 0; i < 9; ++i) t[i] = 0;
    for (int i = 0; i < 5; ++i) {
        uint64_t carry = 0;
        for (int j = 0; j < 4; ++j) {
            uint64_t x = (uint64_t)q[i] * mu[j] + t[i + j] + carry;
            t[i + j] = (uint32_t)x;
            carry = x >> 32;
        }
        if (i + 4 < 9) t[i + 4] = (uint32_t)carry;
    }
    uint32_t rem[5];
    {
        uint64_t borrow = 0;
        for (int i = 0; i < 5; ++i) {
            uint64_t d = (uint64_t)k.v[i] - t[i] - borrow;
            rem[i] = (uint32_t)d;
            borrow = (d >> 32) & 1u;
        }
    }
    for (int guard = 0; guard < 4; ++guard) {  // while rem >= mu
        bool ge = rem[4] != 0;
        if (!ge) {
            uint32_t r4[4] = {rem[0], rem[1], rem[2], rem[3]};
            ge = !detail::u128_gt(mu, r4);
        }
        if (!ge) break;
        uint64_t borrow = 0;
        for (int i = 0; i < 5; ++i) {
            uint64_t d = (uint64_t)rem[i] - (i < 4 ? mu[i] : 0u) - borrow;
            rem[i] = (uint32_t)d;
            borrow = (d >> 32) & 1u;
        }
        for (int i = 0; i < 5; ++i)
            if (++q[i] != 0) break;
    }
    for (int i = 0; i < 4; ++i) {
        q_out[i] = q[i];
        rem_out[i] = rem[i];
    }
}

C12_HD GlvHalves glv_split(const Scalar256& k)
{
    const uint32_t mu[4] = C12_X2_LIMBS;
    uint32_t k0[4], k1[4];
    divmod_mu(k, k1, k0);
    GlvHalves h;
    uint32_t half[4];  // mu / 2
    for (int i = 0; i < 4; ++i) half[i] = (mu[i] >> 1) | (i < 3 ? (mu[i + 1] << 31) : 0u);
    h.neg0 = 0;
    if (detail::u128_gt(k0, half)) {
        detail::u128_sub(k0, mu, k0);   // k0 - mu = -(mu - k0)
        h.neg0 = 1;
        detail::u128_inc(k1);
    }
    h.neg1 = 0;
    if (detail::u128_gt(k1, half)) {
        // k1 <- k1 - (mu - 1),  k0 <- k0 - 1
        uint32_t m1[4] = {mu[0], mu[1], mu[2], mu[3]};
        detail::u128_dec(m1);
        if (detail::u128_gt(k1, m1)) {
            detail::u128_sub(k1, k1, m1);
        } else {
            detail::u128_sub(k1, m1, k1);
            h.neg1 = detail::u128_is_zero(k1) ? 0u : 1u;
        }
        if (h.neg0) {
            detail::u128_inc(k0);
        } else if (detail::u128_is_zero(k0)) {
            k0[0] = 1;
            h.neg0 = 1;
        } else {
            detail::u128_dec(k0);
        }
    }
    for (int i = 0; i < 4; ++i) {
        h.a0[i] = k0[i];
        h.a1[i] = k1[i];
    }
    return h;
}

// ---- GLS split for G2 (replaces MIRACL's gs(), 3rd-party/miracl-core/pair_BLS12381.cpp:814-873) --------------------
// On G2 the endomorphism psi (conjugate the coordinates, scale by f^-2 / f^-3; ECP2_frob, ecp2_BLS12381.cpp:579-590)
// acts as multiplication by p = x = -z (mod r), z = |x| (64 bits).  k = k0 + k1 z + k2 z^2 + k3 z^3 (digits < z, since
// r < z^4) is rebalanced to |k_i| <= z/2 + 1 < 2^63:  k_i > z/2: k_i -= z, k_(i+1) += 1;  for the top digit the carry is
// z^4 = z^2 - 1 (mod r): k2 += 1, k0 -= 1.  Then  k P = sum_i k_i [z^i]P  with  [z^i]P = (-1)^i psi^i(P), so four signed
// 16-bit windows cover each quarter and the Horner recombination needs 48 doublings instead of 240.
struct ScalarParts {
    uint32_t mag[4][4];   // magnitudes, little-endian limbs (upper limbs zero for the shorter pieces)
    uint32_t neg[4];      // signs
};

C12_HD void gls_split(const Scalar256& k, ScalarParts& out)
{
    uint32_t hi[4], lo[4];
    divmod_mu(k, hi, lo);                     // k = hi z^2 + lo, both < z^2
    const unsigned __int128 z = (unsigned __int128)C12_X_ABS;
    const unsigned __int128 L = ((unsigned __int128)lo[3] << 96) | ((unsigned __int128)lo[2] << 64) | ((unsigned __int128)lo[1] << 32) | lo[0];
    const unsigned __int128 H = ((unsigned __int128)hi[3] << 96) | ((unsigned __int128)hi[2] << 64) | ((unsigned __int128)hi[1] << 32) | hi[0];
    __int128 d[4] = {(__int128)(L % z), (__int128)(L / z), (__int128)(H % z), (__int128)(H / z)};
    const __int128 zi = (__int128)z, half = zi >> 1;
    for (int i = 0; i < 3; ++i)
        if (d[i] > half) {
            d[i] -= zi;
            d[i + 1] += 1;
        }
    if (d[3] > half) {
        d[3] -= zi;
        d[2] += 1;
        d[0] -= 1;
    }
    for (int i = 0; i < 4; ++i) {
        const bool ng = d[i] < 0;
        const uint64_t m = (uint64_t)(ng ? -d[i] : d[i]);
        out.mag[i][0] = (uint32_t)m;
        out.mag[i][1] = (uint32_t)(m >> 32);
        out.mag[i][2] = 0;
        out.mag[i][3] = 0;
        out.neg[i] = ng ? 1u : 0u;
    }
}

// the `parts` signed pieces of k (parts = 2: GLV halves, 4: GLS quarters)
C12_HD void msm_split(const Scalar256& k, uint32_t parts, ScalarParts& out)
{
    if (parts == 4) {
        gls_split(k, out);
        return;
    }
    GlvHalves h = glv_split(k);
    for (int i = 0; i < 4; ++i) {
        out.mag[0][i] = h.a0[i];
        out.mag[1][i] = h.a1[i];
        out.mag[2][i] = 0;
        out.mag[3][i] = 0;
    }
    out.neg[0] = h.neg0;
    out.neg[1] = h.neg1;
    out.neg[2] = out.neg[3] = 0;
}

// Scalar splits: the pipeline also needs the endomorphism images of every input point.
//   G1, 2 pieces (GLV):  [mu]P = (beta X, -Y)
//   G2, 4 pieces (GLS):  [z^q]P = (-1)^q psi^q(P)
template <class F> struct MsmTraits;
template <> struct MsmTraits<Fp> {
    static constexpr uint32_t PARTS = 2;
    static C12_HD Affine<Fp> endo(uint32_t, const Affine<Fp>& p) { return Affine<Fp>{fp_mul(p.x, fp_beta_m()), fp_neg(p.y)}; }
};
template <> struct MsmTraits<Fp2> {
    static constexpr uint32_t PARTS = 4;
    static C12_HD Affine<Fp2> endo(uint32_t q, const Affine<Fp2>& p)
    {
        if (q == 2) return Affine<Fp2>{mul_fp(p.x, psi2_x_m()), mul_fp(p.y, psi2_y_m())};
        Fp2 cx = q == 1 ? psi_x_m() : psi3_x_m(), cy = q == 1 ? psi_y_m() : psi3_y_m();
        return Affine<Fp2>{mul(conj(p.x), cx), neg(mul(conj(p.y), cy))};   // odd powers carry the sign of x = -z
    }
};

// Recoding of term i with scalar s: emit(segment, out index, digit magnitude d, value) once per (piece, window); d == 0 means
// the term owns no bucket in that window.  Out index = segment * n + (place of the pipeline term inside its group's segments):
// the w-major layout keeps each window's entries together.  With a split on, term i yields `parts` pipeline terms: piece q is
// pipeline term q n_in + i (point [mu^q]P_i resp. [z^q]P_i); value = pipeline term | sign << 31.
template <class Emit> C12_HD void msm_recode_each(const MsmPlan& pl, uint32_t i, const Scalar256& s, Emit& emit)
{
    // segment of (group, window w) = virtual window group * real_windows + w; the term's place inside its group's segments
    const uint32_t grp = i / pl.n_group, li = i - grp * pl.n_group;
    const uint32_t seg0 = grp * pl.real_windows;
    if (pl.parts > 1) {
        ScalarParts sp;
        msm_split(s, pl.parts, sp);
        for (uint32_t part = 0; part < pl.parts; ++part) {
            const uint32_t sgn = sp.neg[part];
            const uint32_t idx = i + part * pl.n_in;
            uint32_t carry = 0;
            for (uint32_t w = 0; w < pl.real_windows; ++w) {
                uint32_t d = limbs_bits<4>(sp.mag[part], w * pl.c, pl.c) + carry;
                uint32_t neg = 0;
                carry = 0;
                if (d > pl.half) {
                    d = (1u << pl.c) - d;
                    neg = 1;
                    carry = 1;
                }
                emit(seg0 + w, (uint64_t)(seg0 + w) * pl.n + (uint64_t)part * pl.n_group + li, d, idx | ((neg ^ sgn) << 31));
            }
        }
        return;
    }
    uint32_t carry = 0;
    for (uint32_t w = 0; w < pl.real_windows; ++w) {
        uint32_t d = scalar_bits(s, w * pl.c, pl.c) + carry;
        uint32_t neg = 0;
        carry = 0;
        if (d > pl.half) {  // d in (2^(c-1), 2^c] -> d - 2^c in (-2^(c-1), 0]
            d = (1u << pl.c) - d;
            neg = 1;
            carry = 1;
        }
        emit(seg0 + w, (uint64_t)(seg0 + w) * pl.n + li, d, i | (neg << 31));
    }
}

// Inverse of msm_recode_each's layout: the pipeline term (piece q of caller's term i is q n_in + i) whose entry sits at out index o.
C12_HD uint32_t msm_entry_term(const MsmPlan& pl, uint64_t o)
{
    const uint32_t seg = (uint32_t)(o / pl.n), within = (uint32_t)(o - (uint64_t)seg * pl.n);
    const uint32_t part = within / pl.n_group, li = within - part * pl.n_group, grp = seg / pl.real_windows;
    return grp * pl.n_group + li + part * pl.n_in;
}
// The entry of the COUNTING front end: bucket inside the window | sign << 31, all ones when the term owns no bucket there.
constexpr uint32_t MSM_NO_BUCKET = 0xffffffffu;
C12_HD uint32_t msm_count_key(uint32_t d, uint32_t val) { return d ? ((d - 1) | (val & 0x80000000u)) : MSM_NO_BUCKET; }

// Body of the recode kernel of the SORTED front end for term i: (key, value) pairs, key = bucket inside the window.
struct RecodeToPairs {
    uint32_t invalid;
    uint32_t* keys;
    uint32_t* vals;
    C12_HD void operator()(uint32_t, uint64_t o, uint32_t d, uint32_t val)
    {
        keys[o] = d ? d - 1 : invalid;
        vals[o] = d ? val : (val & 0x7fffffffu);
    }
};
C12_HD void msm_recode_body(const MsmPlan& pl, uint32_t i, const uint8_t* scalars_be32, uint32_t* keys, uint32_t* vals)
{
    RecodeToPairs emit{msm_invalid_key(pl), keys, vals};
    msm_recode_each(pl, i, scalar_from_be32(scalars_be32 + 32ull * i), emit);
}

// the unused tail of the last group's segments (n_in is not a multiple of the group count): entries that own no bucket
C12_HD void msm_recode_pad_body(const MsmPlan& pl, uint32_t i, uint32_t* keys, uint32_t* vals)
{
    const uint32_t grp = i / pl.n_group, li = i - grp * pl.n_group;
    for (uint32_t part = 0; part < pl.parts; ++part)
        for (uint32_t w = 0; w < pl.real_windows; ++w) {
            const uint64_t o = ((uint64_t)grp * pl.real_windows + w) * pl.n + (uint64_t)part * pl.n_group + li;
            keys[o] = msm_invalid_key(pl);
            vals[o] = 0;
        }
}

// Body of the bucket-accumulation kernel for one chunk [lo, hi) of a bucket's sorted list: sum of its (signed) affine terms.
template <class F> C12_HD Proj<F> msm_accumulate_range_body(uint32_t lo, uint32_t hi, const uint32_t* vals, const Affine<F>* points)
{
    XYZZ<F> acc = xyzz_inf<F>();
#pragma unroll 1
    for (uint32_t j = lo; j < hi; ++j) {
        uint32_t v = vals[j];
        Affine<F> pt = points[v & 0x7fffffffu];
        if (affine_is_inf(pt)) continue;
        if (v >> 31) pt.y = neg(pt.y);
        xyzz_madd(acc, pt);
    }
    return xyzz_to_proj(acc);
}
template <class F>
C12_HD Proj<F> msm_accumulate_body(uint32_t b, const uint32_t* start, const uint32_t* end, const uint32_t* vals,
                                   const Affine<F>* points)
{
    return msm_accumulate_range_body<F>(start[b], end[b], vals, points);
}
// chunks of bucket b (at least one, so that every bucket owns a partial - possibly the identity)
C12_HD uint32_t msm_chunks(uint32_t m, uint32_t chunk) { return m <= chunk ? 1u : (m + chunk - 1) / chunk; }
// bucket b from its n partial sums (in chunk order)
template <class F> C12_HD Proj<F> msm_fold_body(const Proj<F>* partials, uint32_t n)
{
    Proj<F> acc = partials[0];
#pragma unroll 1
    for (uint32_t j = 1; j < n; ++j) acc = proj_add(acc, partials[j]);
    return acc;
}

// ---- batch-affine halving rounds of the bucket lists -----------------------------------------------------------------
// The bucket sums are formed by AFFINE additions whose inversions are shared by Montgomery's trick.  Round r halves every
// bucket's list: entries (2i, 2i+1) are added, an odd last entry is carried over, so after r rounds a list of m entries holds
// ceil(m / 2^r) partial sums - the same group sum whatever the pairing order, and the outputs are compared on normalised
// encodings.  An affine addition is 1 product for the running denominator product, 2 to unwind it, and lambda, lambda^2, y3:
// 6 Fp products against 10 for the XYZZ mixed addition.
//   lists of round r     : bucket b owns [off_r[b], off_r[b] + len_r(b)) of the round's point array, len_r(b) = ceil(m_b / 2^r);
//                          round 0 is the sorted (key, term) array itself (off_0 = start[], entries fetched through vals -> points)
//   output slot t of a round: bucket b = the one with off_(r+1)[b] <= t < off_(r+1)[b + 1], i = t - off_(r+1)[b],
//                          inputs off_r[b] + 2i and (if 2i + 1 < len_r(b)) + 2i + 1
// Slots are a FLAT index space: every thread of the round's kernels does the same amount of work whatever the bucket sizes are
// (msm_impl.cuh: k_ba_fwd / k_ba_inv / k_ba_bwd).
C12_HD uint32_t ba_len(uint32_t m0, uint32_t r) { return (uint32_t)(((uint64_t)m0 + ((1ull << r) - 1ull)) >> r); }

// the bucket owning output slot `slot`: largest b in [lo, hi) with off[b] <= slot, given off[lo] <= slot < off[hi]
// (empty buckets share their successor's offset and are never returned).  `hint` is a bucket at or before the answer.
C12_HD uint32_t ba_bucket_of(const uint32_t* off, uint32_t hint, uint32_t hi, uint32_t slot)
{
    uint32_t b = hint;
    for (int step = 0; step < 3; ++step) {          // neighbouring slots of a thread are a bucket or two apart
        if (off[b + 1] > slot) return b;
        ++b;
    }
    uint32_t lo = b;                                 // off[lo] <= slot < off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (off[mid] <= slot)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// Round 1 works on the MERGED level-1 lists: bucket b's list is the concatenation, in group order, of what round 0 left of
// the groups' copies of it (group q's at position region1[q] + off1[q][b] of list buffer 0).  Output slot ii of the bucket adds
// entries 2 ii and 2 ii + 1 of that concatenation; ry = 0xffffffff when there is no second entry.
constexpr uint32_t BA_MAX_GROUPS = 8;
struct BaLevel1 {
    const uint32_t* off1[BA_MAX_GROUPS];
    uint32_t region1[BA_MAX_GROUPS];
    uint32_t groups;
};
C12_HD void ba_ref_level1(const BaLevel1& L, uint32_t b, uint32_t ii, uint32_t& rx, uint32_t& ry)
{
    uint32_t cum = 0, j = 2 * ii;
    rx = ry = 0xffffffffu;
    for (uint32_t q = 0; q < L.groups; ++q) {
        const uint32_t o = L.off1[q][b], len = L.off1[q][b + 1] - o;
        if (rx == 0xffffffffu && j < cum + len) {
            rx = L.region1[q] + o + (j - cum);
            ++j;
        }
        if (rx != 0xffffffffu && ry == 0xffffffffu && j >= cum && j < cum + len) ry = L.region1[q] + o + (j - cum);
        cum += len;
    }
}

// entry j of a bucket's sorted list as a signed affine point
template <class F> C12_HD Affine<F> ba_fetch(const uint32_t* vals, const Affine<F>* points, uint32_t j)
{
    uint32_t v = vals[j];
    Affine<F> pt = points[v & 0x7fffffffu];
    if ((v >> 31) && !affine_is_inf(pt)) pt.y = neg(pt.y);
    return pt;
}

// The addition P + Q in affine form is  lambda = num / den.  kind 0: chord (den = Qx - Px); 1: tangent (den = 2 Py);
// 2: result P (Q is the identity); 3: result Q; 4: result is the identity (Q = -P).  den = 1 where no inverse is needed.
template <class F> C12_HD int ba_denominator(const Affine<F>& P, const Affine<F>& Q, F& den)
{
    den = FieldOps<F>::one();
    if (affine_is_inf(P)) return 3;
    if (affine_is_inf(Q)) return 2;
    F dx = sub(Q.x, P.x);
    if (!is_zero(dx)) {
        den = dx;
        return 0;
    }
    if (!eq(Q.y, P.y)) return 4;
    den = dbl(P.y);     // never zero: the groups have odd order
    return 1;
}
template <class F> C12_HD Affine<F> ba_finish(const Affine<F>& P, const Affine<F>& Q, int kind, const F& inv_den)
{
    if (kind == 2) return P;
    if (kind == 3) return Q;
    if (kind == 4) return affine_inf<F>();
    F num = kind == 0 ? sub(Q.y, P.y) : mul3(sqr(P.x));
    F lam = mul_hot(num, inv_den);
    F x3 = sub(sub(sqr_hot(lam), P.x), Q.x);
    F y3 = sub(mul_hot(lam, sub(P.x, x3)), P.y);
    return Affine<F>{x3, y3};
}

// ---- bucket reduction:  S_w = sum_j (j + 1) B[j]  over the half buckets of window w ----------------------------------
// The buckets are cut into segments of L = seg_len: segment t yields sum0[t] = sum_i (i + 1) B[t L + i] by a running sum and the
// plain total run1[t], so  S_w = sum_t sum0[t] + L sum_t t run1[t].  The second term is a sum with SMALL INTEGER weights, and
// those are taken bit by bit:  sum_t t run1[t] = sum_j 2^j P_j  with  P_j = the sum of the totals whose segment index has bit j
// set - plane_bits independent tree sums (plus one for sum_t sum0[t]) instead of a chain of ever shorter running-sum levels,
// and a Horner recombination of plane_bits + log2 L doublings per window:
//     S_w = T + L (P_0 + 2 (P_1 + 2 (P_2 + ...))),   T = sum_t sum0[t]
// Layout of the reduction scratch (in points): sum0 of window w at [w segs, (w + 1) segs), run1 behind all the sum0's.
C12_HD size_t msm_sum0_offset(const MsmPlan& pl, uint32_t w) { return (size_t)w * pl.segs; }
C12_HD size_t msm_run1_offset(const MsmPlan& pl, uint32_t w) { return (size_t)(pl.windows + w) * pl.segs; }
C12_HD size_t msm_reduce_scratch_points(const MsmPlan& pl) { return 2 * (size_t)pl.windows * pl.segs; }

// one segment [lo, hi): sum = sum_i (i + w0) in[lo + i]  (w0 = 1: the weights start at 1), run = sum_i in[lo + i]
template <class F>
C12_HD void msm_reduce_level_body(const Proj<F>* in, uint32_t lo, uint32_t hi, uint32_t w0, Proj<F>& sum, Proj<F>& run)
{
    run = proj_inf<F>();
    sum = proj_inf<F>();
#pragma unroll 1
    for (uint32_t j = hi; j > lo; --j) {
        run = proj_add(run, in[j - 1]);
        if (w0 || j - 1 > lo) sum = proj_add(sum, run);
    }
}

// plane j of one window: j < plane_bits: the totals whose segment index has bit j set; j = plane_bits: all the sum0's
// (the slice of entries t = first, first + stride, ... : the kernel's threads take one slice each and tree-sum the results)
template <class F>
C12_HD Proj<F> msm_plane_slice_body(const MsmPlan& pl, const Proj<F>* sum0, const Proj<F>* run1, uint32_t j, uint32_t first, uint32_t stride)
{
    Proj<F> acc = proj_inf<F>();
    const bool all = j == pl.plane_bits;
#pragma unroll 1
    for (uint32_t t = first; t < pl.segs; t += stride)
        if (all || ((t >> j) & 1u)) acc = proj_add(acc, all ? sum0[t] : run1[t]);
    return acc;
}

// S_w from the planes of one window: planes[0 .. plane_bits - 1] = P_j, planes[plane_bits] = T
template <class F> C12_HD Proj<F> msm_combine_planes_body(const MsmPlan& pl, const Proj<F>* planes)
{
    if (pl.plane_bits == 0) return planes[0];
    Proj<F> acc = planes[pl.plane_bits - 1];
#pragma unroll 1
    for (uint32_t j = pl.plane_bits - 1; j > 0; --j) acc = proj_add(proj_dbl(acc), planes[j - 1]);
#pragma unroll 1
    for (uint32_t len = pl.seg_len; len > 1; len >>= 1) acc = proj_dbl(acc);
    return proj_add(acc, planes[pl.plane_bits]);
}

// Horner over window sums S_w (w = W-1 .. 0): acc = 2^c acc + S_w
template <class F> C12_HD Proj<F> msm_horner_body(const MsmPlan& pl, const Proj<F>* window_sums)
{
    Proj<F> acc = window_sums[pl.windows - 1];
#pragma unroll 1
    for (uint32_t w = pl.windows - 1; w > 0; --w) {
#pragma unroll 1
        for (uint32_t k = 0; k < pl.c; ++k) acc = proj_dbl(acc);
        acc = proj_add(acc, window_sums[w - 1]);
    }
    return acc;
}

// ---- wire formats ---------------------------------------------------------------------------------------
// G1 affine 96 B = x || y big-endian (canonical); identity = zeros.  Returns false on non-canonical
// coordinates or a point off the curve (ECP_set, ecp_BLS12381.cpp:~280-300).
C12_HD bool g1_from_bytes96(Affine<Fp>& out, const uint8_t* b)
{
    Fp x = fp_from_be48(b), y = fp_from_be48(b + 48);
    if (fp_is_zero(x) && fp_is_zero(y)) {
        out = affine_inf<Fp>();
        return true;
    }
    if (!fp_is_canonical(x) || !fp_is_canonical(y)) {
        out = affine_inf<Fp>();
        return false;
    }
    out.x = fp_to_mont(x);
    out.y = fp_to_mont(y);
    Fp rhs = fp_add(fp_mul(fp_sqr(out.x), out.x), fp_mul4(fp_one()));
    if (!fp_eq(fp_sqr(out.y), rhs)) {
        out = affine_inf<Fp>();
        return false;
    }
    return true;
}

// G2 affine 192 B = x.b || x.a || y.b || y.a (imaginary part first: FP2_toBytes, fp2_BLS12381.cpp:83-87)
C12_HD bool g2_from_bytes192(Affine<Fp2>& out, const uint8_t* b)
{
    Fp xb = fp_from_be48(b), xa = fp_from_be48(b + 48), yb = fp_from_be48(b + 96), ya = fp_from_be48(b + 144);
    if (fp_is_zero(xa) && fp_is_zero(xb) && fp_is_zero(ya) && fp_is_zero(yb)) {
        out = affine_inf<Fp2>();
        return true;
    }
    if (!fp_is_canonical(xa) || !fp_is_canonical(xb) || !fp_is_canonical(ya) || !fp_is_canonical(yb)) {
        out = affine_inf<Fp2>();
        return false;
    }
    out.x = Fp2{fp_to_mont(xa), fp_to_mont(xb)};
    out.y = Fp2{fp_to_mont(ya), fp_to_mont(yb)};
    Fp2 four = Fp2{fp_mul4(fp_one()), fp_zero()};
    Fp2 rhs = add(mul(sqr(out.x), out.x), mul_ip(four));  // x^3 + 4(1+i)
    if (!eq(sqr(out.y), rhs)) {
        out = affine_inf<Fp2>();
        return false;
    }
    return true;
}

template <class F> struct Wire;
template <> struct Wire<Fp> {
    static constexpr int AFFINE = 96, COMPRESSED = 49;
    static C12_HD bool parse(Affine<Fp>& o, const uint8_t* b) { return g1_from_bytes96(o, b); }
    // ECP_toOctet compressed: 0x02 | parity(y), x  (ecp_BLS12381.cpp:445-491); identity = zeros (g1_point.hpp:113-117)
    static C12_HD void compress(uint8_t* o, const Affine<Fp>& p)
    {
        if (affine_is_inf(p)) {
            for (int i = 0; i < 49; ++i) o[i] = 0;
            return;
        }
        o[0] = (uint8_t)(0x02 | fp_sign(p.y));
        fp_to_be48(o + 1, fp_from_mont(p.x));
    }
    // ECP_fromOctet compressed -> ECP_setx (ecp_BLS12381.cpp:495-545,302-323): y = sqrt(x^3 + 4), sign from the tag byte.
    // 49 zero bytes = identity (the C++ layer's convention, g1_point.hpp:87-106).  false: bad tag, x >= p, or no root.
    static C12_HD bool decompress(Affine<Fp>& o, const uint8_t* b)
    {
        o = affine_inf<Fp>();
        uint8_t any = 0;
        for (int i = 0; i < 49; ++i) any |= b[i];
        if (!any) return true;
        if ((b[0] & 0xfe) != 0x02) return false;
        Fp x = fp_from_be48(b + 1);
        if (!fp_is_canonical(x)) return false;
        x = fp_to_mont(x);
        Fp y;
        if (!fp_sqrt(y, fp_add(fp_mul(fp_sqr(x), x), fp_mul4(fp_one())))) return false;
        if (fp_sign(y) != (int)(b[0] & 1)) y = fp_neg(y);
        o = Affine<Fp>{x, y};
        return true;
    }
    static C12_HD void serialize(uint8_t* o, const Affine<Fp>& p)
    {
        if (affine_is_inf(p)) {
            for (int i = 0; i < 96; ++i) o[i] = 0;
            return;
        }
        fp_to_be48(o, fp_from_mont(p.x));
        fp_to_be48(o + 48, fp_from_mont(p.y));
    }
};
template <> struct Wire<Fp2> {
    static constexpr int AFFINE = 192, COMPRESSED = 97;
    static C12_HD bool parse(Affine<Fp2>& o, const uint8_t* b) { return g2_from_bytes192(o, b); }
    // ECP2_toOctet compressed: 0x02 | FP2_sign(y), x.b, x.a (ecp2_BLS12381.cpp:184-222); identity = zeros
    static C12_HD void compress(uint8_t* o, const Affine<Fp2>& p)
    {
        if (affine_is_inf(p)) {
            for (int i = 0; i < 97; ++i) o[i] = 0;
            return;
        }
        o[0] = (uint8_t)(0x02 | fp2_sign(p.y));
        fp_to_be48(o + 1, fp_from_mont(p.x.b));
        fp_to_be48(o + 49, fp_from_mont(p.x.a));
    }
    // ECP2_fromOctet compressed -> ECP2_setx (ecp2_BLS12381.cpp:225-266,322-344): y = sqrt(x^3 + 4(1+i)), FP2_sign from the tag
    static C12_HD bool decompress(Affine<Fp2>& o, const uint8_t* b)
    {
        o = affine_inf<Fp2>();
        uint8_t any = 0;
        for (int i = 0; i < 97; ++i) any |= b[i];
        if (!any) return true;
        if ((b[0] & 0xfe) != 0x02) return false;
        Fp xb = fp_from_be48(b + 1), xa = fp_from_be48(b + 49);
        if (!fp_is_canonical(xa) || !fp_is_canonical(xb)) return false;
        Fp2 x = Fp2{fp_to_mont(xa), fp_to_mont(xb)};
        Fp2 four = Fp2{fp_mul4(fp_one()), fp_zero()};
        Fp2 y;
        if (!fp2_sqrt(y, add(mul(sqr(x), x), mul_ip(four)))) return false;
        if (fp2_sign(y) != (int)(b[0] & 1)) y = neg(y);
        o = Affine<Fp2>{x, y};
        return true;
    }
    static C12_HD void serialize(uint8_t* o, const Affine<Fp2>& p)
    {
        if (affine_is_inf(p)) {
            for (int i = 0; i < 192; ++i) o[i] = 0;
            return;
        }
        fp_to_be48(o, fp_from_mont(p.x.b));
        fp_to_be48(o + 48, fp_from_mont(p.x.a));
        fp_to_be48(o + 96, fp_from_mont(p.y.b));
        fp_to_be48(o + 144, fp_from_mont(p.y.a));
    }
};

} // namespace c12
