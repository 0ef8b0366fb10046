// TEST-ONLY host mirror: compiles the product's host+device templates (crypto12381_b200/csrc/*.cuh) with a
// plain C++ compiler and runs the per-thread kernel BODIES serially on the CPU, so the algorithms, indexing
// and wire formats can be checked against the oracle in the GPU-less build container.  The product never
// loads this library; the device build uses the inline-PTX field primitives instead of the portable ones.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../crypto12381_b200/csrc/msm_core.cuh"
#include "../../crypto12381_b200/csrc/scalar_mul.cuh"
#include "../../crypto12381_b200/csrc/pairing.cuh"
#include "../../crypto12381_b200/csrc/miracl_pod.cuh"
#include "../../crypto12381_b200/csrc/hash.cuh"

using namespace c12;

namespace {
MsmPlan make_plan(uint32_t n, uint32_t c, uint32_t seg_len, uint32_t parts)
{
    MsmPlan pl = msm_make_plan(n, c, parts);
    if (seg_len) msm_plan_levels(pl, seg_len > pl.half ? pl.half : seg_len);   // a power of two
    return pl;
}

// parts = 1: no split; MsmTraits<F>::PARTS: the split the device entries use
template <class F> int msm(const uint8_t* pts, const uint8_t* sc, uint32_t n_in, uint32_t c, uint32_t seg_len, uint8_t* out, uint32_t parts = 1, uint32_t rounds = 0, uint32_t chunk_override = 0, uint32_t groups = 1, uint32_t counting = 0)
{
    using W = Wire<F>;
    if (n_in == 0) {
        W::compress(out, affine_inf<F>());
        return 0;
    }
    const MsmPlan rpl = make_plan(n_in, c, seg_len, parts);     // the reduction's plan
    MsmPlan pl = msm_list_plan(rpl, groups);                    // the list side: (group, window) segments with buckets of their own
    const uint32_t n = pl.n;
    std::vector<Affine<F>> P(rpl.n);
    for (uint32_t i = 0; i < n_in; ++i) {
        if (!W::parse(P[i], pts + (size_t)W::AFFINE * i)) return -1;
        for (uint32_t q = 1; q < parts; ++q) P[(size_t)q * n_in + i] = MsmTraits<F>::endo(q, P[i]);
    }
    size_t N = (size_t)n * pl.windows;
    std::vector<uint32_t> keys(N, 0xdeadbeefu), vals(N, 0xdeadbeefu);
    for (uint32_t i = 0; i < n_in; ++i) msm_recode_body(pl, i, sc, keys.data(), vals.data());
    for (uint32_t i = n_in; i < pl.groups * pl.n_group; ++i) msm_recode_pad_body(pl, i, keys.data(), vals.data());
    for (size_t i = 0; i < N; ++i)
        if (keys[i] == 0xdeadbeefu) return -4;          // every entry of every segment is written
    if (groups > 1)                                     // a group's lists name that group's terms only
        for (size_t i = 0; i < N; ++i)
            if (keys[i] < pl.half && ((vals[i] & 0x7fffffffu) % n_in) / pl.n_group != (i / n) / pl.real_windows) return -5;
    // segmented stable sort: window w owns [w*n, (w+1)*n), keys are window-local (as the device sort does)
    std::vector<uint32_t> order(N);
    for (size_t i = 0; i < N; ++i) order[i] = (uint32_t)i;
    for (uint32_t w = 0; w < pl.windows; ++w)
        std::stable_sort(order.begin() + (size_t)w * n, order.begin() + (size_t)(w + 1) * n,
                         [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint32_t> sk(N), sv(N);
    for (size_t i = 0; i < N; ++i) { sk[i] = keys[order[i]]; sv[i] = vals[order[i]]; }
    std::vector<uint32_t> start(pl.total, 0), end(pl.total, 0);
    if (counting) {
        // The device's default front end (k_recode_count -> scan -> k_bucket_scatter): every entry takes a rank inside its bucket
        // from a counter, the lists are filled by (bucket start + rank).  The atomic arrival order is arbitrary on the device;
        // here the entries arrive in a scrambled order (counting = the stride of the walk), which must not change the sum.
        struct Emit {
            std::vector<uint32_t>& k;
            void operator()(uint32_t, uint64_t o, uint32_t d, uint32_t val) { k[o] = msm_count_key(d, val); }
        };
        std::vector<uint32_t> ck(N, 0xdeadbeefu), rank(N, 0), counts(pl.total, 0);
        Emit emit{ck};
        Scalar256 zero;
        for (int k = 0; k < 8; ++k) zero.v[k] = 0;
        for (uint32_t i = 0; i < pl.groups * pl.n_group; ++i) msm_recode_each(pl, i, i < n_in ? scalar_from_be32(sc + 32ull * i) : zero, emit);
        size_t stride = counting;
        auto gcd = [](size_t a, size_t b) { while (b) { size_t t = a % b; a = b; b = t; } return a; };
        while (gcd(stride, N) != 1) ++stride;
        for (size_t j = 0, o = 0; j < N; ++j, o = (o + stride) % N) {
            if (ck[o] == 0xdeadbeefu) return -6;
            if (ck[o] == MSM_NO_BUCKET) continue;
            rank[o] = counts[(o / n) * pl.half + (ck[o] & 0x7fffffffu)]++;
        }
        for (uint32_t seg = 0, pos = 0; seg < pl.windows; ++seg) {
            pos = (uint32_t)((size_t)seg * n);                    // a segment's lists start at its first entry, as on the device per group
            if (seg % pl.real_windows) pos = end[(size_t)seg * pl.half - 1];
            for (uint32_t k = 0; k < pl.half; ++k) {
                const size_t b = (size_t)seg * pl.half + k;
                start[b] = pos;
                pos += counts[b];
                end[b] = pos;
            }
        }
        std::fill(sv.begin(), sv.end(), 0xdeadbeefu);
        for (size_t o = 0; o < N; ++o) {
            if (ck[o] == MSM_NO_BUCKET) continue;
            const size_t b = (o / n) * pl.half + (ck[o] & 0x7fffffffu);
            const uint32_t val = msm_entry_term(pl, o) | (ck[o] & 0x80000000u);
            // the same (term | sign) the sorted front end attaches to this entry
            if (val != vals[o]) return -7;
            sv[start[b] + rank[o]] = val;
        }
    } else
    for (uint32_t w = 0; w < pl.windows; ++w)
        for (size_t i = 0; i < n; ++i) {
            size_t g = (size_t)w * n + i;
            uint32_t k = sk[g];
            if (k >= pl.half) continue;
            size_t b = (size_t)w * pl.half + k;
            if (i == 0 || sk[g - 1] != k) start[b] = (uint32_t)g;
            if (i + 1 == n || sk[g + 1] != k) end[b] = (uint32_t)g + 1;
        }
    std::vector<Proj<F>> buckets(pl.total);
    if (rounds == 0) {
        // k_accumulate over chunks of the bucket lists + k_fold
        if (chunk_override) pl.chunk = chunk_override;
        for (uint32_t b = 0; b < pl.total; ++b) {
            uint32_t m = end[b] - start[b], nch = msm_chunks(m, pl.chunk);
            std::vector<Proj<F>> parts(nch);
            for (uint32_t j = 0; j < nch; ++j) {
                uint32_t lo = start[b] + j * pl.chunk, hi = lo + pl.chunk > end[b] ? end[b] : lo + pl.chunk;
                parts[j] = msm_accumulate_range_body<F>(lo, hi, sv.data(), P.data());
            }
            buckets[b] = msm_fold_body<F>(parts.data(), nch);
        }
    } else {
        // the batch-affine halving rounds (k_ba_fwd / k_ba_inv / k_ba_bwd over the flat slot space), one inversion per pair here
        // instead of one per 32 J additions: round 0 per upload group on its own copies of the buckets, the rounds behind it on
        // the merged lists (a bucket's level-1 list = the concatenation of its groups', ba_ref_level1), then the accumulation over
        // what the rounds leave (k_accumulate<F, true>) on the REAL buckets
        const uint32_t B = rpl.total, G = groups;
        auto add_pair = [&](const Affine<F>& X, const Affine<F>& Y) {
            F den;
            int kind = ba_denominator(X, Y, den);
            return ba_finish(X, Y, kind, inv(den));
        };
        // round 0
        std::vector<std::vector<uint32_t>> off1(G, std::vector<uint32_t>(B + 1, 0));
        std::vector<uint32_t> region1(G, 0);
        for (uint32_t g = 0; g < G; ++g) {
            for (uint32_t b = 0; b < B; ++b) off1[g][b + 1] = off1[g][b] + ba_len(end[(size_t)g * B + b] - start[(size_t)g * B + b], 1);
            if (g + 1 < G) region1[g + 1] = region1[g] + off1[g][B] + 3;     // regions apart, as the capacity-based ones on the device
        }
        std::vector<Affine<F>> lvl1(region1[G - 1] + off1[G - 1][B] + 3, affine_inf<F>());
        for (uint32_t g = 0; g < G; ++g) {
            uint32_t hint = 0;
            for (uint32_t slot = 0; slot < off1[g][B]; ++slot) {
                const uint32_t b = ba_bucket_of(off1[g].data(), (slot & 7u) ? hint : 0u, B, slot);   // with and without a hint
                hint = b;
                if (!(off1[g][b] <= slot && slot < off1[g][b + 1])) return -3;
                const uint32_t vb = g * B + b, ii = slot - off1[g][b], m = end[vb] - start[vb], in0 = start[vb] + 2 * ii;
                Affine<F> X = ba_fetch<F>(sv.data(), P.data(), in0);
                lvl1[region1[g] + slot] = 2 * ii + 1 < m ? add_pair(X, ba_fetch<F>(sv.data(), P.data(), in0 + 1)) : X;
            }
        }
        // merged rounds 1 .. rounds - 1
        BaLevel1 L1;
        L1.groups = G;
        for (uint32_t g = 0; g < G; ++g) {
            L1.off1[g] = off1[g].data();
            L1.region1[g] = region1[g];
        }
        std::vector<uint32_t> m1(B, 0);
        for (uint32_t b = 0; b < B; ++b)
            for (uint32_t g = 0; g < G; ++g) m1[b] += off1[g][b + 1] - off1[g][b];
        std::vector<Affine<F>> cur, next;
        std::vector<uint32_t> off_in, off_out(B + 1, 0);
        for (uint32_t r = 1; r < rounds; ++r) {
            off_in = off_out;
            for (uint32_t b = 0; b < B; ++b) off_out[b + 1] = off_out[b] + ba_len(m1[b], r);
            next.assign(off_out[B], affine_inf<F>());
            for (uint32_t b = 0; b < B; ++b)
                for (uint32_t ii = 0; ii < off_out[b + 1] - off_out[b]; ++ii) {
                    if (r == 1) {
                        uint32_t rx, ry;
                        ba_ref_level1(L1, b, ii, rx, ry);
                        if (rx == 0xffffffffu) return -6;
                        next[off_out[b] + ii] = ry != 0xffffffffu ? add_pair(lvl1[rx], lvl1[ry]) : lvl1[rx];
                    } else {
                        const uint32_t len = off_in[b + 1] - off_in[b], in0 = off_in[b] + 2 * ii;
                        next[off_out[b] + ii] = 2 * ii + 1 < len ? add_pair(cur[in0], cur[in0 + 1]) : cur[in0];
                    }
                }
            cur.swap(next);
        }
        buckets.assign(B, proj_inf<F>());
        for (uint32_t b = 0; b < B; ++b) {
            XYZZ<F> acc = xyzz_inf<F>();
            if (rounds == 1) {          // one round: the level-1 lists are what is left (the device runs this with a single group only)
                for (uint32_t g = 0; g < G; ++g)
                    for (uint32_t j = off1[g][b]; j < off1[g][b + 1]; ++j)
                        if (!affine_is_inf(lvl1[region1[g] + j])) xyzz_madd(acc, lvl1[region1[g] + j]);
            } else {
                for (uint32_t j = off_out[b]; j < off_out[b + 1]; ++j)
                    if (!affine_is_inf(cur[j])) xyzz_madd(acc, cur[j]);
            }
            buckets[b] = xyzz_to_proj(acc);
        }
        groups = 1;     // merged
    }
    if (groups > 1) {       // k_fold: the groups' copies of a bucket are added up
        std::vector<Proj<F>> merged(rpl.total);
        for (uint32_t b = 0; b < rpl.total; ++b) {
            Proj<F> acc = buckets[b];
            for (uint32_t g = 1; g < groups; ++g) acc = proj_add(acc, buckets[(size_t)g * rpl.total + b]);
            merged[b] = acc;
        }
        buckets.swap(merged);
    }
    pl = rpl;
    std::vector<Proj<F>> wsum(pl.windows);
    for (uint32_t w = 0; w < pl.windows; ++w) {
        // the reduction of msm_core.cuh: segment running sums (k_reduce_level0), one tree sum per bit plane of the segment
        // index (k_reduce_planes, here with 3 threads' slices per plane), recombination (k_finish)
        const Proj<F>* B = buckets.data() + (size_t)w * pl.half;
        std::vector<Proj<F>> sum0(pl.segs), run1(pl.segs), planes(pl.plane_bits + 1);
        for (uint32_t t = 0; t < pl.segs; ++t) {
            uint32_t lo = t * pl.seg_len, hi = lo + pl.seg_len > pl.half ? pl.half : lo + pl.seg_len;
            msm_reduce_level_body<F>(B, lo, hi, 1u, sum0[t], run1[t]);
        }
        for (uint32_t j = 0; j <= pl.plane_bits; ++j) {
            Proj<F> acc = proj_inf<F>();
            for (uint32_t first = 0; first < 3; ++first) acc = proj_add(acc, msm_plane_slice_body<F>(pl, sum0.data(), run1.data(), j, first, 3));
            planes[j] = acc;
        }
        wsum[w] = msm_combine_planes_body<F>(pl, planes.data());
    }
    Proj<F> r = msm_horner_body<F>(pl, wsum.data());
    W::compress(out, proj_to_affine(r));
    return 0;
}
} // namespace

extern "C" {
// plain-integer (48 B BE) field ops through the Montgomery code path: op 0 mul, 1 add, 2 sub, 3 inv, 4 neg, 5 sqrt
void hm_fp_op(int op, const uint8_t* a48, const uint8_t* b48, uint8_t* out48)
{
    Fp a = fp_to_mont(fp_from_be48(a48));
    Fp b = fp_to_mont(fp_from_be48(b48));
    Fp r;
    switch (op) {
    case 0: r = fp_mul(a, b); break;
    case 1: r = fp_add(a, b); break;
    case 2: r = fp_sub(a, b); break;
    case 3: r = fp_inv(a); break;
    case 4: r = fp_neg(a); break;
    default: r = fp_sqrt_candidate(a); break;
    }
    fp_to_be48(out48, fp_from_mont(r));
}

int hm_g1_msm(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint32_t seg, uint8_t* out49) { return msm<Fp>(p, s, n, c, seg, out49); }
int hm_g2_msm(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint32_t seg, uint8_t* out97) { return msm<Fp2>(p, s, n, c, seg, out97); }
// the device entries' scalar split with bucket lists cut into chunks of `chunk` entries (k_accumulate + k_fold)
int hm_g1_msm_chunked(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint32_t chunk, uint8_t* out49)
{
    return msm<Fp>(p, s, n, c, 0, out49, MsmTraits<Fp>::PARTS, 0, chunk);
}
// with `rounds` batch-affine pre-reduction rounds and the device entries' scalar split
int hm_g1_msm_ba(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint32_t rounds, uint8_t* out49)
{
    return msm<Fp>(p, s, n, c, 0, out49, MsmTraits<Fp>::PARTS, rounds);
}
int hm_g2_msm_ba(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint32_t rounds, uint8_t* out97)
{
    return msm<Fp2>(p, s, n, c, 0, out97, MsmTraits<Fp2>::PARTS, rounds);
}
// the same with the terms cut into `groups` upload groups (virtual windows per group, merged after the accumulation)
int hm_g1_msm_groups(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint32_t rounds, uint32_t groups, uint8_t* out49)
{
    return msm<Fp>(p, s, n, c, 0, out49, MsmTraits<Fp>::PARTS, rounds, 0, groups);
}
int hm_g2_msm_groups(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint32_t rounds, uint32_t groups, uint8_t* out97)
{
    return msm<Fp2>(p, s, n, c, 0, out97, MsmTraits<Fp2>::PARTS, rounds, 0, groups);
}
// the same with the bucket lists made by counting (the device's default front end), entries arriving in a walk of stride `arrival`
int hm_g1_msm_counting(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint32_t rounds, uint32_t groups, uint32_t arrival, uint8_t* out49)
{
    return msm<Fp>(p, s, n, c, 0, out49, MsmTraits<Fp>::PARTS, rounds, 0, groups, arrival ? arrival : 1);
}
int hm_g2_msm_counting(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint32_t rounds, uint32_t groups, uint32_t arrival, uint8_t* out97)
{
    return msm<Fp2>(p, s, n, c, 0, out97, MsmTraits<Fp2>::PARTS, rounds, 0, groups, arrival ? arrival : 1);
}

int hm_g1_mul(const uint8_t* p96, const uint8_t* s32, uint32_t n, uint8_t* out49)
{
    for (uint32_t i = 0; i < n; ++i)
        if (!scalar_mul_body<Fp>(p96 + 96 * (size_t)i, s32 + 32 * (size_t)i, out49 + 49 * (size_t)i)) return -1;
    return 0;
}
int hm_g2_mul(const uint8_t* p192, const uint8_t* s32, uint32_t n, uint8_t* out97)
{
    for (uint32_t i = 0; i < n; ++i)
        if (!scalar_mul_body<Fp2>(p192 + 192 * (size_t)i, s32 + 32 * (size_t)i, out97 + 97 * (size_t)i)) return -1;
    return 0;
}
void hm_g1_fixed_base(const uint8_t* s32, uint32_t n, uint8_t* out96)
{
    for (uint32_t i = 0; i < n; ++i) fixed_base_body<Fp>(s32 + 32 * (size_t)i, out96 + 96 * (size_t)i);
}
void hm_g2_fixed_base(const uint8_t* s32, uint32_t n, uint8_t* out192)
{
    for (uint32_t i = 0; i < n; ++i) fixed_base_body<Fp2>(s32 + 32 * (size_t)i, out192 + 192 * (size_t)i);
}

// pairings: B instances x k pairs; mode 0 raw Miller value, 1 final-exponentiated GT
int hm_pairing_product(const uint8_t* g1, const uint8_t* g2, uint32_t B, uint32_t k, int mode, uint8_t* out576)
{
    for (uint32_t b = 0; b < B; ++b)
        if (!pairing_product_body(g1 + 96 * (size_t)b * k, g2 + 192 * (size_t)b * k, k, mode, out576 + 576 * (size_t)b)) return -1;
    return 0;
}
void hm_final_exp(const uint8_t* in576, uint32_t B, uint8_t* out576)
{
    for (uint32_t b = 0; b < B; ++b) final_exp_body(in576 + 576 * (size_t)b, out576 + 576 * (size_t)b);
}
void hm_gt_mul(const uint8_t* a, const uint8_t* b, uint32_t B, uint8_t* out576)
{
    for (uint32_t i = 0; i < B; ++i) gt_mul_body(a + 576 * (size_t)i, b + 576 * (size_t)i, out576 + 576 * (size_t)i);
}
void hm_gt_pow(const uint8_t* a, const uint8_t* s32, uint32_t B, uint8_t* out576)
{
    for (uint32_t i = 0; i < B; ++i) gt_pow_body(a + 576 * (size_t)i, s32 + 32 * (size_t)i, out576 + 576 * (size_t)i);
}

void hm_gt_pow_gs(const uint8_t* a, const uint8_t* s32, uint32_t B, uint8_t* out576)
{
    for (uint32_t i = 0; i < B; ++i) gt_pow_gs_body(a + 576 * (size_t)i, s32 + 32 * (size_t)i, out576 + 576 * (size_t)i);
}

// reference PODs <-> wire formats (bodies of the k_pod_* kernels)
int hm_pod_points_to_wire(int g2, const uint8_t* pods, uint32_t n, uint8_t* wire)
{
    bool ok = true;
    for (uint32_t i = 0; i < n; ++i)
        ok = (g2 ? pod_point_to_wire<Fp2>(pods + (size_t)POD_P2 * i, wire + 192 * (size_t)i)
                 : pod_point_to_wire<Fp>(pods + (size_t)POD_P1 * i, wire + 96 * (size_t)i)) && ok;
    return ok ? 0 : -1;
}
int hm_wire_to_pod_points(int g2, const uint8_t* wire, uint32_t n, uint8_t* pods)
{
    bool ok = true;
    for (uint32_t i = 0; i < n; ++i)
        ok = (g2 ? wire_to_pod_point<Fp2>(wire + 192 * (size_t)i, pods + (size_t)POD_P2 * i)
                 : wire_to_pod_point<Fp>(wire + 96 * (size_t)i, pods + (size_t)POD_P1 * i)) && ok;
    return ok ? 0 : -1;
}
int hm_pod_bigs_to_scalars(const uint8_t* bigs, uint32_t n, uint8_t* out32)
{
    bool ok = true;
    for (uint32_t i = 0; i < n; ++i) ok = pod_big_to_scalar(bigs + (size_t)POD_BIG * i, out32 + 32 * (size_t)i) && ok;
    return ok ? 0 : -1;
}
int hm_pod_fp12_to_wire(const uint8_t* pods, uint32_t n, uint8_t* wire)
{
    bool ok = true;
    for (uint32_t e = 0; e < n; ++e)
        for (uint32_t j = 0; j < 12; ++j) ok = pod_fp12_coeff_to_wire(pods + (size_t)POD_FP12 * e, j, wire + 576 * (size_t)e) && ok;
    return ok ? 0 : -1;
}
void hm_wire_to_pod_fp12(const uint8_t* wire, uint32_t n, uint8_t* pods)
{
    for (uint32_t e = 0; e < n; ++e)
        for (uint32_t j = 0; j < 12; ++j) wire_to_pod_fp12_coeff(wire + 576 * (size_t)e, j, pods + (size_t)POD_FP12 * e);
}
uint32_t hm_choose_window(uint64_t n) { return msm_choose_window(n); }
uint32_t hm_choose_window_glv(uint64_t n) { return msm_choose_window(2 * n, 128); }
// k (32 B BE) -> |k0|, |k1| (16 B BE each) and signs, k = +-k0 +- k1 * x^2 (mod r)
void hm_glv_split(const uint8_t* k32, uint8_t* a0, uint8_t* a1, uint32_t* signs)
{
    GlvHalves h = glv_split(scalar_from_be32(k32));
    for (int i = 0; i < 4; ++i)
        for (int b = 0; b < 4; ++b) {
            a0[15 - (4 * i + b)] = (uint8_t)(h.a0[i] >> (8 * b));
            a1[15 - (4 * i + b)] = (uint8_t)(h.a1[i] >> (8 * b));
        }
    signs[0] = h.neg0;
    signs[1] = h.neg1;
}
// auto plan (what the device entry uses): window from msm_choose_window, segments from msm_make_plan
int hm_g1_msm_auto(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint8_t* out49)
{
    // exactly the device entry's plan: GLV split on, window from msm_choose_window(2n, 128)
    return msm<Fp>(p, s, n, c ? c : msm_choose_window(2 * (uint64_t)n, 128), 0, out49, MsmTraits<Fp>::PARTS);
}
// the G2 device entry's plan: GLS split in four, window from msm_choose_window(4n, 64)
int hm_g2_msm_auto(const uint8_t* p, const uint8_t* s, uint32_t n, uint32_t c, uint8_t* out97)
{
    return msm<Fp2>(p, s, n, c ? c : msm_choose_window(4 * (uint64_t)n, 64), 0, out97, MsmTraits<Fp2>::PARTS);
}
// k (32 B BE) -> |k_i| (8 B BE each) and signs, k = sum_i +-k_i z^i (mod r)
void hm_gls_split(const uint8_t* k32, uint8_t* mags32, uint32_t* signs)
{
    ScalarParts sp;
    gls_split(scalar_from_be32(k32), sp);
    for (int q = 0; q < 4; ++q) {
        uint64_t m = (uint64_t)sp.mag[q][0] | ((uint64_t)sp.mag[q][1] << 32);
        for (int b = 0; b < 8; ++b) mags32[8 * q + 7 - b] = (uint8_t)(m >> (8 * b));
        signs[q] = sp.neg[q];
    }
}
// compressed -> affine (Wire<F>::decompress, the body of k_decompress); returns 0 ok, -1 on a malformed encoding
int hm_g1_decompress(const uint8_t* in49, uint32_t n, uint8_t* out96)
{
    int rc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        Affine<Fp> p;
        if (!Wire<Fp>::decompress(p, in49 + 49 * (size_t)i)) rc = -1;
        Wire<Fp>::serialize(out96 + 96 * (size_t)i, p);
    }
    return rc;
}
int hm_g2_decompress(const uint8_t* in97, uint32_t n, uint8_t* out192)
{
    int rc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        Affine<Fp2> p;
        if (!Wire<Fp2>::decompress(p, in97 + 97 * (size_t)i)) rc = -1;
        Wire<Fp2>::serialize(out192 + 192 * (size_t)i, p);
    }
    return rc;
}
// subgroup membership bodies (k_subgroup_check); returns -1 if a point does not parse
int hm_g1_member(const uint8_t* p96, uint32_t n, uint8_t* out)
{
    int rc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        Affine<Fp> p;
        if (!Wire<Fp>::parse(p, p96 + 96 * (size_t)i)) rc = -1;
        out[i] = subgroup_member(p) ? 1 : 0;
    }
    return rc;
}
int hm_g2_member(const uint8_t* p192, uint32_t n, uint8_t* out)
{
    int rc = 0;
    for (uint32_t i = 0; i < n; ++i) {
        Affine<Fp2> p;
        if (!Wire<Fp2>::parse(p, p192 + 192 * (size_t)i)) rc = -1;
        out[i] = subgroup_member(p) ? 1 : 0;
    }
    return rc;
}
// SHA3-512 / hash-to-Zp bodies (k_sha3_512)
void hm_sha3_512(const uint8_t* msg, uint32_t len, uint8_t* out64) { sha3_512(msg, len, out64); }
void hm_hash_to_zp(const uint8_t* msg, uint32_t len, uint8_t* out32) { hash_to_zp_body(msg, len, out32); }
// hash-to-G1 bodies (k_hash_to_g1)
void hm_hash_to_g1(const uint8_t* msg, uint32_t len, uint8_t* out49) { hash_to_g1_body(msg, len, out49); }
int hm_map_to_g1(const uint8_t* u48, uint8_t* out49) { return map_to_g1_body(u48, out49) ? 0 : 1; }
}

#if defined(C12_COUNT_FP_MUL)
// tools/count_fp_mul.py: Montgomery products executed by the kernel bodies (host build with -DC12_COUNT_FP_MUL)
extern "C" unsigned long long hm_fp_addsub_count(int reset)
{
    unsigned long long n = c12::host::addsub_counter();
    if (reset) c12::host::addsub_counter() = 0;
    return n;
}
extern "C" unsigned long long hm_fp_mul_count(int reset)
{
    unsigned long long n = c12::host::mont_mul_counter();
    if (reset) c12::host::mont_mul_counter() = 0;
    return n;
}
#endif
