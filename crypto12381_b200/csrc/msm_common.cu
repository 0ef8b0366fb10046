// Field-independent parts of the MSM pipeline: scalar recoding, the segmented radix sort driver, bucket bounds,
// and the input-validation flag word.  (Kernels here are launched only through the host functions below, so the
// per-field translation units never reference a __global__ symbol of another TU.)
#include "msm_impl.cuh"
#include "sort.cuh"

namespace c12 {

__global__ void __launch_bounds__(128) k_recode(MsmPlan pl, uint32_t first, uint32_t last, const uint8_t* __restrict__ scalars, uint32_t* __restrict__ keys,
                                                uint32_t* __restrict__ vals, int* flags)
{
    uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;       // terms [first, last): everything, or one upload group
    if (i >= last) return;
    if (i >= pl.n_in) {         // the unused tail of the last group's segments
        msm_recode_pad_body(pl, i, keys, vals);
        return;
    }
    if (!scalar_is_canonical(scalar_from_be32(scalars + 32ull * i))) atomicOr(flags, FLAG_BAD_SCALAR);
    msm_recode_body(pl, i, scalars, keys, vals);
}

// group < 0: all terms; else the terms of upload group `group` (and, for the last group, the padding of its segments)
int launch_recode(const MsmPlan& pl, int group, const uint8_t* d_scalars, uint32_t* keys, uint32_t* vals, int* flags, cudaStream_t s)
{
    uint32_t first = 0, last = pl.groups * pl.n_group;
    if (group >= 0) {
        first = (uint32_t)group * pl.n_group;
        last = first + pl.n_group;
    }
    if (last > first) k_recode<<<cdiv(last - first, 128), 128, 0, s>>>(pl, first, last, d_scalars, keys, vals, flags);
    C12_LAUNCHED();
    return C12381_OK;
}

// ---- bucket lists by counting (the default front end) -----------------------------------------------------------------
// A window's entries only have to be GROUPED by bucket: the order inside a bucket's list is irrelevant to the sum (the group
// law is exact and the result leaves as a canonical encoding), so nothing has to be sorted.  Three steps instead of the
// radix passes and the bounds search:
//   k_recode_count    term -> digits; every entry takes its rank inside its bucket from an atomic counter (counts = the `end`
//                     array, zeroed first) and parks (bucket | sign, rank)
//   k_count_scan_*    exclusive scan of the counters -> start / end of every bucket (and, for the halving rounds, the level-1
//                     slot offsets off1 = scan of ceil(count / 2) in the same pass)
//   k_bucket_scatter  entry -> vals[start[bucket] + rank] = pipeline term | sign
// The counters are 4 B x buckets (1 MiB at c = 16), resident in L2; ranks come back from the L2 atomic unit.  Equal keys in
// neighbouring lanes (equal or tiny scalars: a whole window in one bucket) would serialise on one counter, so a warp that sees any
// aggregates its equal keys with match_any first - one atomic per distinct bucket.
constexpr uint32_t COUNT_NONE = MSM_NO_BUCKET;
__global__ void k_ba_plan_top(uint32_t* __restrict__ tile_sums, uint32_t ntiles);

struct RecodeCount {
    uint32_t half;
    bool live;
    uint32_t* keys;
    uint32_t* ranks;
    uint32_t* counts;
    __device__ __forceinline__ void operator()(uint32_t seg, uint64_t o, uint32_t d, uint32_t val)
    {
        const uint32_t full = 0xffffffffu, lane = threadIdx.x & 31;
        const uint32_t b = (live && d) ? seg * half + (d - 1) : COUNT_NONE;
        const uint32_t nb = __shfl_down_sync(full, b, 1);
        uint32_t rank = 0;
        if (__any_sync(full, lane < 31 && b == nb && b != COUNT_NONE)) {
            const uint32_t peers = __match_any_sync(full, b);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if ((int)lane == leader && b != COUNT_NONE) base = atomicAdd(counts + b, (uint32_t)__popc(peers));
            rank = __shfl_sync(full, base, leader) + __popc(peers & ((1u << lane) - 1u));
        } else if (b != COUNT_NONE) {
            rank = atomicAdd(counts + b, 1u);
        }
        if (live) {
            keys[o] = msm_count_key(d, val);
            ranks[o] = rank;
        }
    }
};

__global__ void __launch_bounds__(128) k_recode_count(MsmPlan pl, uint32_t first, uint32_t last, const uint8_t* __restrict__ scalars,
                                                      uint32_t* __restrict__ keys, uint32_t* __restrict__ ranks, uint32_t* __restrict__ counts, int* flags)
{
    // every lane walks the loops (the emitter's shuffles are warp-wide); terms past `last` write nothing, the unused tail of
    // the last group's segments (i >= n_in) recodes a zero scalar: entries that own no bucket
    const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < last, real = live && i < pl.n_in;
    Scalar256 s;
#pragma unroll
    for (int k = 0; k < 8; ++k) s.v[k] = 0;
    if (real) {
        s = scalar_from_be32(scalars + 32ull * i);
        if (!scalar_is_canonical(s)) {
            atomicOr(flags, FLAG_BAD_SCALAR);
#pragma unroll
            for (int k = 0; k < 8; ++k) s.v[k] = 0;
        }
    }
    RecodeCount emit{pl.half, live, keys, ranks, counts};
    msm_recode_each(pl, live ? i : first, s, emit);
}

constexpr int COUNT_SCAN_ITEMS = 8, COUNT_SCAN_TILE = 256 * COUNT_SCAN_ITEMS;

// tile sums of the counts and of ceil(count / 2)
__global__ void __launch_bounds__(256) k_count_scan_tiles(const uint32_t* __restrict__ counts, uint32_t total, uint32_t* __restrict__ tile_sums, uint32_t ntiles)
{
    const uint32_t base = blockIdx.x * COUNT_SCAN_TILE + threadIdx.x * COUNT_SCAN_ITEMS;
    uint32_t s = 0, h = 0;
#pragma unroll
    for (int i = 0; i < COUNT_SCAN_ITEMS; ++i) {
        const uint32_t m = base + i < total ? counts[base + i] : 0u;
        s += m;
        h += (m + 1) >> 1;
    }
    uint32_t tot;
    block_exclusive_scan_256(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
    block_exclusive_scan_256(h, &tot);
    if (threadIdx.x == 0) tile_sums[ntiles + blockIdx.x] = tot;
}

// counts (in `end`) -> start[b] = pos0 + sum of the counts before b, end[b] = start[b] + count; off1[b] = sum of ceil(count / 2)
// before b (total + 1 entries; only if off1 != nullptr); with refs the odd lists' last slots get their empty second reference
__global__ void __launch_bounds__(256) k_count_scan_apply(uint32_t total, uint32_t pos0, const uint32_t* __restrict__ tile_sums, uint32_t ntiles,
                                                          uint32_t* __restrict__ start, uint32_t* __restrict__ end, uint32_t* __restrict__ off1, uint2* __restrict__ refs)
{
    const uint32_t base = blockIdx.x * COUNT_SCAN_TILE + threadIdx.x * COUNT_SCAN_ITEMS;
    uint32_t v[COUNT_SCAN_ITEMS], s = 0, h = 0;
#pragma unroll
    for (int i = 0; i < COUNT_SCAN_ITEMS; ++i) {
        v[i] = base + i < total ? end[base + i] : 0u;
        s += v[i];
        h += (v[i] + 1) >> 1;
    }
    uint32_t tot;
    uint32_t e = block_exclusive_scan_256(s, &tot) + tile_sums[blockIdx.x] + pos0;
    uint32_t o = block_exclusive_scan_256(h, &tot) + tile_sums[ntiles + blockIdx.x];
#pragma unroll
    for (int i = 0; i < COUNT_SCAN_ITEMS; ++i) {
        if (base + i < total) {
            start[base + i] = e;
            end[base + i] = e + v[i];
        }
        if (off1 && base + i <= total) off1[base + i] = o;          // entry `total` is the level's slot count
        if (refs && (v[i] & 1u)) refs[o + (v[i] >> 1)].y = BA_NONE;  // an odd list's last entry has no partner (the scatter writes only entries)
        e += v[i];
        o += (v[i] + 1) >> 1;
    }
}

// refs == nullptr: the lists themselves (vals).  Otherwise the slot references of halving round 0 directly (what k_ba_map would
// derive from the lists): slot off1[bucket] + rank / 2 adds the entries of ranks 2 i and 2 i + 1 (an odd list's last slot got its
// BA_NONE from the scan) - one gather and one 4-byte store per entry.
__global__ void __launch_bounds__(256) k_bucket_scatter(MsmPlan pl, uint32_t seg0, uint64_t entries, const uint32_t* __restrict__ keys,
                                                        const uint32_t* __restrict__ ranks, const uint32_t* __restrict__ start, uint32_t* __restrict__ vals,
                                                        const uint32_t* __restrict__ off1, uint2* __restrict__ refs)
{
    const uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= entries) return;
    const uint64_t o = (uint64_t)seg0 * pl.n + t;
    const uint32_t key = keys[o];
    if (key == COUNT_NONE) return;
    const uint32_t val = msm_entry_term(pl, o) | (key & 0x80000000u);
    const size_t b = (size_t)(o / pl.n) * pl.half + (key & 0x7fffffffu);
    const uint32_t rank = ranks[o];
    if (!refs) {
        vals[start[b] + rank] = val;
        return;
    }
    reinterpret_cast<uint32_t*>(refs + (off1[b - (size_t)seg0 * pl.half] + (rank >> 1)))[rank & 1] = val;
}

size_t count_scan_scratch_words(uint32_t total) { return 2 * (size_t)cdiv((size_t)total + 1, COUNT_SCAN_TILE) + 8; }

// Group `group` (< 0: everything) of the list plan: from its scalars to the bucket lists.  counts/end: zeroed here.  The bucket
// bounds of the group's segments land in start / end (global positions in `vals`, the group's region starting at its first
// segment), off1 (optional) gets the group's level-1 slot offsets.
int launch_bucket_lists(const MsmPlan& pl, int group, const uint8_t* d_scalars, uint32_t* keys, uint32_t* ranks, uint32_t* vals, uint32_t* start,
                        uint32_t* end, uint32_t* off1, uint2* refs, uint32_t* tile_sums, int* flags, cudaStream_t s, cudaEvent_t after_recode, cudaEvent_t after_scan)
{
    const uint32_t g = group < 0 ? 0u : (uint32_t)group, ngroups = group < 0 ? pl.groups : 1u;
    const uint32_t first = g * pl.n_group, last = first + ngroups * pl.n_group;
    const uint32_t seg0 = g * pl.real_windows, nseg = ngroups * pl.real_windows, total = nseg * pl.half;
    uint32_t *gstart = start + (size_t)seg0 * pl.half, *gend = end + (size_t)seg0 * pl.half;
    if (last <= first) return C12381_OK;
    C12_CUDA(cudaMemsetAsync(gend, 0, 4 * (size_t)total, s));
    k_recode_count<<<cdiv(last - first, 128), 128, 0, s>>>(pl, first, last, d_scalars, keys, ranks, gend - (size_t)seg0 * pl.half, flags);
    C12_LAUNCHED();
    if (after_recode) C12_CUDA(cudaEventRecord(after_recode, s));
    const uint32_t ntiles = cdiv((size_t)total + 1, COUNT_SCAN_TILE);
    k_count_scan_tiles<<<ntiles, 256, 0, s>>>(gend, total, tile_sums, ntiles);
    C12_LAUNCHED();
    k_ba_plan_top<<<2, 256, 0, s>>>(tile_sums, ntiles);
    C12_LAUNCHED();
    k_count_scan_apply<<<ntiles, 256, 0, s>>>(total, (uint32_t)((uint64_t)seg0 * pl.n), tile_sums, ntiles, gstart, gend, off1, off1 ? refs : nullptr);
    C12_LAUNCHED();
    if (after_scan) C12_CUDA(cudaEventRecord(after_scan, s));        // bounds and level-1 offsets are in place: all the merged plan needs
    const uint64_t entries = (uint64_t)nseg * pl.n;
    k_bucket_scatter<<<(unsigned)cdiv(entries, (uint64_t)256), 256, 0, s>>>(pl, seg0, entries, keys, ranks, start, vals, off1, off1 ? refs : nullptr);
    C12_LAUNCHED();
    return C12381_OK;
}

// segments [seg0, seg0 + nseg) of the sorted key array -> bounds of their buckets (global positions, global bucket ids)
int launch_bucket_bounds(const MsmPlan& pl, uint32_t seg0, uint32_t nseg, const uint32_t* keys, uint32_t* start, uint32_t* end, cudaStream_t s)
{
    C12_CUDA(cudaMemsetAsync(start + (size_t)seg0 * pl.half, 0, 4 * (size_t)nseg * pl.half, s));
    C12_CUDA(cudaMemsetAsync(end + (size_t)seg0 * pl.half, 0, 4 * (size_t)nseg * pl.half, s));
    k_bucket_bounds<<<dim3(cdiv(pl.n, 256), nseg), 256, 0, s>>>(keys, pl.n, pl.half, start, end, seg0);
    C12_LAUNCHED();
    return C12381_OK;
}

// ---- offsets of the bucket lists after each batch-affine halving round ------------------------------------------------
// Level k = the lists after k halvings.  Level 1 is per upload group (round 0 adds up each group's sorted entries pairwise);
// from there on a bucket's list is the CONCATENATION of its groups' level-1 lists (msm_impl.cuh), so with
//   M1_b = sum over groups of ceil(m_(g,b) / 2)      the level-k list of bucket b holds  ceil(M1_b / 2^(k-1))  entries.
// off_k[b] = sum over b' < b of that, total + 1 entries per level; three launches for all requested levels together
// (blockIdx.y = level - first level): tile sums, scan of the tile sums, apply.
constexpr int BA_PLAN_ITEMS = 8, BA_PLAN_TILE = 256 * BA_PLAN_ITEMS;

struct BaPlanGeom {
    const uint32_t* start;      // bucket bounds: group g's copy of bucket b is entry g * gstride + b
    const uint32_t* end;
    uint32_t total, groups, gstride, first_level;
};

__device__ __forceinline__ uint32_t ba_plan_len(const BaPlanGeom& g, uint32_t b, uint32_t level)
{
    if (b >= g.total) return 0u;
    uint32_t m1 = 0;
    for (uint32_t q = 0; q < g.groups; ++q) m1 += ba_len(g.end[(size_t)q * g.gstride + b] - g.start[(size_t)q * g.gstride + b], 1);
    return ba_len(m1, level - 1);
}

__global__ void __launch_bounds__(256) k_ba_plan_tiles(BaPlanGeom g, uint32_t* __restrict__ tile_sums, uint32_t ntiles)
{
    const uint32_t level = g.first_level + blockIdx.y, base = blockIdx.x * BA_PLAN_TILE + threadIdx.x * BA_PLAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < BA_PLAN_ITEMS; ++i) s += ba_plan_len(g, base + i, level);
    uint32_t tot;
    block_exclusive_scan_256(s, &tot);
    if (threadIdx.x == 0) tile_sums[(size_t)blockIdx.y * ntiles + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(256) k_ba_plan_top(uint32_t* __restrict__ tile_sums, uint32_t ntiles)
{
    uint32_t* t = tile_sums + (size_t)blockIdx.x * ntiles;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < ntiles; base += 256) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v = idx < ntiles ? t[idx] : 0;
        uint32_t tot;
        const uint32_t e = block_exclusive_scan_256(v, &tot);
        if (idx < ntiles) t[idx] = e + carry;
        carry += tot;
    }
}

__global__ void __launch_bounds__(256) k_ba_plan_apply(BaPlanGeom g, const uint32_t* __restrict__ tile_sums, uint32_t ntiles, uint32_t* __restrict__ off)
{
    const uint32_t level = g.first_level + blockIdx.y, base = blockIdx.x * BA_PLAN_TILE + threadIdx.x * BA_PLAN_ITEMS;
    uint32_t v[BA_PLAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < BA_PLAN_ITEMS; ++i) {
        v[i] = ba_plan_len(g, base + i, level);
        s += v[i];
    }
    uint32_t tot;
    uint32_t e = block_exclusive_scan_256(s, &tot) + tile_sums[(size_t)blockIdx.y * ntiles + blockIdx.x];
    uint32_t* o = off + (size_t)blockIdx.y * (g.total + 1);
#pragma unroll
    for (int i = 0; i < BA_PLAN_ITEMS; ++i) {
        if (base + i <= g.total) o[base + i] = e;       // entry `total` is the level's slot count
        e += v[i];
    }
}

size_t ba_plan_scratch_words(uint32_t total, uint32_t levels) { return (size_t)levels * (cdiv((size_t)total + 1, BA_PLAN_TILE) + 1) + 64; }

// off[(k - first_level) * (total + 1) + b] for k = first_level .. first_level + levels - 1, numbered from 0.
// One group's level 1: groups = 1 with that group's bounds; the merged levels >= 2: all groups' bounds, gstride apart.
int launch_ba_plan(uint32_t total, const uint32_t* start, const uint32_t* end, uint32_t groups, uint32_t gstride, uint32_t first_level, uint32_t levels,
                   uint32_t* off, uint32_t* tile_sums, cudaStream_t s)
{
    if (levels == 0) return C12381_OK;
    BaPlanGeom g;
    g.start = start;
    g.end = end;
    g.total = total;
    g.groups = groups;
    g.gstride = gstride;
    g.first_level = first_level;
    const uint32_t ntiles = cdiv((size_t)total + 1, BA_PLAN_TILE);
    k_ba_plan_tiles<<<dim3(ntiles, levels), 256, 0, s>>>(g, tile_sums, ntiles);
    C12_LAUNCHED();
    k_ba_plan_top<<<levels, 256, 0, s>>>(tile_sums, ntiles);
    C12_LAUNCHED();
    k_ba_plan_apply<<<dim3(ntiles, levels), 256, 0, s>>>(g, tile_sums, ntiles, off);
    C12_LAUNCHED();
    return C12381_OK;
}

// where the accumulation finds what the halving rounds left of every bucket list: lane l (a pipeline of a group) wrote its last
// round's output from final_region[l] on, in its own slot numbering
__global__ void __launch_bounds__(256) k_ba_list_bounds(BaListGeom g, uint32_t* __restrict__ lstart, uint32_t* __restrict__ lend)
{
    const uint32_t vb = blockIdx.x * 256 + threadIdx.x;
    if (vb >= g.vtotal) return;
    uint32_t l = 0;
    while (l + 1 < g.lanes && vb >= g.vb0[l + 1]) ++l;
    const uint32_t b = g.b_lo[l] + (vb - g.vb0[l]);
    const uint32_t org = g.off[l][g.b_lo[l]];
    lstart[vb] = g.final_region[l] + (g.off[l][b] - org);
    lend[vb] = g.final_region[l] + (g.off[l][b + 1] - org);
}

int launch_ba_list_bounds(const BaListGeom& g, uint32_t* lstart, uint32_t* lend, cudaStream_t s)
{
    k_ba_list_bounds<<<cdiv(g.vtotal, 256), 256, 0, s>>>(g, lstart, lend);
    C12_LAUNCHED();
    return C12381_OK;
}

// every output slot of the rounds [g.first_round, g.first_round + g.rounds) -> the two inputs it adds (blockIdx.y = round -
// first round; see msm_impl.cuh).  A thread resolves BA_MAP_RUN consecutive slots: one binary search for the first, then the
// bucket only moves forward by a step or two.
//   round 0 (a group's own): inputs are (term | sign) values of the group's sorted entries
//   round 1: inputs are positions in the groups' level-1 lists, a bucket's list being their concatenation in group order
//   rounds >= 2: positions in the previous round's output, pipeline by pipeline
constexpr uint32_t BA_MAP_RUN = 8;
__global__ void __launch_bounds__(256) k_ba_map(BaMapGeom g)
{
    const uint32_t r = g.first_round + blockIdx.y;
    const uint32_t* off_out = g.off_out[blockIdx.y];
    const uint32_t n_slots = off_out[g.total];
    uint32_t t = (blockIdx.x * 256 + threadIdx.x) * BA_MAP_RUN;
    if (t >= n_slots) return;
    const uint32_t t_end = t + BA_MAP_RUN < n_slots ? t + BA_MAP_RUN : n_slots;
    const uint32_t* off_in = g.off_in[blockIdx.y];       // rounds >= 2
    uint32_t b = ba_bucket_of(off_out, 0u, g.total, t), p = 0;
    while (p + 1 < g.pipes && b >= g.b_lo[p + 1]) ++p;
    uint32_t lo = 0, hi = 0, p_slot0 = off_out[g.b_lo[p]];
    bool fresh = true;
    for (; t < t_end; ++t) {
        if (fresh || t >= hi) {
            if (!fresh) b = ba_bucket_of(off_out, b + 1, g.total, t);
            fresh = false;
            lo = off_out[b];
            hi = off_out[b + 1];
            if (p + 1 < g.pipes && b >= g.b_lo[p + 1]) {
                while (p + 1 < g.pipes && b >= g.b_lo[p + 1]) ++p;
                p_slot0 = off_out[g.b_lo[p]];
            }
        }
        const uint32_t ii = t - lo;
        uint2 ref;
        if (r == 0) {
            const uint32_t m = g.end[b] - g.start[b], pos = g.start[b] + 2 * ii;
            ref.x = g.vals[pos];
            ref.y = 2 * ii + 1 < m ? g.vals[pos + 1] : BA_NONE;
        } else if (r == 1) {
            ba_ref_level1(g.level1, b, ii, ref.x, ref.y);
        } else {
            const uint32_t len = off_in[b + 1] - off_in[b];
            const uint32_t pos = g.list_region[(r - 1) & 1][p] + (off_in[b] + 2 * ii - off_in[g.b_lo[p]]);
            ref.x = pos;
            ref.y = 2 * ii + 1 < len ? pos + 1 : BA_NONE;
        }
        g.refs[(size_t)g.ref_region[blockIdx.y][p] + (t - p_slot0)] = ref;
    }
}

int launch_ba_map(const BaMapGeom& g, uint32_t max_slots, cudaStream_t s)
{
    if (g.rounds == 0) return C12381_OK;
    k_ba_map<<<dim3(cdiv(max_slots, 256 * BA_MAP_RUN), g.rounds), 256, 0, s>>>(g);
    C12_LAUNCHED();
    return C12381_OK;
}

// ---- chunking of the bucket lists (virtual buckets) ------------------------------------------------------------------
__global__ void k_chunk_counts(uint32_t total, uint32_t chunk, const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                               uint32_t* __restrict__ out)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < total) out[b] = msm_chunks(end[b] - start[b], chunk);
    if (b == total) out[b] = 0;
}
__global__ void k_chunk_pad(uint32_t vmax, uint32_t* __restrict__ keys, uint32_t* __restrict__ ids)
{
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= vmax) return;
    keys[v] = 255u;             // sorts behind every real chunk
    ids[v] = 0xffffffffu;
}
__global__ void k_chunk_map(uint32_t total, uint32_t chunk, const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                            const uint32_t* __restrict__ vstart, uint32_t* __restrict__ vbucket, uint32_t* __restrict__ keys,
                            uint32_t* __restrict__ ids)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total) return;
    const uint32_t m = end[b] - start[b], n = msm_chunks(m, chunk), v0 = vstart[b];
    for (uint32_t j = 0; j < n; ++j) {
        uint32_t sz = m - j * chunk;
        if (sz > chunk) sz = chunk;
        vbucket[v0 + j] = b;
        keys[v0 + j] = 255u - (sz > 255u ? 255u : sz);   // ascending key = descending length
        ids[v0 + j] = v0 + j;
    }
}

// layout of the scratch: vstart[total + 1] | vbucket[vmax] | keys, ids, keys2, ids2 [vmax each] | hist | scan tiles
size_t chunk_order_scratch_words(const MsmPlan& pl)
{
    size_t tile_words = 0;
    size_t hist = sort_scratch_words(pl.vmax, 1, &tile_words);
    size_t scan_tiles = cdiv((size_t)pl.total + 1, SCAN_TILE) + 2;
    return (size_t)pl.total + 1 + 5 * (size_t)pl.vmax + hist + tile_words + scan_tiles + 64;
}

int launch_chunk_order(const MsmPlan& pl, const uint32_t* start, const uint32_t* end, uint32_t* scratch, uint32_t** vstart_out,
                       uint32_t** vbucket_out, uint32_t** order, cudaStream_t s)
{
    size_t tile_words = 0;
    size_t hist_words = sort_scratch_words(pl.vmax, 1, &tile_words);
    uint32_t* vstart = scratch;
    uint32_t* vbucket = vstart + pl.total + 1;
    uint32_t* keys = vbucket + pl.vmax;
    uint32_t* ids = keys + pl.vmax;
    uint32_t* keys2 = ids + pl.vmax;
    uint32_t* ids2 = keys2 + pl.vmax;
    uint32_t* hist = ids2 + pl.vmax;
    uint32_t* tiles = hist + hist_words;
    uint32_t* scan_tiles = tiles + tile_words;
    const uint32_t m = pl.total + 1;
    k_chunk_counts<<<cdiv(m, 256), 256, 0, s>>>(pl.total, pl.chunk, start, end, vstart);
    C12_LAUNCHED();
    const uint32_t ntiles = cdiv(m, SCAN_TILE);
    k_scan_tile_sums<<<ntiles, SCAN_THREADS, 0, s>>>(vstart, m, scan_tiles);
    C12_LAUNCHED();
    k_scan_top<<<1, SCAN_THREADS, 0, s>>>(scan_tiles, ntiles);
    C12_LAUNCHED();
    k_scan_apply<<<ntiles, SCAN_THREADS, 0, s>>>(vstart, m, scan_tiles);
    C12_LAUNCHED();
    k_chunk_pad<<<cdiv(pl.vmax, 256), 256, 0, s>>>(pl.vmax, keys, ids);
    C12_LAUNCHED();
    k_chunk_map<<<cdiv(pl.total, 256), 256, 0, s>>>(pl.total, pl.chunk, start, end, vstart, vbucket, keys, ids);
    C12_LAUNCHED();
    int rc = sort_pairs_segmented(keys, ids, keys2, ids2, pl.vmax, 1, 8, hist, tiles, s);
    if (rc) return rc;
    *vstart_out = vstart;
    *vbucket_out = vbucket;
    *order = ids;
    return C12381_OK;
}

size_t sort_scratch_words(uint32_t n, uint32_t nseg, size_t* tile_words)
{
    size_t nblk = cdiv(n, SORT_TILE);
    size_t m = (size_t)nseg * 256 * nblk;
    if (tile_words) *tile_words = cdiv(m, SCAN_TILE) + 1;
    return m;
}

// Stable LSD sort of nseg independent segments of n (key, value) pairs each, keys < 2^key_bits.
// On return `keys`/`vals` point at the sorted arrays (the pointers are swapped per pass).
int sort_pairs_segmented(uint32_t*& keys, uint32_t*& vals, uint32_t*& keys_alt, uint32_t*& vals_alt, uint32_t n, uint32_t nseg,
                         uint32_t key_bits, uint32_t* hist, uint32_t* tile_sums, cudaStream_t s)
{
    const uint32_t nblk = cdiv(n, SORT_TILE);
    const size_t m = (size_t)nseg * 256 * nblk;
    const uint32_t ntiles = cdiv(m, SCAN_TILE);
    for (uint32_t shift = 0; shift < key_bits; shift += 8) {
        k_radix_hist<<<dim3(nblk, nseg), SORT_THREADS, 0, s>>>(keys, n, (int)shift, hist, nblk);
        C12_LAUNCHED();
        k_scan_tile_sums<<<ntiles, SCAN_THREADS, 0, s>>>(hist, m, tile_sums);
        C12_LAUNCHED();
        k_scan_top<<<1, SCAN_THREADS, 0, s>>>(tile_sums, ntiles);
        C12_LAUNCHED();
        k_scan_apply<<<ntiles, SCAN_THREADS, 0, s>>>(hist, m, tile_sums);
        C12_LAUNCHED();
        k_radix_scatter<<<dim3(nblk, nseg), SORT_THREADS, 0, s>>>(keys, vals, keys_alt, vals_alt, n, (int)shift, hist, nblk);
        C12_LAUNCHED();
        uint32_t* t = keys; keys = keys_alt; keys_alt = t;
        t = vals; vals = vals_alt; vals_alt = t;
    }
    return C12381_OK;
}

// host entries: the host flag word is cleared on the way in and read on the way out (HostScope in msm_impl.cuh keeps
// Ctx::host_depth > 0 in between, which is what routes the kernels' reports to that word)
int flags_reset(cudaStream_t s)
{
    C12_CUDA(cudaMemsetAsync(ctx().d_flags + HOST_FLAG_WORD, 0, sizeof(int), s));
    return C12381_OK;
}

static int flags_collect_word(int word, cudaStream_t s)
{
    Ctx& c = ctx();
    C12_CUDA(cudaMemcpyAsync(c.h_flags + word, c.d_flags + word, sizeof(int), cudaMemcpyDeviceToHost, s));
    C12_CUDA(cudaStreamSynchronize(s));
    C12_CUDA(cudaGetLastError());
    int f = c.h_flags[word];
    if (f) {
        cudaMemsetAsync(c.d_flags + word, 0, sizeof(int), s);
        const char* what = (f & FLAG_BAD_POINT) ? "malformed input: non-canonical coordinate or point off the curve"
                           : (f & FLAG_BAD_SCALAR) ? "malformed input: scalar >= group order"
                                                   : "malformed input: reference POD outside its documented range";
        return set_error(C12381_EINPUT, what);
    }
    return C12381_OK;
}

int flags_collect(cudaStream_t s) { return flags_collect_word(HOST_FLAG_WORD, s); }

} // namespace c12

using namespace c12;

extern "C" int c12381_sync_status(void* stream)
{
    C12_REQUIRE_CTX();
    return flags_collect_word(0, (cudaStream_t)stream);
}
