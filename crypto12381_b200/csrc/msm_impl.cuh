// __global__ wrappers and the host-side pipeline of the MSM / batched scalar multiplication entries, templated
// on the coordinate field (Fp -> G1, Fp2 -> G2).  Included by msm_g1.cu and msm_g2.cu, which instantiate one
// field each so the two compile in parallel.  The per-thread bodies are in msm_core.cuh / scalar_mul.cuh.
//
// Pipeline of one MSM (all on the caller's stream, no host synchronisation inside):
//   k_recode          scalars -> pieces -> signed window digits -> (window-local key, term|sign) pairs, window-major
//   radix sort        segmented per window, ceil(c/8) passes of hist / scan / scatter  (sort.cuh)
//   k_bucket_bounds   [start, end) of every bucket in the sorted pairs
//   chunk order       long lists cut into chunks, chunk ids sorted by length (msm_common.cu)
//   k_parse_points    wire bytes -> Montgomery affine points + endomorphism images (canonical + on-curve checks)
//   k_accumulate      one thread per chunk: XYZZ mixed additions of its terms         <- the dominant kernel
//   k_fold            partial sums of a bucket's chunks -> the bucket
//   k_reduce_level0   running sums over segments of the buckets:  sum_i (i + 1) B_i  and the plain total of each segment
//   k_reduce_planes   one tree sum per bit plane of the segment index over the totals (+ one over the segment sums)
//   k_finish          plane recombination, Horner over windows (lane-cooperative doublings), normalisation, encoding
// At large n the accumulation is preceded by batch-affine halving rounds of the bucket lists (k_ba_fwd / k_ba_inv / k_ba_bwd,
// c12381_set_msm_batch_affine): k_accumulate then only walks what they leave, normally one sum per bucket.
#pragma once
#include "common.cuh"
#include "scalar_mul.cuh"

namespace c12 {

enum { FLAG_BAD_POINT = 1, FLAG_BAD_SCALAR = 2 };
enum { OUT_COMPRESSED = 0, OUT_AFFINE = 1 };

template <class T> __device__ __forceinline__ T shfl_down_obj(const T& x, int delta)
{
    static_assert(sizeof(T) % 4 == 0, "word-sized objects only");
    T r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&x);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); ++i) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta);
    return r;
}

template <class T> __device__ __forceinline__ T shfl_bcast_obj(const T& x, int src)
{
    static_assert(sizeof(T) % 4 == 0, "word-sized objects only");
    T r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&x);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); ++i) d[i] = __shfl_sync(0xffffffffu, s[i], src);
    return r;
}

// Sum of one value per thread over a 256-thread block (fixed tree: deterministic).  Result valid in thread 0.
template <class F> __device__ Proj<F> block_sum_256(Proj<F> acc)
{
    __shared__ Proj<F> warp_part[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll 1
    for (int off = 16; off >= 1; off >>= 1) {
        Proj<F> o = shfl_down_obj(acc, off);
        acc = proj_add(acc, o);
    }
    if (lane == 0) warp_part[w] = acc;
    __syncthreads();
    if (w == 0) {
        acc = lane < 8 ? warp_part[lane] : proj_inf<F>();
#pragma unroll 1
        for (int off = 4; off >= 1; off >>= 1) {
            Proj<F> o = shfl_down_obj(acc, off);
            acc = proj_add(acc, o);
        }
    }
    return acc;
}

template <class F>
__global__ void __launch_bounds__(128) k_parse_points(const uint8_t* __restrict__ bytes, uint32_t first, uint32_t last, uint32_t n, uint32_t parts,
                                                      Affine<F>* __restrict__ pts, int* flags)
{
    uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;      // terms [first, last) of n: one upload group
    if (i >= last) return;
    Affine<F> p;
    if (!Wire<F>::parse(p, bytes + (size_t)Wire<F>::AFFINE * i)) atomicOr(flags, FLAG_BAD_POINT);
    pts[i] = p;
    for (uint32_t q = 1; q < parts; ++q) pts[(size_t)q * n + i] = MsmTraits<F>::endo(q, p);   // the identity (0, 0) maps to itself
}

// defined once in msm_common.cu (kernels there are launched through these host functions)
int launch_recode(const MsmPlan& pl, int group, const uint8_t* d_scalars, uint32_t* keys, uint32_t* vals, int* flags, cudaStream_t s);
int launch_bucket_lists(const MsmPlan& pl, int group, const uint8_t* d_scalars, uint32_t* keys, uint32_t* ranks, uint32_t* vals, uint32_t* start,
                        uint32_t* end, uint32_t* off1, uint2* refs, uint32_t* tile_sums, int* flags, cudaStream_t s, cudaEvent_t after_recode, cudaEvent_t after_scan);
size_t count_scan_scratch_words(uint32_t total);
int launch_bucket_bounds(const MsmPlan& pl, uint32_t seg0, uint32_t nseg, const uint32_t* keys, uint32_t* start, uint32_t* end, cudaStream_t s);
// chunking of the bucket lists: vstart[b] = first chunk of bucket b (total + 1 entries), vbucket[v] = bucket of chunk v,
// order[] = chunk ids by decreasing length, padded with 0xffffffff up to pl.vmax
int launch_chunk_order(const MsmPlan& pl, const uint32_t* start, const uint32_t* end, uint32_t* scratch, uint32_t** vstart, uint32_t** vbucket,
                       uint32_t** order, cudaStream_t s);
size_t chunk_order_scratch_words(const MsmPlan& pl);

// offsets of the bucket lists after k halvings, for a run of levels k (msm_common.cu): total + 1 entries per level, the last
// one being the level's slot count
size_t ba_plan_scratch_words(uint32_t total, uint32_t levels);
int launch_ba_plan(uint32_t total, const uint32_t* start, const uint32_t* end, uint32_t groups, uint32_t gstride, uint32_t first_level, uint32_t levels,
                   uint32_t* off, uint32_t* tile_sums, cudaStream_t s);

// ---- batch-affine halving rounds (msm_core.cuh "batch-affine halving rounds") -----------------------------------------
// k_ba_map (msm_common.cu, ONE launch for all rounds, scalar-only: it runs while the points are still being uploaded)
// resolves every output slot of every round to the two inputs it adds: bucket search in the round's offsets, then either the
// (term | sign) values of the sorted array (round 0) or positions in the previous round's list buffer.
// One round is then three launches over the FLAT space of the round's output slots [off_out[b_lo], off_out[b_hi]):
//   k_ba_fwd : warp w owns 32 J consecutive slots, lane l the slots  base + 32 i + l  (coalesced loads and stores).  A lane
//              multiplies up the denominators of its J additions (prefixes parked in HBM, 48 B per slot); the 32 lane totals
//              go through a 5-step shuffle butterfly that leaves in every lane the warp total AND the product of the OTHER
//              lanes' totals; the warp total goes to the round's inversion pool.
//   k_ba_inv : one thread per pool entry inverts it - all 32 lanes of a warp invert (SIMT makes one inversion cost a whole
//              warp's issue slots, so they are pooled across the grid: one inversion per 32 J additions).
//   k_ba_bwd : 1 / (lane total) = pool inverse x product of the others; the lane unwinds its prefixes (2 products per addition)
//              and finishes each addition (lambda, lambda^2, y3).
// Equal / opposite / identity operands are classified per pair (ba_denominator) and take no part in the inversion.
// Pipelines run rounds apart from each other, and the slot numbering shrinks from round to round, so every pipeline keeps its
// lists, prefixes and slot references in regions of its own, indexed from the start of ITS slot range (k_ba_list_bounds turns
// the last round's regions into one pair of bound arrays for the accumulation).
constexpr uint32_t BA_MAX_ROUNDS = 16, BA_MAX_PIPES = BA_MAX_GROUPS, BA_NONE = 0xffffffffu;

struct BaMapGeom {
    // round 0 (a group's own): its bucket bounds in the sorted array and the sorted (term | sign) values
    const uint32_t* start;
    const uint32_t* end;
    const uint32_t* vals;
    // the rounds resolved by this launch: first_round .. first_round + rounds - 1, arrays indexed by round - first_round
    const uint32_t* off_out[BA_MAX_ROUNDS];             // offsets of the round's OUTPUT lists (total + 1 entries): the slot numbering
    const uint32_t* off_in[BA_MAX_ROUNDS];              // offsets of its INPUT lists (rounds >= 2)
    uint32_t ref_region[BA_MAX_ROUNDS][BA_MAX_PIPES];   // where the slots of (round, pipeline) start in refs[]
    BaLevel1 level1;            // round 1: a bucket's input list is the concatenation of the groups' level-1 lists (msm_core.cuh)
    uint2* refs;                // out: the two inputs of every slot (second = BA_NONE: the odd last entry of a list, carried over)
    uint32_t total, first_round, rounds, pipes;
    uint32_t b_lo[BA_MAX_PIPES + 1];                    // pipeline p owns buckets [b_lo[p], b_lo[p + 1])
    uint32_t list_region[2][BA_MAX_PIPES];              // pipeline p's region in list buffer 0 / 1 (rounds >= 2)
};
int launch_ba_map(const BaMapGeom& g, uint32_t max_slots, cudaStream_t s);

struct BaGeom {
    const uint32_t* off_out;    // offsets of the round's OUTPUT lists (total + 1 entries): the slot numbering
    const uint2* refs;          // this (round, pipeline)'s slot references
    uint32_t b_lo, b_hi, J;
    uint32_t out_region, scratch_region;
};

// k_ba_list_bounds (msm_common.cu): the reduced lists of all lanes (a lane = one pipeline of one upload group) as ONE pair of
// bound arrays over the virtual buckets, for the chunking and the accumulation
struct BaListGeom {
    const uint32_t* off[BA_MAX_PIPES];      // the lane's last-round offsets (its group's array)
    uint32_t b_lo[BA_MAX_PIPES];            // first bucket of the lane in that array
    uint32_t vb0[BA_MAX_PIPES + 1];         // first virtual bucket of the lane
    uint32_t final_region[BA_MAX_PIPES];    // where the lane's last round wrote
    uint32_t lanes, vtotal;
};
int launch_ba_list_bounds(const BaListGeom& g, uint32_t* lstart, uint32_t* lend, cudaStream_t s);

template <class F> struct BaIo {
    const Affine<F>* pts;       // round 0: the parsed points (references are term | sign)
    const Affine<F>* lists;     // rounds >= 1: the previous round's output (references are positions)
    Affine<F>* out;
    F* prefix;                  // per slot: running product of the lane's denominators up to and including this slot
    F* pool;                    // per warp: total, replaced by its inverse (k_ba_inv)
    F* others;                  // per lane: product of the other 31 lane totals
};

template <class F> __device__ __forceinline__ F shfl_xor_obj(const F& x, int mask)
{
    static_assert(sizeof(F) % 4 == 0, "word-sized objects only");
    F r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&x);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(F) / 4); ++i) d[i] = __shfl_xor_sync(0xffffffffu, s[i], mask);
    return r;
}

#ifndef C12_BA_THREADS
#define C12_BA_THREADS 128
#endif
#ifndef C12_BA_MIN_BLOCKS
#define C12_BA_MIN_BLOCKS 4       // 128 registers (16 bytes of spill) and 16 warps per SM against 138 and 12: accumulation 4.85 -> 4.71 ms (profiles/r02q)
#endif
#ifndef C12_BA2_MIN_BLOCKS
#define C12_BA2_MIN_BLOCKS 2      // over Fp2 the unwinding pass holds twice the state: 168 registers spill, 255 do not
#endif
template <class F> struct BaShape {
    static constexpr int MIN_BLOCKS = sizeof(F) == sizeof(Fp) ? C12_BA_MIN_BLOCKS : C12_BA2_MIN_BLOCKS;
    // automatic rounds from this many (term, window) entries on: 2^22 over Fp, 2^21 over Fp2 (an Fp2 addition saves more per round:
    // G2 n = 2^17 5.06 -> 4.83 ms, profiles/r03q_rounds_sweep.txt)
    static constexpr uint32_t AUTO_MIN_LOG = sizeof(F) == sizeof(Fp) ? 22 : 21;
};

// the input a slot reference names, as a signed affine point
template <class F, bool FIRST> __device__ __forceinline__ Affine<F> ba_input(const BaIo<F>& io, uint32_t ref)
{
    if (!FIRST) return io.lists[ref];
    Affine<F> pt = io.pts[ref & 0x7fffffffu];
    if ((ref >> 31) && !affine_is_inf(pt)) pt.y = neg(pt.y);
    return pt;
}
template <class F, bool FIRST> __device__ __forceinline__ const Affine<F>* ba_input_ptr(const BaIo<F>& io, uint32_t ref)
{
    return FIRST ? io.pts + (ref & 0x7fffffffu) : io.lists + ref;
}
// The next slot's operands used to be pulled into L2 ahead of their use.  With 16 resident warps per SM (4 blocks, r02q) the
// warps hide that latency themselves and the prefetches only add traffic (the ncu capture of round 0 showed 0.5 GB of reads
// beyond what the 64-byte fetch granularity explains): without them G1 n = 2^20 -0.4 %, n = 2^22 -0.6 %, G2 n = 2^20 -1.5 %
// (profiles/r03y_prefetch_ab.txt).  -DC12_BA_PREFETCH brings them back, -DC12_BA_PREFETCH_L1 aims them at L1.
#if !defined(C12_BA_PREFETCH) && !defined(C12_BA_PREFETCH_L1)
__device__ __forceinline__ void prefetch_l2(const void*) {}
#elif defined(C12_BA_PREFETCH_L1)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#else
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

// -DC12_BA_LOCKSTEP: the warps of a block walk the slot loops together (one barrier per slot, as k_accumulate does): the
// unwinding pass is ~35 KB of straight-line products.  Every thread then stays in the loop for all J trips.
#if defined(C12_BA_LOCKSTEP)
#define C12_BA_ENTER(g, L)                  \
    if (!ba_lane(g, L)) {                   \
        L.count = 0;                        \
        L.wl = 0;                           \
    }
#define C12_BA_TRIPS(g, L) (g).J
#define C12_BA_STEP(i, L)     \
    __syncthreads();          \
    if ((i) >= (L).count) continue
#define C12_BA_LEAVE(L) \
    if ((L).n_slots <= (uint64_t)(L).gw * 32u * g.J) return
#else
#define C12_BA_ENTER(g, L) \
    if (!ba_lane(g, L)) return
#define C12_BA_TRIPS(g, L) (L).count
#define C12_BA_STEP(i, L) ((void)0)
#define C12_BA_LEAVE(L) ((void)0)
#endif
// slots of this lane: local indices  wl + 32 i + lane,  i < count  (wl = the warp's first local slot)
struct BaLane {
    uint32_t wl, count, gw, lane, n_slots, slot_lo;
};
__device__ __forceinline__ bool ba_lane(const BaGeom& g, BaLane& L)
{
    L.slot_lo = g.off_out[g.b_lo];
    L.n_slots = g.off_out[g.b_hi] - L.slot_lo;
    L.gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    L.lane = threadIdx.x & 31;
    const uint64_t wl = (uint64_t)L.gw * 32u * g.J;
    if (wl >= L.n_slots) return false;
    L.wl = (uint32_t)wl;
    const uint32_t left = L.n_slots - L.wl;
    L.count = left > L.lane ? (left - L.lane + 31u) / 32u : 0u;
    if (L.count > g.J) L.count = g.J;
    return true;
}

template <class F, bool FIRST>
__global__ void __launch_bounds__(C12_BA_THREADS) k_ba_fwd(BaGeom g, BaIo<F> io)
{
    BaLane L;
    C12_BA_ENTER(g, L);
    const uint2* refs = g.refs + L.wl + L.lane;
    F* prefix = io.prefix + g.scratch_region + L.wl + L.lane;
    F c = FieldOps<F>::one();
    // software pipeline: the references run two slots ahead, the x coordinates one slot ahead of the products
    uint2 r1 = L.count > 0 ? refs[0] : make_uint2(0u, BA_NONE), r2 = L.count > 1 ? refs[32] : make_uint2(0u, BA_NONE);
    F px = FieldOps<F>::zero(), qx = px;
    if (L.count > 0 && r1.y != BA_NONE) {
        px = ba_input_ptr<F, FIRST>(io, r1.x)->x;
        qx = ba_input_ptr<F, FIRST>(io, r1.y)->x;
    }
#pragma unroll 1
    for (uint32_t i = 0; i < C12_BA_TRIPS(g, L); ++i) {
        C12_BA_STEP(i, L);
        const uint2 r0 = r1;
        const F px0 = px, qx0 = qx;
        r1 = r2;
        if (i + 2 < L.count) r2 = refs[(size_t)(i + 2) * 32u];
        if (i + 1 < L.count && r1.y != BA_NONE) {
            px = ba_input_ptr<F, FIRST>(io, r1.x)->x;
            qx = ba_input_ptr<F, FIRST>(io, r1.y)->x;
        }
        if (r0.y != BA_NONE) {
            // the common case needs the two x coordinates only; anything else (identity operands, equal x) is classified in full
            F den = sub(qx0, px0);
            bool use = !is_zero(den) && !is_zero(px0) && !is_zero(qx0);
            if (!use) use = ba_denominator(ba_input<F, FIRST>(io, r0.x), ba_input<F, FIRST>(io, r0.y), den) < 2;
            if (use) c = mul_hot(c, den);
        }
        prefix[(size_t)i * 32u] = c;
    }
    C12_BA_LEAVE(L);
    // butterfly over the 32 lane totals: w = product of my group, o = product of everyone outside my lane
    F w = c, o = FieldOps<F>::one();
#pragma unroll 1
    for (int k = 1; k < 32; k <<= 1) {
        const F peer = shfl_xor_obj(w, k);
        o = mul_hot(o, peer);
        w = mul_hot(w, peer);
    }
    io.others[(size_t)L.gw * 32u + L.lane] = o;
    if (L.lane == 0) io.pool[L.gw] = w;
}

template <class F> __global__ void __launch_bounds__(128) k_ba_inv(BaGeom g, F* pool)
{
    const uint32_t n_slots = g.off_out[g.b_hi] - g.off_out[g.b_lo];
    const uint32_t t = blockIdx.x * 128 + threadIdx.x;
    if ((uint64_t)t * 32u * g.J >= n_slots) return;
    pool[t] = inv(pool[t]);
}

template <class F, bool FIRST>
__global__ void __launch_bounds__(C12_BA_THREADS, BaShape<F>::MIN_BLOCKS) k_ba_bwd(BaGeom g, BaIo<F> io)
{
    BaLane L;
    C12_BA_ENTER(g, L);
    const uint2* refs = g.refs + L.wl + L.lane;
    const F* prefix = io.prefix + g.scratch_region + L.wl + L.lane;
    Affine<F>* out = io.out + g.out_region + L.wl + L.lane;
    F iv = FieldOps<F>::one();
    if (L.count) iv = mul_hot(io.pool[L.gw], io.others[(size_t)L.gw * 32u + L.lane]);      // 1 / (this lane's total)
    // the references run two slots ahead; the points of the next slot are pulled into L2 while this one is finished
    uint2 r1 = L.count > 0 ? refs[(size_t)(L.count - 1) * 32u] : make_uint2(0u, BA_NONE);
    uint2 r2 = L.count > 1 ? refs[(size_t)(L.count - 2) * 32u] : make_uint2(0u, BA_NONE);
#pragma unroll 1
    for (uint32_t i = C12_BA_TRIPS(g, L); i-- > 0;) {
        C12_BA_STEP(i, L);
        const uint2 r0 = r1;
        r1 = r2;
        if (i >= 2) r2 = refs[(size_t)(i - 2) * 32u];
        if (i >= 1) {
            const char* a = reinterpret_cast<const char*>(ba_input_ptr<F, FIRST>(io, r1.x));
            prefetch_l2(a);
            prefetch_l2(a + sizeof(Affine<F>) - 16);
            if (FIRST && r1.y != BA_NONE) {      // later rounds: the partner is the next list entry, on the same or the following line
                const char* b = reinterpret_cast<const char*>(ba_input_ptr<F, FIRST>(io, r1.y));
                prefetch_l2(b);
                prefetch_l2(b + sizeof(Affine<F>) - 16);
            } else if (!FIRST) {
                prefetch_l2(a + 2 * sizeof(Affine<F>) - 16);
            }
        }
        const Affine<F> P = ba_input<F, FIRST>(io, r0.x);
        if (r0.y == BA_NONE) {               // odd last entry of its list: carried over
            out[(size_t)i * 32u] = P;
            continue;
        }
        const Affine<F> Q = ba_input<F, FIRST>(io, r0.y);
        F den;
        const int kind = ba_denominator(P, Q, den);
        F inv_den = iv;
        if (kind < 2) {
            if (i) inv_den = mul_hot(iv, prefix[(size_t)(i - 1) * 32u]);
            iv = mul_hot(iv, den);
        }
        out[(size_t)i * 32u] = ba_finish(P, Q, kind, inv_den);
    }
}

// Accumulation over CHUNKS of the bucket lists (virtual buckets): thread t takes chunk order[t] - chunks are visited in
// order of decreasing length (launch_chunk_order), so the 32 lanes of a warp run loops of equal length and the heaviest
// start first - and leaves a partial sum; k_fold adds up the partials of each bucket.
// Block shape (measured, profiles/r01aj_acc_lockstep_ab.txt): 128 threads x 3 resident blocks over Fp, 64 threads x 6 over Fp2.
// The warps of a block walk the loop TOGETHER (one barrier per addition, -DC12_ACC_NO_LOCKSTEP removes it): the straight-line
// mixed addition is ~60 KB of code, and warps that drift apart each stream it through the instruction cache on their own (ncu:
// no_instruction stalls 1.0 per issued instruction before, profiles/r01z_k_accumulate_full.md).  Same additions, same order.
#if !defined(C12_ACC_NO_LOCKSTEP) && !defined(C12_ACC_LOCKSTEP)
#define C12_ACC_LOCKSTEP 1
#endif
template <class F> struct AccShape {
#if defined(C12_ACC_THREADS)
#if !defined(C12_ACC_MIN_BLOCKS)
#define C12_ACC_MIN_BLOCKS 1
#endif
    static constexpr int THREADS = C12_ACC_THREADS, MIN_BLOCKS = C12_ACC_MIN_BLOCKS;
#else
#if !defined(C12_ACC2_MIN_BLOCKS)
#define C12_ACC2_MIN_BLOCKS 6
#endif
    static constexpr int THREADS = sizeof(F) == sizeof(Fp) ? 128 : 64, MIN_BLOCKS = sizeof(F) == sizeof(Fp) ? 3 : C12_ACC2_MIN_BLOCKS;
#endif
};
// DIRECT: the lists are arrays of points already (the output of the batch-affine halving rounds: start[] / end[] are that
// round's offsets), not (term | sign) values to be looked up in pts.
template <class F, bool DIRECT>
__global__ void __launch_bounds__(AccShape<F>::THREADS, AccShape<F>::MIN_BLOCKS) k_accumulate(uint32_t vmax, uint32_t chunk, const uint32_t* __restrict__ start,
                                                    const uint32_t* __restrict__ end, const uint32_t* __restrict__ vals,
                                                    const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ order,
                                                    const uint32_t* __restrict__ vbucket, const uint32_t* __restrict__ vstart,
                                                    Proj<F>* __restrict__ vpartial)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
#if defined(C12_ACC_LOCKSTEP)
    // every thread of the block stays in the loop up to the block's longest chunk (chunks are length-ordered: nearly equal)
    const uint32_t v = t < vmax ? order[t] : 0xffffffffu;
    uint32_t lo = 0, hi = 0;
    if (v != 0xffffffffu) {
        const uint32_t b = vbucket[v];
        lo = start[b] + (v - vstart[b]) * chunk;
        hi = lo + chunk;
        if (hi > end[b]) hi = end[b];
    }
    __shared__ uint32_t s_len;
    if (threadIdx.x == 0) s_len = 0;
    __syncthreads();
    atomicMax(&s_len, hi - lo);
    __syncthreads();
    const uint32_t len = s_len;
    XYZZ<F> acc = xyzz_inf<F>();
    uint32_t wnext = (!DIRECT && lo < hi) ? vals[lo] : 0u;
#pragma unroll 1
    for (uint32_t i = 0; i < len; ++i) {
        __syncthreads();
        const uint32_t j = lo + i;
        if (j < hi) {
            const uint32_t w = DIRECT ? j : wnext;
            Affine<F> pt = pts[w & 0x7fffffffu];
#if defined(C12_ACC_PREFETCH)
            // the index of the NEXT term is read one addition ahead and its point is pulled towards the SM while this one is added:
            // the block's warps reach their gathers together, so nothing else on the block hides that latency
            if (j + 1 < hi) {
                wnext = vals[j + 1];
                const char* nx = reinterpret_cast<const char*>(pts + (wnext & 0x7fffffffu));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(nx + sizeof(Affine<F>) - 4));
            }
#else
            if (!DIRECT && j + 1 < hi) wnext = vals[j + 1];
#endif
            if (!affine_is_inf(pt)) {
                if (!DIRECT && (w >> 31)) pt.y = neg(pt.y);
                xyzz_madd(acc, pt);
            }
        }
    }
    if (v != 0xffffffffu) vpartial[v] = xyzz_to_proj(acc);
#else
    static_assert(!DIRECT, "the direct lists need the lockstep build");
    if (t >= vmax) return;
    const uint32_t v = order[t];
    if (v == 0xffffffffu) return;                   // padding behind the last chunk
    const uint32_t b = vbucket[v];
    const uint32_t lo = start[b] + (v - vstart[b]) * chunk;
    uint32_t hi = lo + chunk;
    if (hi > end[b]) hi = end[b];
    vpartial[v] = msm_accumulate_range_body<F>(lo, hi, vals, pts);
#endif
}

// bucket b = the partial sums of its chunks, over all upload groups (group g's copy of the bucket is list g * total + b).
// A bucket with more than FOLD_SERIAL_MAX chunks (skewed scalars: a whole window in one bucket is thousands of chunks) is not
// folded by its one thread - that chain was 120 ms for 2^18 equal scalars over G2 - but handed to k_fold_heavy through a list.
constexpr uint32_t FOLD_SERIAL_MAX = 16;
template <class F>
__global__ void __launch_bounds__(128) k_fold(uint32_t total, uint32_t groups, const uint32_t* __restrict__ vstart, const Proj<F>* __restrict__ vpartial,
                                              Proj<F>* __restrict__ buckets, uint32_t* __restrict__ heavy)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total) return;
    const uint32_t v0 = vstart[b], nch = vstart[b + 1] - v0;
    if (heavy && groups == 1 && nch > FOLD_SERIAL_MAX) {
        heavy[1 + atomicAdd(heavy, 1u)] = b;          // heavy[0]: count, then the bucket ids
        return;
    }
    Proj<F> acc = msm_fold_body<F>(vpartial + v0, nch);
#pragma unroll 1
    for (uint32_t g = 1; g < groups; ++g) {
        const uint32_t v = vstart[g * total + b];
        acc = proj_add(acc, msm_fold_body<F>(vpartial + v, vstart[g * total + b + 1] - v));
    }
    buckets[b] = acc;
}

// level 0 of the bucket reduction: one thread per (window, segment of seg_len buckets)
template <class F>
__global__ void __launch_bounds__(128, sizeof(F) == sizeof(Fp) ? 3 : 1) k_reduce_level0(MsmPlan pl, const Proj<F>* __restrict__ buckets, Proj<F>* __restrict__ scratch, uint32_t w0)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t w = w0 + blockIdx.y;
    if (t >= pl.segs) return;
    uint32_t lo = t * pl.seg_len, hi = lo + pl.seg_len;
    if (hi > pl.half) hi = pl.half;
    Proj<F> sum, run;
    msm_reduce_level_body<F>(buckets + (size_t)w * pl.half, lo, hi, 1u, sum, run);
    scratch[msm_sum0_offset(pl, w) + t] = sum;
    scratch[msm_run1_offset(pl, w) + t] = run;
}

// ---- lane-cooperative point arithmetic (the serial tail of an MSM) -------------------------------------------------
// One thread's dependent chain of Montgomery products is latency-bound (~0.5 us per product), and the Horner
// combination of the window sums is c (W - 1) sequential doublings.  The RCB formulas have two layers of independent
// products (4 + 4 for a doubling, 6 + 6 for an addition): lane i of the warp computes the i-th product of a layer and
// the results are broadcast back with shuffles, so a doubling costs two product latencies instead of eight.
// Every lane holds the same point before and after each call (all 32 lanes must call these together).
// Groups of 8 lanes: every lane of a group holds the same point before and after each call; `mask` names the lanes of
// the calling group(s) (all of them must call together).
template <class T> __device__ __forceinline__ T group_bcast_obj(const T& x, int src, uint32_t mask)
{
    static_assert(sizeof(T) % 4 == 0, "word-sized objects only");
    T r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&x);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); ++i) d[i] = __shfl_sync(mask, s[i], src, 8);
    return r;
}

// out[i] = a[i] * b[i], product i computed by lane i of the group (N <= 8)
template <class F, int N> __device__ __forceinline__ void coop_mul(F (&out)[N], const F (&a)[N], const F (&b)[N], uint32_t mask)
{
    const int lane = threadIdx.x & 7;
    F x = a[0], y = b[0];
#pragma unroll
    for (int i = 1; i < N; ++i) {
        x = select(lane == i, a[i], x);
        y = select(lane == i, b[i], y);
    }
    F m = mul_hot(x, y);     // straight-line product: these are single-warp latency chains
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = group_bcast_obj(m, i, mask);
}

// WARP-WIDE variant (all 32 lanes hold the same operands; k_finish, k_fixed_base_windows).  Over Fp it is the 8-lane scheme.  Over
// Fp2 the N Karatsuba products are 3 N Fp products on 3 N lanes - lane 3 i + j computes a.a b.a / a.b b.b / (a.a + a.b)(b.a + b.b)
// of product i, lane 3 i collects the three and finishes the recombination - so a layer costs one Fp product latency instead
// of three (G2 Horner chain: ~10 us -> ~5 us per doubling).
template <int N> __device__ __forceinline__ void coop_mul_warp(Fp (&out)[N], const Fp (&a)[N], const Fp (&b)[N]) { coop_mul<Fp, N>(out, a, b, 0xffffffffu); }
template <int N> __device__ __forceinline__ void coop_mul_warp(Fp2 (&out)[N], const Fp2 (&a)[N], const Fp2 (&b)[N])
{
    static_assert(3 * N <= 32, "one lane per Fp product");
    const int lane = threadIdx.x & 31, i = lane / 3, j = lane - 3 * i;
    Fp2 xa = a[0], xb = b[0];
#pragma unroll
    for (int k = 1; k < N; ++k) {
        xa = select(i == k, a[k], xa);
        xb = select(i == k, b[k], xb);
    }
    const Fp x = fp_select(j == 0, xa.a, fp_select(j == 1, xa.b, fp_add(xa.a, xa.b)));
    const Fp y = fp_select(j == 0, xb.a, fp_select(j == 1, xb.b, fp_add(xb.a, xb.b)));
    const Fp m = fp_mul_inl(x, y);
    const Fp t1 = shfl_down_obj(m, 1), t2 = shfl_down_obj(m, 2);       // meaningful on lanes 3 i: m = a.a b.a, t1 = a.b b.b, t2 = the cross product
    const Fp2 r = Fp2{fp_sub(m, t1), fp_sub(fp_sub(t2, m), t1)};
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = shfl_bcast_obj(r, 3 * k);
}

// proj_dbl (RCB15 Algorithm 9), same value
template <class F, bool WARP = false> __device__ Proj<F> coop_dbl(const Proj<F>& p, uint32_t mask)
{
    F a1[4] = {p.y, p.y, p.z, p.x}, b1[4] = {p.y, p.z, p.z, p.y}, m[4];
    if constexpr (WARP)
        coop_mul_warp<4>(m, a1, b1);
    else
        coop_mul<F, 4>(m, a1, b1, mask);                // t0 = Y^2, t1 = YZ, t2 = Z^2, XY
    F z3 = mul8(m[0]);
    F t2 = FieldOps<F>::mul_b3(m[2]);
    F y3 = add(m[0], t2);
    F t0 = sub(m[0], mul3(t2));
    F a2[4] = {t2, z3, y3, t0}, b2[4] = {z3, m[1], t0, m[3]}, n[4];
    if constexpr (WARP)
        coop_mul_warp<4>(n, a2, b2);
    else
        coop_mul<F, 4>(n, a2, b2, mask);                // x3, Z3, y3 t0, t0 XY
    return Proj<F>{dbl(n[3]), add(n[2], n[0]), n[1]};
}

// proj_add (RCB15 Algorithm 7), same value
template <class F, bool WARP = false> __device__ Proj<F> coop_add(const Proj<F>& p, const Proj<F>& q, uint32_t mask)
{
    F a1[6] = {p.x, p.y, p.z, add(p.x, p.y), add(p.y, p.z), add(p.x, p.z)};
    F b1[6] = {q.x, q.y, q.z, add(q.x, q.y), add(q.y, q.z), add(q.x, q.z)}, m[6];
    if constexpr (WARP)
        coop_mul_warp<6>(m, a1, b1);
    else
        coop_mul<F, 6>(m, a1, b1, mask);
    F t3 = sub(m[3], add(m[0], m[1]));
    F t4 = sub(m[4], add(m[1], m[2]));
    F y3 = FieldOps<F>::mul_b3(sub(m[5], add(m[0], m[2])));
    F t0 = mul3(m[0]);
    F t2 = FieldOps<F>::mul_b3(m[2]);
    F z3 = add(m[1], t2);
    F t1 = sub(m[1], t2);
    F a2[6] = {y3, t3, y3, t1, t0, z3}, b2[6] = {t4, t1, t0, z3, t3, t4}, n[6];
    if constexpr (WARP)
        coop_mul_warp<6>(n, a2, b2);
    else
        coop_mul<F, 6>(n, a2, b2, mask);
    return Proj<F>{sub(n[1], n[0]), add(n[2], n[3]), add(n[5], n[4])};
}

// Sum of one value per thread over a block of THREADS (128 or 256), built for LATENCY: a thread's own complete addition is
// a chain of 14 dependent products (~14 us with one or two warps per scheduler), the lane-cooperative one ~3 us.  Two shuffle
// levels on whole threads (32 -> 8 per warp), then the remaining levels on groups of 8 lanes through shared memory.  Fixed
// tree: deterministic.  Result in sh[0] after the call (all threads must call; sh holds THREADS / 4 points).
template <class F, int THREADS> __device__ void block_sum_coop(Proj<F> acc, Proj<F>* sh)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    acc = proj_add(acc, shfl_down_obj(acc, 16));
    acc = proj_add(acc, shfl_down_obj(acc, 8));
    if (lane < 8) sh[w * 8 + lane] = acc;
    __syncthreads();
    const uint32_t g = tid >> 3, mask = 0xffu << (tid & 24);
#pragma unroll 1
    for (uint32_t n = THREADS / 8; n >= 1; n >>= 1) {
        if (g < n) {
            const Proj<F> r = coop_add(sh[g], sh[g + n], mask);
            if ((tid & 7) == 0) sh[g] = r;
        }
        __syncthreads();
    }
}

// the buckets k_fold left: one block per bucket at a time, every thread adds up a strided share of the chunk partials, the block's
// tree (block_sum_coop) the rest.  A group sum: any order gives the same bucket.  With no heavy bucket (random scalars) the blocks
// read the count and leave.
template <class F> __global__ void __launch_bounds__(128) k_fold_heavy(const uint32_t* __restrict__ heavy, const uint32_t* __restrict__ vstart,
                                                                          const Proj<F>* __restrict__ vpartial, Proj<F>* __restrict__ buckets)
{
    __shared__ Proj<F> sh[128 / 4];
    const uint32_t count = heavy[0];
#pragma unroll 1
    for (uint32_t h = blockIdx.x; h < count; h += gridDim.x) {
        const uint32_t b = heavy[1 + h], v0 = vstart[b], nch = vstart[b + 1] - v0;
        Proj<F> acc = proj_inf<F>();
#pragma unroll 1
        for (uint32_t j = threadIdx.x; j < nch; j += 128) acc = proj_add(acc, vpartial[v0 + j]);
        block_sum_coop<F, 128>(acc, sh);
        if (threadIdx.x == 0) buckets[b] = sh[0];
        __syncthreads();
    }
}

// the bit planes of the segment index (msm_core.cuh "bucket reduction"): blocks (j, w, z < splits) tree-sum plane j of window w
// (j = plane_bits: the sum of all sum0's, twice the entries, twice the blocks); the block that finishes last adds up the
// partials in block order -> planes[w][j].  Every plane is an independent sum: one launch, all in parallel, and sized to be
// resident at once (two blocks per SM: 256 threads at 128 registers over Fp, 128 threads at 255 over Fp2).
template <class F> struct PlaneShape {
    static constexpr int THREADS = sizeof(F) == sizeof(Fp) ? 256 : 128;
    static constexpr uint32_t SPLITS = sizeof(F) == sizeof(Fp) ? 2 : 4;      // blocks per bit plane (the plane of the sum0's takes twice as many)
};
template <class F>
__global__ void __launch_bounds__(PlaneShape<F>::THREADS, 2) k_reduce_planes(MsmPlan pl, const Proj<F>* __restrict__ scratch, Proj<F>* __restrict__ parts,
                                                                            uint32_t* __restrict__ tickets, Proj<F>* __restrict__ planes, uint32_t w0)
{
    constexpr int THREADS = PlaneShape<F>::THREADS;
    constexpr uint32_t SPLITS = PlaneShape<F>::SPLITS;
    __shared__ Proj<F> sh[THREADS / 4];
    __shared__ uint32_t s_last;
    const uint32_t j = blockIdx.x, w = w0 + blockIdx.y, z = blockIdx.z;
    const uint32_t splits = j == pl.plane_bits ? 2 * SPLITS : SPLITS;
    if (z >= splits) return;
    Proj<F> acc = msm_plane_slice_body<F>(pl, scratch + msm_sum0_offset(pl, w), scratch + msm_run1_offset(pl, w), j, z * THREADS + threadIdx.x,
                                          splits * THREADS);
    block_sum_coop<F, THREADS>(acc, sh);
    const size_t slot = (size_t)w * MSM_WPART_SLOTS + j;
    Proj<F>* mine = parts + slot * (2 * SPLITS);
    if (threadIdx.x == 0) {
        mine[z] = sh[0];
        __threadfence();
        s_last = atomicAdd(&tickets[slot], 1u) == splits - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 8) return;
    __threadfence();
    Proj<F> sum = mine[0];
#pragma unroll 1
    for (uint32_t k = 1; k < splits; ++k) sum = coop_add(sum, mine[k], 0xffu);
    if (threadIdx.x == 0) {
        planes[slot] = sum;
        tickets[slot] = 0;          // ready for the next call on this arena layout
    }
}

template <class F> __device__ void write_point(uint8_t* out, const Proj<F>& r, int out_mode)
{
    Affine<F> a = proj_to_affine(r);
    if (out_mode == OUT_AFFINE)
        Wire<F>::serialize(out, a);
    else
        Wire<F>::compress(out, a);
}

// One block of 8 warps over the windows [w_lo, w_hi).  Warp v, for windows w_lo + v, + 8, ...: recombines the planes
// S_w = T + L (P_0 + 2 (P_1 + ...))  with lane-cooperative doublings and additions.  Then warp 0 runs the Horner combination
// acc = 2^c acc + S_w  down to w_lo, multiplies by 2^(c shift) (the windows below a HIGH part), adds add_in (the high part's
// result, for the low part), and either parks the point (part_out) or normalises and encodes it.  One launch over all windows
// with shift = 0 is the whole tail; two launches let the high windows' chain of doublings - c (W - 1) of them whatever the
// split - run while the low windows are still being added up (msm_run, "split tail").
template <class F>
__global__ void __launch_bounds__(256) k_finish(MsmPlan pl, const Proj<F>* __restrict__ planes, uint8_t* out, int out_mode, uint32_t w_lo, uint32_t w_hi,
                                                uint32_t shift, const Proj<F>* __restrict__ add_in, Proj<F>* __restrict__ part_out)
{
    extern __shared__ __align__(16) unsigned char finish_smem[];
    Proj<F>* wsum = reinterpret_cast<Proj<F>*>(finish_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t full = 0xffffffffu;
    for (uint32_t w = w_lo + warp; w < w_hi; w += 8) {
        const Proj<F>* slot = planes + (size_t)w * MSM_WPART_SLOTS;
        Proj<F> acc = slot[pl.plane_bits];
        if (pl.plane_bits > 0) {
            acc = slot[pl.plane_bits - 1];
#pragma unroll 1
            for (uint32_t j = pl.plane_bits - 1; j > 0; --j) {
                acc = coop_dbl<F, true>(acc, full);
                acc = coop_add<F, true>(acc, slot[j - 1], full);
            }
#pragma unroll 1
            for (uint32_t len = pl.seg_len; len > 1; len >>= 1) acc = coop_dbl<F, true>(acc, full);
            acc = coop_add<F, true>(acc, slot[pl.plane_bits], full);
        }
        if (lane == 0) wsum[w - w_lo] = acc;
    }
    __syncthreads();
    if (warp) return;
    Proj<F> acc = wsum[w_hi - 1 - w_lo];
#pragma unroll 1
    for (uint32_t w = w_hi - 1; w > w_lo; --w) {
#pragma unroll 1
        for (uint32_t k = 0; k < pl.c; ++k) acc = coop_dbl<F, true>(acc, full);
        acc = coop_add<F, true>(acc, wsum[w - 1 - w_lo], full);
    }
#pragma unroll 1
    for (uint32_t k = 0; k < shift * pl.c; ++k) acc = coop_dbl<F, true>(acc, full);
    if (add_in) acc = coop_add<F, true>(acc, *add_in, full);
    if (part_out) {
        if (lane == 0) *part_out = acc;
    } else if (lane == 0) {
        write_point<F>(out, acc, out_mode);
    }
}

// the two halves of a split tail -> the result (one warp, lane-cooperative addition)
template <class F> __global__ void __launch_bounds__(32) k_combine2(const Proj<F>* __restrict__ a, const Proj<F>* __restrict__ b, uint8_t* out, int out_mode)
{
    const Proj<F> r = coop_add<F, true>(*a, *b, 0xffffffffu);
    if ((threadIdx.x & 31) == 0) write_point<F>(out, r, out_mode);
}

// sum of n wire-format points (merging all-gathered per-rank partials; also a general point-sum entry)
template <class F>
__global__ void __launch_bounds__(256) k_sum_points(const uint8_t* __restrict__ bytes, uint32_t n, uint8_t* out, int out_mode, int* flags)
{
    Proj<F> acc = proj_inf<F>();
    for (uint32_t i = threadIdx.x; i < n; i += 256) {
        Affine<F> p;
        if (!Wire<F>::parse(p, bytes + (size_t)Wire<F>::AFFINE * i)) atomicOr(flags, FLAG_BAD_POINT);
        acc = proj_add_affine(acc, p);
    }
    acc = block_sum_256(acc);
    if (threadIdx.x == 0) write_point<F>(out, acc, out_mode);
}

template <class F>
__global__ void __launch_bounds__(128) k_scalar_mul(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ scalars, uint32_t n,
                                                    uint8_t* __restrict__ out, int out_mode, int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Scalar256 k = scalar_from_be32(scalars + 32ull * i);
    if (!scalar_is_canonical(k)) atomicOr(flags, FLAG_BAD_SCALAR);
    const size_t stride = out_mode == OUT_AFFINE ? Wire<F>::AFFINE : Wire<F>::COMPRESSED;
    if (!scalar_mul_body<F>(pts + (size_t)Wire<F>::AFFINE * i, scalars + 32ull * i, out + stride * i, out_mode == OUT_AFFINE))
        atomicOr(flags, FLAG_BAD_POINT);
}

// ---- fixed base g^x (the `select` path, g1_point.hpp:355-369 / g2_point.hpp:129-143) ------------------------------
// A per-context table  T[w][d-1] = d 2^(8w) G  (w < 32, d = 1..128, affine, Montgomery form) is built once on the
// device; a scalar is recoded into 32 signed 8-bit digits and g^x is 32 mixed additions - no doublings.
constexpr uint32_t FB_WINDOWS = 32, FB_HALF = 128;

// Table construction in two steps.  k_fixed_base_windows: warp j -> W[j][w] = 2^(8w) base_j for w = 0..31 (the only long
// dependent chain: 248 lane-cooperative doublings).  k_fixed_base_table: thread (j, w, d) -> d W[j][w] by
// double-and-add over the 8 bits of d (at most 7 + 7 point operations), normalised to affine.
// bases = nullptr: the default generator (one table); else m wire-format affine bases, table j at [j * 4096].
template <class F>
__global__ void __launch_bounds__(32) k_fixed_base_windows(const uint8_t* __restrict__ bases, uint32_t m, Proj<F>* __restrict__ wbase, int* flags)
{
    // one warp per base: W[j][w] = 2^(8w) base_j is ONE chain of 248 doublings, so it runs on the lane-cooperative doubling
    // (two product latencies each, as in k_finish) instead of 32 threads redoing ever longer prefixes of it on their own
    const uint32_t j = blockIdx.x, lane = threadIdx.x;
    Affine<F> base = generator<F>();
    if (bases && !Wire<F>::parse(base, bases + (size_t)Wire<F>::AFFINE * j) && lane == 0) atomicOr(flags, FLAG_BAD_POINT);
    Proj<F> acc = proj_from_affine(base);
#pragma unroll 1
    for (uint32_t w = 0; w < FB_WINDOWS; ++w) {
        if (lane == 0) wbase[j * FB_WINDOWS + w] = acc;
        if (w + 1 < FB_WINDOWS) {
#pragma unroll 1
            for (int i = 0; i < 8; ++i) acc = coop_dbl<F, true>(acc, 0xffffffffu);
        }
    }
}

template <class F>
__global__ void __launch_bounds__(128) k_fixed_base_table(uint32_t m, const Proj<F>* __restrict__ wbase, Affine<F>* __restrict__ table)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * FB_WINDOWS * FB_HALF) return;
    const uint32_t d = t % FB_HALF + 1;
    const Proj<F> base = wbase[t / FB_HALF];
    table[t] = proj_to_affine(proj_mul_small(base, d));
}

// acc += k * (the base of `table`), k < r as 32 signed 8-bit digits: 32 mixed additions, no doublings
template <class F> __device__ __forceinline__ void fixed_base_accumulate(XYZZ<F>& acc, const Scalar256& k, const Affine<F>* __restrict__ table)
{
    uint32_t carry = 0;
#pragma unroll 1
    for (uint32_t w = 0; w < FB_WINDOWS; ++w) {
        uint32_t d = ((k.v[w >> 2] >> ((w & 3u) * 8)) & 255u) + carry;
        carry = 0;
        bool ng = false;
        if (d > FB_HALF) {          // (128, 256] -> d - 256 in (-128, 0]
            d = 256u - d;
            ng = true;
            carry = 1;
        }
        if (d == 0) continue;
        Affine<F> pt = table[w * FB_HALF + d - 1];
        if (affine_is_inf(pt)) continue;     // identity base
        if (ng) pt.y = neg(pt.y);
        xyzz_madd(acc, pt);
    }
    // scalars are < r < 2^255, so the top digit is at most 0x73 + 1 and never carries out
}

// out[b] = sum_j scalars[b][j] * base_j  (m = 1 with the generator's table: g^x)
template <class F>
__global__ void __launch_bounds__(128) k_fixed_base(const uint8_t* __restrict__ scalars, uint32_t n, uint32_t m, const Affine<F>* __restrict__ table,
                                                    uint8_t* __restrict__ out, int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ<F> acc = xyzz_inf<F>();
#pragma unroll 1
    for (uint32_t j = 0; j < m; ++j) {
        Scalar256 k = scalar_from_be32(scalars + 32ull * ((size_t)i * m + j));
        if (!scalar_is_canonical(k)) atomicOr(flags, FLAG_BAD_SCALAR);
        fixed_base_accumulate<F>(acc, k, table + (size_t)j * FB_WINDOWS * FB_HALF);
    }
    Wire<F>::serialize(out + (size_t)Wire<F>::AFFINE * i, proj_to_affine(xyzz_to_proj(acc)));
}

// ---- wire-format conversions in batch (K13): compressed <-> affine -----------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_decompress(const uint8_t* __restrict__ in, uint32_t n, uint8_t* __restrict__ out, int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p;
    if (!Wire<F>::decompress(p, in + (size_t)Wire<F>::COMPRESSED * i)) atomicOr(flags, FLAG_BAD_POINT);
    Wire<F>::serialize(out + (size_t)Wire<F>::AFFINE * i, p);
}
template <class F>
__global__ void __launch_bounds__(128) k_compress(const uint8_t* __restrict__ in, uint32_t n, uint8_t* __restrict__ out, int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p;
    if (!Wire<F>::parse(p, in + (size_t)Wire<F>::AFFINE * i)) atomicOr(flags, FLAG_BAD_POINT);
    Wire<F>::compress(out + (size_t)Wire<F>::COMPRESSED * i, p);
}

// ---- host pipeline ------------------------------------------------------------------------------------------
int sort_pairs_segmented(uint32_t*& keys, uint32_t*& vals, uint32_t*& keys_alt, uint32_t*& vals_alt, uint32_t n, uint32_t nseg,
                         uint32_t key_bits, uint32_t* hist, uint32_t* tile_sums, cudaStream_t s);
size_t sort_scratch_words(uint32_t n, uint32_t nseg, size_t* tile_words);

// Schedule of the batch-affine halving rounds: how many rounds; round 0 runs per upload GROUP (a group's sorted entries need
// only that group's points, so its additions start while the later groups are still on the wire), rounds >= 1 run on the merged
// lists - a bucket's level-1 list is the concatenation of its groups' - in PIPELINES (window groups on their own streams: one's
// inversion kernel hides behind the other's additions).  For every lane the slots per thread J and the launch bound in warps.
// Everything is derived on the host from UPPER bounds (a window segment holds at most lp.n entries; a level's slot count is at
// most entries / 2^k + a few per bucket); the kernels read the exact slot ranges from the offset arrays on the device.
#ifndef C12_BA_AUTO_MIN_LOG
#define C12_BA_AUTO_MIN_LOG 22      // automatic rounds only from 2^22 (term, window) entries on: below, a round's launches and its inversion latency cost more than they save
#endif
// Ctx::knob: [0] J is chosen so that a lane's round still spans about this many waves of resident warps; [1] its upper
// limit; [2] the automatic round count stops this many halvings early - lists of a few entries are cheaper to finish with
// XYZZ additions than with rounds that are all launch and inversion latency
struct BaSchedule {
    uint32_t rounds = 0, groups = 1, pipes = 1, lanes = 1;
    bool split_tail = false;
    uint32_t b_lo[BA_MAX_PIPES + 1];                            // pipeline p owns the real buckets [b_lo[p], b_lo[p + 1])
    uint32_t J0[BA_MAX_PIPES], warps0[BA_MAX_PIPES];            // round 0, per group
    uint32_t region0[BA_MAX_PIPES];                             // a group's region in list buffer 0 / the prefix array / the round-0 slot references
    uint32_t slots0 = 0;                                        // bound on one group's level-1 slot count
    uint32_t J[BA_MAX_ROUNDS][BA_MAX_PIPES], warps[BA_MAX_ROUNDS][BA_MAX_PIPES];   // rounds >= 1, per pipeline
    uint32_t ref_region[BA_MAX_ROUNDS][BA_MAX_PIPES];           // rounds >= 1: where (round, pipeline)'s slot references start
    uint32_t region[2][BA_MAX_PIPES];                           // rounds >= 1: pipeline p's region in list buffer (round & 1)
    uint32_t scratch_region[BA_MAX_PIPES];                      // rounds >= 1: pipeline p's region in the prefix array
    uint32_t final_region[BA_MAX_PIPES];                        // where the last round writes (buffer 2; one round only: buffer 0 by group)
    uint32_t slots_merged = 0;                                  // bound on the level-2 slot count (all pipelines)
    // list buffers: [0] level 1 (round 0's output, by group), [1] / [3] the odd / even rounds >= 1 (by pipeline), [2] the last
    // round.  The pipelines run rounds apart from each other, so no buffer a pipeline still reads is ever another one's target.
    size_t buf[4] = {0, 0, 0, 0}, prefix = 0, refs = 0, pool_stride = 0;
};

inline BaSchedule msm_ba_schedule(const MsmPlan& pl, const MsmPlan& lp, uint32_t blocks_per_sm, uint32_t auto_min_log = C12_BA_AUTO_MIN_LOG)
{
    BaSchedule sc;
    const Ctx& c = ctx();
    const uint64_t N = (uint64_t)lp.n * lp.windows;
    if (c.ba_rounds == 0 || N >= (1ull << 31) || pl.total == 0) return sc;
    uint32_t R = 0;
    while (R < 12 && (1ull << R) * pl.total < 2 * N) ++R;      // lists of twice the mean load would end up as single sums
    if (c.ba_rounds > 0)
        R = (uint32_t)c.ba_rounds;
    else if (N < (1ull << auto_min_log))
        return sc;
    else
        R = R > (uint32_t)c.knob[2] ? R - (uint32_t)c.knob[2] : 0;
    if (R == 0) return sc;
    sc.rounds = R > BA_MAX_ROUNDS ? BA_MAX_ROUNDS : R;
    sc.groups = lp.groups;
    sc.pipes = (uint32_t)c.ba_pipes < pl.windows ? (uint32_t)c.ba_pipes : pl.windows;
    if (sc.pipes < 1) sc.pipes = 1;
    sc.lanes = sc.groups > sc.pipes ? sc.groups : sc.pipes;
    // one lane's round is sized for the warps that are resident at once: BaShape<F>::MIN_BLOCKS blocks per SM (4 over Fp, 2 over
    // Fp2; a fixed 3 until r03o: accumulation -1 % over G1, -2 % over G2 at n = 2^20, profiles/r03o_fill_sweep.txt)
    const uint32_t resident = (uint32_t)(c.sm_count > 0 ? c.sm_count : 148) * (C12_BA_THREADS / 32) * blocks_per_sm;
    const uint64_t jmax = c.knob[1] > 2 ? (uint64_t)c.knob[1] : 2, waves = (uint64_t)(c.knob[0] > 0 ? c.knob[0] : 1);
    auto shape = [&](size_t slots, uint32_t& J, uint32_t& warps) {
        uint64_t j = slots * 100ull / (32ull * resident * waves * (uint64_t)(c.ba_fill_pct > 0 ? c.ba_fill_pct : 100));
        j = j < 2 ? 2 : (j > jmax ? jmax : j);
        J = (uint32_t)j;
        warps = (uint32_t)((slots + 32 * j - 1) / (32 * j));
        if (warps + 1 > sc.pool_stride) sc.pool_stride = warps + 1;
    };
    for (uint32_t p = 0; p <= sc.pipes; ++p) sc.b_lo[p] = (pl.windows * p / sc.pipes) * pl.half;
    // split tail (msm_run): two pipelines, the HIGH windows the smaller one (3 of 8, 1 of 4) - it finishes its rounds first and its
    // reduction and chain of doublings run under the low windows' last rounds
    sc.split_tail = c.split_tail == 1 && sc.rounds > 1 && sc.pipes == 2 && pl.windows >= 4;
    if (sc.split_tail) {
        const uint32_t high = pl.windows * 3 / 8 > 0 ? pl.windows * 3 / 8 : 1;
        sc.b_lo[1] = (pl.windows - high) * pl.half;
    }
    // round 0: group g's level-1 lists hold at most (its entries) / 2 + one per bucket slots
    const size_t cap0 = (size_t)(((uint64_t)pl.windows * lp.n) >> 1) + pl.total + 1;
    sc.slots0 = (uint32_t)cap0;
    for (uint32_t g = 0; g < sc.groups; ++g) {
        sc.region0[g] = (uint32_t)(g * cap0);
        shape(cap0, sc.J0[g], sc.warps0[g]);
    }
    // level k >= 2 of pipeline p: ceil(M1_b / 2^(k-1)) <= m_b / 2^k + groups + 1 per bucket
    auto cap = [&](uint32_t k, uint32_t p) {
        const uint64_t wins = (sc.b_lo[p + 1] - sc.b_lo[p]) / pl.half;
        return (size_t)((wins * lp.n * lp.groups) >> k) + (size_t)wins * pl.half * (lp.groups + 1) + 1;
    };
    size_t ref_base = cap0 * sc.groups;
    for (uint32_t r = 1; r < sc.rounds; ++r) {
        size_t base = 0;
        for (uint32_t p = 0; p < sc.pipes; ++p) {
            shape(cap(r + 1, p), sc.J[r][p], sc.warps[r][p]);
            sc.ref_region[r][p] = (uint32_t)(ref_base + base);
            base += cap(r + 1, p);
        }
        if (r == 1) sc.slots_merged = (uint32_t)base;
        ref_base += base;
    }
    sc.refs = ref_base;
    size_t lvl[4] = {0, 0, 0, 0};           // total capacity of levels 2, 3 and the last
    for (uint32_t k = 2; k <= 3; ++k) {
        size_t base = 0;
        for (uint32_t p = 0; p < sc.pipes; ++p) {
            sc.region[(k - 1) & 1][p] = (uint32_t)base;           // level k is the output of round k - 1: buffer (k - 1) & 1
            if (k == 2) sc.scratch_region[p] = (uint32_t)base;
            base += cap(k, p);
        }
        lvl[k] = base;
    }
    size_t fin = 0;
    for (uint32_t p = 0; p < sc.pipes; ++p) {
        sc.final_region[p] = (uint32_t)fin;
        fin += cap(sc.rounds, p);
    }
    sc.buf[0] = cap0 * sc.groups;
    sc.buf[1] = sc.rounds > 1 ? lvl[2] : 1;
    sc.buf[2] = sc.rounds > 1 ? fin : 1;
    sc.buf[3] = sc.rounds > 2 ? lvl[3] : 1;
    sc.prefix = cap0 * sc.groups > lvl[2] ? cap0 * sc.groups : lvl[2];
    return sc;
}

// lp: the list-side plan (msm_list_plan), pl: the plain plan of the reduction
template <class F> size_t msm_scratch_bytes(const MsmPlan& pl, const MsmPlan& lp)
{
    size_t N = (size_t)lp.n * lp.windows;
    size_t tile_words = 0;
    size_t hist_words = sort_scratch_words(lp.n, lp.windows, &tile_words);
    size_t b = 0;
    b += align_up(sizeof(Affine<F>) * (size_t)pl.n);
    b += 4 * align_up(4 * N);
    b += align_up(4 * hist_words) + align_up(4 * (tile_words + 16));
    b += align_up(4 * (size_t)BA_MAX_PIPES * count_scan_scratch_words(pl.total));
    b += 2 * align_up(4 * (size_t)lp.total);
    b += 2 * (align_up(4 * chunk_order_scratch_words(pl)) + align_up(sizeof(Proj<F>) * (size_t)pl.vmax)) + align_up(2 * sizeof(Proj<F>));      // twice: the split tail orders and accumulates each pipeline's buckets on their own
    b += align_up(sizeof(Proj<F>) * (size_t)pl.total);
    const BaSchedule sc = msm_ba_schedule(pl, lp, BaShape<F>::MIN_BLOCKS, BaShape<F>::AUTO_MIN_LOG);
    if (sc.rounds) {
        // level offsets: one level-1 array per group, levels 2 .. rounds merged; the plan scratch of both; the list bounds
        b += align_up(4 * (size_t)(sc.groups + sc.rounds) * ((size_t)pl.total + 1));
        b += align_up(4 * (sc.groups + 1) * ba_plan_scratch_words(pl.total, sc.rounds));
        b += 2 * align_up(4 * ((size_t)pl.total + 1));
        for (int k = 0; k < 4; ++k) b += align_up(sizeof(Affine<F>) * sc.buf[k]);
        b += align_up(sizeof(F) * sc.prefix) + align_up(8 * sc.refs);
        b += align_up(sizeof(F) * sc.pool_stride * sc.lanes) + align_up(sizeof(F) * sc.pool_stride * sc.lanes * 32);
    }
    b += align_up(sizeof(Proj<F>) * msm_reduce_scratch_points(pl));
    b += align_up(sizeof(Proj<F>) * (size_t)pl.windows * MSM_WPART_SLOTS) * (1 + 2 * PlaneShape<F>::SPLITS) + align_up(4 * ((size_t)pl.windows * MSM_WPART_SLOTS + 8));
    b += align_up(4 * 2 * ((size_t)pl.total + 1));
    return b + 65536;
}

// one lane of a halving round
template <class F> struct BaLaneArgs {
    cudaStream_t stream;
    BaGeom g;
    BaIo<F> io;
    uint32_t warps;
    bool first;
};

template <class F> int msm_ba_round_launch(const BaLaneArgs<F>& a)
{
    const unsigned blocks = cdiv(a.warps, C12_BA_THREADS / 32);
    if (a.first)
        k_ba_fwd<F, true><<<blocks, C12_BA_THREADS, 0, a.stream>>>(a.g, a.io);
    else
        k_ba_fwd<F, false><<<blocks, C12_BA_THREADS, 0, a.stream>>>(a.g, a.io);
    C12_LAUNCHED();
    k_ba_inv<F><<<cdiv(a.warps, 128), 128, 0, a.stream>>>(a.g, a.io.pool);
    C12_LAUNCHED();
    if (a.first)
        k_ba_bwd<F, true><<<blocks, C12_BA_THREADS, 0, a.stream>>>(a.g, a.io);
    else
        k_ba_bwd<F, false><<<blocks, C12_BA_THREADS, 0, a.stream>>>(a.g, a.io);
    C12_LAUNCHED();
    return C12381_OK;
}

// the plans of an n-term MSM under the current settings: window width, then the list-side plan for `groups` upload groups
// (groups only exist where at least two halving rounds run: round 0 is what makes a group's lists independent of the later
// groups' points, the merged rounds behind it are what keeps the groups from costing anything)
template <class F> bool msm_plans(size_t n, uint32_t groups, MsmPlan& pl, MsmPlan& lp)
{
    if (n == 0 || n > (1ull << 26) / MsmTraits<F>::PARTS * 2) return false;
    constexpr uint32_t parts = MsmTraits<F>::PARTS;
    uint32_t cbits = ctx().forced_window ? (uint32_t)ctx().forced_window : msm_choose_window(parts * n, 256 / parts);
    if (cbits < 2 || cbits > 16) return false;
    // segment running sums of the bucket reduction: over Fp2 one block per SM is resident (148 x 128 threads), so the grid is kept
    // within 16 Ki threads - one wave - there (G2 n = 2^18: tail 2.45 -> 2.22 ms, profiles/r02l_tail_tune.txt)
    const uint32_t seg_wave = ctx().knob[3] > 0 ? (uint32_t)ctx().knob[3] : (sizeof(F) == sizeof(Fp2) ? 16384u : 0u);
    pl = msm_make_plan((uint32_t)n, cbits, parts, seg_wave);
    if (groups > BA_MAX_PIPES) groups = BA_MAX_PIPES;
    lp = msm_list_plan(pl, groups);
    if (groups > 1 && msm_ba_schedule(pl, lp, BaShape<F>::MIN_BLOCKS, BaShape<F>::AUTO_MIN_LOG).rounds < 2) lp = pl;
    return true;
}

// scratch bound for an n-term MSM under the current settings (callers arena_begin with at least this much)
template <class F> size_t msm_scratch_for(size_t n, uint32_t groups = 1)
{
    MsmPlan pl, lp;
    if (!msm_plans<F>(n, groups, pl, lp)) return 0;
    return msm_scratch_bytes<F>(pl, lp);
}

// The caller has arena_begin()'d at least msm_scratch_for<F>(n, groups) bytes beyond what it took itself.
// d_points: n wire-format affine points; d_scalars: n x 32 B big-endian; d_out: Wire<F>::COMPRESSED or ::AFFINE bytes.
// Everything is enqueued on `s` (and the context's side streams, forked from and joined back into `s`); malformed input is
// reported through the context flag word (c12381_sync_status / the host entry's return code), never by a different code path.
// scalars_ready / points_ready: optional events, one per upload group, after which that group's slice of d_scalars / d_points is
// valid.  The host entries upload group by group on a second stream; every group is a lane of its own from its scalars to its
// level-1 lists (recode, sort, bounds, slot map; then parse and round 0), so group 0's additions start while the later groups are
// still on the wire; the rounds behind that run on the merged lists.  The groups are ceil(n / groups) consecutive terms each.
template <class F>
int msm_run(const uint8_t* d_points, const uint8_t* d_scalars, size_t n_sz, uint8_t* d_out, int out_mode, cudaStream_t s,
            uint32_t groups = 1, const cudaEvent_t* scalars_ready = nullptr, const cudaEvent_t* points_ready = nullptr)
{
    Ctx& c = ctx();
    const int out_bytes = out_mode == OUT_AFFINE ? Wire<F>::AFFINE : Wire<F>::COMPRESSED;
    if (n_sz == 0) {
        C12_CUDA(cudaMemsetAsync(d_out, 0, out_bytes, s));
        return C12381_OK;
    }
    if (n_sz > (1ull << 26) / MsmTraits<F>::PARTS * 2) return set_error(C12381_EARG, "msm: too many terms per call (2^26 over G1, 2^25 over G2)");
    const uint32_t n = (uint32_t)n_sz;
    MsmPlan pl, lp;
    if (!msm_plans<F>(n_sz, groups, pl, lp)) return set_error(C12381_EARG, "msm: window width must be in [2, 16]");
    const uint32_t groups_req = groups;
    groups = lp.groups;
    const size_t N = (size_t)lp.n * lp.windows;
    const uint32_t B = pl.total;                                // real buckets; group g's copy of bucket b is list g * B + b

    int rc;
    const size_t hist_words = sort_scratch_words(lp.n, lp.windows, nullptr);
    size_t gtile_words = 0;
    const size_t ghist_words = sort_scratch_words(lp.n, pl.windows, &gtile_words);      // one group's segments
    Affine<F>* pts = (Affine<F>*)arena_take(sizeof(Affine<F>) * (size_t)pl.n);
    uint32_t* keys = (uint32_t*)arena_take(4 * N);
    uint32_t* vals = (uint32_t*)arena_take(4 * N);
    uint32_t* keys2 = (uint32_t*)arena_take(4 * N);
    uint32_t* vals2 = (uint32_t*)arena_take(4 * N);
    uint32_t* hist = (uint32_t*)arena_take(4 * hist_words);
    uint32_t* tiles = (uint32_t*)arena_take(4 * (size_t)groups * (gtile_words + 2));
    const size_t count_words = count_scan_scratch_words(B);
    uint32_t* count_tiles = (uint32_t*)arena_take(4 * (size_t)groups * count_words);
    uint32_t* start = (uint32_t*)arena_take(4 * (size_t)lp.total);
    uint32_t* end = (uint32_t*)arena_take(4 * (size_t)lp.total);
    uint32_t* chunk_scratch = (uint32_t*)arena_take(4 * chunk_order_scratch_words(pl));
    Proj<F>* vpartial = (Proj<F>*)arena_take(sizeof(Proj<F>) * (size_t)pl.vmax);
    uint32_t* chunk_scratch2 = (uint32_t*)arena_take(4 * chunk_order_scratch_words(pl));      // split tail: the second pipeline's
    Proj<F>* vpartial2 = (Proj<F>*)arena_take(sizeof(Proj<F>) * (size_t)pl.vmax);
    Proj<F>* hpart = (Proj<F>*)arena_take(2 * sizeof(Proj<F>));
    Proj<F>* buckets = (Proj<F>*)arena_take(sizeof(Proj<F>) * (size_t)pl.total);
    const BaSchedule sc = msm_ba_schedule(pl, lp, BaShape<F>::MIN_BLOCKS, BaShape<F>::AUTO_MIN_LOG);
    const uint32_t R = sc.rounds;
    uint32_t *off1 = nullptr, *offm = nullptr, *ba_tiles = nullptr, *ba_lstart = nullptr, *ba_lend = nullptr;
    uint2* ba_refs = nullptr;
    Affine<F>* E[4] = {nullptr, nullptr, nullptr, nullptr};    // list buffers (BaSchedule::buf)
    F *ba_prefix = nullptr, *ba_pool = nullptr, *ba_others = nullptr;
    const size_t lvl_words = (size_t)B + 1, plan_words = ba_plan_scratch_words(B, R ? R : 1);
    if (R) {
        off1 = (uint32_t*)arena_take(4 * lvl_words * groups);                  // level 1, per group
        offm = (uint32_t*)arena_take(4 * lvl_words * R);                       // levels 2 .. R, merged: level k at (k - 2) * lvl_words
        ba_tiles = (uint32_t*)arena_take(4 * plan_words * (groups + 1));
        ba_lstart = (uint32_t*)arena_take(4 * lvl_words);
        ba_lend = (uint32_t*)arena_take(4 * lvl_words);
        for (int k = 0; k < 4; ++k) E[k] = (Affine<F>*)arena_take(sizeof(Affine<F>) * sc.buf[k]);
        ba_prefix = (F*)arena_take(sizeof(F) * sc.prefix);
        ba_refs = (uint2*)arena_take(8 * sc.refs);
        ba_pool = (F*)arena_take(sizeof(F) * sc.pool_stride * sc.lanes);
        ba_others = (F*)arena_take(sizeof(F) * sc.pool_stride * sc.lanes * 32);
    }
    Proj<F>* partial = (Proj<F>*)arena_take(sizeof(Proj<F>) * msm_reduce_scratch_points(pl));
    Proj<F>* wsum = (Proj<F>*)arena_take(sizeof(Proj<F>) * (size_t)pl.windows * MSM_WPART_SLOTS);
    Proj<F>* plane_parts = (Proj<F>*)arena_take(sizeof(Proj<F>) * (size_t)pl.windows * MSM_WPART_SLOTS * 2 * PlaneShape<F>::SPLITS);
    uint32_t* plane_tickets = (uint32_t*)arena_take(4 * ((size_t)pl.windows * MSM_WPART_SLOTS + 8));      // + the two counts of the heavy-bucket lists (k_fold)
    uint32_t* heavy_lists = (uint32_t*)arena_take(4 * 2 * ((size_t)pl.total + 1));
    if (!heavy_lists) return set_error(C12381_ECUDA, "msm: scratch arena bound too small");
    uint32_t* heavy[2] = {heavy_lists, heavy_lists + pl.total + 1};
    C12_CUDA(cudaMemsetAsync(plane_tickets, 0, 4 * (size_t)pl.windows * MSM_WPART_SLOTS, s));
    C12_CUDA(cudaMemsetAsync(heavy[0], 0, 4, s));
    C12_CUDA(cudaMemsetAsync(heavy[1], 0, 4, s));

    C12_CUDA(cudaEventRecord(c.ev[0], s));
    C12_CUDA(cudaEventRecord(c.pev[0], s));
    // Streams: group g runs on `s` (g = 0) or side stream g - 1; with several groups the merged plan runs on the plan stream so
    // that it does not hold up group 0's round 0; pipeline p of the merged rounds on `s` (p = 0) or side stream p - 1.
    auto lane_stream = [&](uint32_t i) { return i ? c.side[i - 1] : s; };
    cudaStream_t plan_stream = groups > 1 ? c.plan_stream : s;      // the merged plan needs only the bucket COUNTS: it runs beside the scatter / the groups' own stages
    if (groups > 1) {           // the arena is ours from this point of `s` on
        C12_CUDA(cudaEventRecord(c.side_ev[0], s));
        for (uint32_t g = 1; g < groups; ++g) C12_CUDA(cudaStreamWaitEvent(c.side[g - 1], c.side_ev[0], 0));
        C12_CUDA(cudaStreamWaitEvent(plan_stream, c.side_ev[0], 0));
    }
    // One group: the scalar-only stages (which wait on the L2 atomic unit and on launch latencies) run on the high-priority front
    // stream, the parse of the points (integer pipe) beside them on `s`, filling what they leave idle.
    const bool parse_aside = groups == 1 && c.parse_aside;
    if (parse_aside) {
        C12_CUDA(cudaEventRecord(c.parse_ev[0], s));
        C12_CUDA(cudaStreamWaitEvent(c.front_stream, c.parse_ev[0], 0));
        plan_stream = c.plan_stream;
        C12_CUDA(cudaStreamWaitEvent(plan_stream, c.parse_ev[0], 0));
    }
    // ---- per group: the scalar-only stages (recode, sort, bucket bounds, level-1 offsets, round-0 slot map) -----------------
    uint32_t *skeys = keys, *svals = vals;       // where the sorted pairs end up (the passes ping-pong between the two buffers)
    for (uint32_t g = 0; g < groups; ++g) {
        cudaStream_t sg = parse_aside ? c.front_stream : lane_stream(g);
        if (scalars_ready) {
            if (groups > 1)
                C12_CUDA(cudaStreamWaitEvent(sg, scalars_ready[g], 0));
            else
                for (uint32_t q = 0; q < groups_req; ++q) C12_CUDA(cudaStreamWaitEvent(sg, scalars_ready[q], 0));
        }
        const size_t seg_off = (size_t)g * pl.windows * lp.n;
        uint32_t* off1_g = R ? off1 + g * lvl_words : nullptr;
        if (c.front_end == 0) {
            // bucket lists by counting: (bucket | sign, rank) in keys / vals, the lists in vals2
            rc = launch_bucket_lists(lp, groups > 1 ? (int)g : -1, d_scalars, keys, vals, vals2, start, end, off1_g, R ? ba_refs + sc.region0[g] : nullptr, count_tiles + g * count_words,
                                     flags_word(), sg, g == 0 ? c.pev[1] : nullptr, plan_stream != s ? c.msm_ev[g] : nullptr);
            if (rc) return rc;
            svals = vals2;
            if (g == 0) C12_CUDA(cudaEventRecord(c.pev[2], sg));
        } else {
            rc = launch_recode(lp, groups > 1 ? (int)g : -1, d_scalars, keys, vals, flags_word(), sg);
            if (rc) return rc;
            if (g == 0) C12_CUDA(cudaEventRecord(c.pev[1], sg));
            uint32_t *k1 = keys + seg_off, *v1 = vals + seg_off, *k2 = keys2 + seg_off, *v2 = vals2 + seg_off;
            rc = sort_pairs_segmented(k1, v1, k2, v2, lp.n, pl.windows, lp.c, hist + g * ghist_words, tiles + g * (gtile_words + 2), sg);
            if (rc) return rc;
            skeys = k1 - seg_off;
            svals = v1 - seg_off;
            if (g == 0) C12_CUDA(cudaEventRecord(c.pev[2], sg));
            rc = launch_bucket_bounds(lp, g * pl.windows, pl.windows, skeys, start, end, sg);
            if (rc) return rc;
            if (R) {
                rc = launch_ba_plan(B, start + (size_t)g * B, end + (size_t)g * B, 1, 0, 1, 1, off1_g, ba_tiles + g * plan_words, sg);
                if (rc) return rc;
            }
            if (plan_stream != s) C12_CUDA(cudaEventRecord(c.msm_ev[g], sg));      // this group's bounds and level-1 offsets are in place (by counting: recorded behind the scan)
        }
        if (R) {
            BaMapGeom mg;
            mg.start = start + (size_t)g * B;
            mg.end = end + (size_t)g * B;
            mg.vals = svals;
            mg.off_out[0] = off1_g;
            mg.off_in[0] = nullptr;
            mg.ref_region[0][0] = sc.region0[g];
            mg.level1.groups = 0;
            mg.refs = ba_refs;
            mg.total = B;
            mg.first_round = 0;
            mg.rounds = 1;
            mg.pipes = 1;
            mg.b_lo[0] = 0;
            mg.b_lo[1] = B;
            if (c.front_end != 0) rc = launch_ba_map(mg, sc.slots0, sg);      // by counting: the scatter wrote round 0's references itself
            if (rc) return rc;
        }
    }
    // ---- the merged levels: offsets of levels 2 .. R, slot maps of rounds 1 .. R - 1, the lists left for the accumulation ------
    uint32_t *order = nullptr, *vstart = nullptr, *vbucket = nullptr;
    const uint32_t *lstart = start, *lend = end;
    const Affine<F>* final_lists = nullptr;
    if (plan_stream != s)
        for (uint32_t g = 0; g < groups; ++g) C12_CUDA(cudaStreamWaitEvent(plan_stream, c.msm_ev[g], 0));
    if (R) {
        rc = launch_ba_plan(B, start, end, groups, B, 2, R - 1, offm, ba_tiles + groups * plan_words, plan_stream);
        if (rc) return rc;
        if (R > 1) {
            BaMapGeom mg;
            mg.start = mg.end = mg.vals = nullptr;
            for (uint32_t r = 1; r < R; ++r) {
                mg.off_out[r - 1] = offm + (size_t)(r - 1) * lvl_words;                       // level r + 1
                mg.off_in[r - 1] = r >= 2 ? offm + (size_t)(r - 2) * lvl_words : nullptr;     // level r
                for (uint32_t p = 0; p < sc.pipes; ++p) mg.ref_region[r - 1][p] = sc.ref_region[r][p];
            }
            for (uint32_t g = 0; g < groups; ++g) {
                mg.level1.off1[g] = off1 + g * lvl_words;
                mg.level1.region1[g] = sc.region0[g];
            }
            mg.level1.groups = groups;
            mg.refs = ba_refs;
            mg.total = B;
            mg.first_round = 1;
            mg.rounds = R - 1;
            mg.pipes = sc.pipes;
            for (uint32_t p = 0; p <= sc.pipes; ++p) mg.b_lo[p] = sc.b_lo[p];
            for (uint32_t p = 0; p < sc.pipes; ++p) {
                mg.list_region[0][p] = sc.region[0][p];
                mg.list_region[1][p] = sc.region[1][p];
            }
            rc = launch_ba_map(mg, sc.slots_merged, plan_stream);
            if (rc) return rc;
        }
        // what the rounds leave: level R - of the pipelines in buffer 2, or (one round only) of the single group in buffer 0
        BaListGeom lg;
        lg.lanes = R > 1 ? sc.pipes : 1;
        lg.vtotal = B;
        for (uint32_t p = 0; p < lg.lanes; ++p) {
            lg.off[p] = R > 1 ? offm + (size_t)(R - 2) * lvl_words : off1;
            lg.b_lo[p] = R > 1 ? sc.b_lo[p] : 0;
            lg.vb0[p] = R > 1 ? sc.b_lo[p] : 0;
            lg.final_region[p] = R > 1 ? sc.final_region[p] : sc.region0[0];
        }
        lg.vb0[lg.lanes] = B;
        final_lists = R > 1 ? E[2] : E[0];
        rc = launch_ba_list_bounds(lg, ba_lstart, ba_lend, plan_stream);
        if (rc) return rc;
        lstart = ba_lstart;
        lend = ba_lend;
    }
    // bucket lists longer than pl.chunk entries are cut into chunks (XYZZ additions, one thread per chunk)
    // split tail: each pipeline's buckets are ordered, accumulated, folded and reduced on their own (bucket ids relative to b_lo)
    const bool split_tail = sc.split_tail;
    MsmPlan tp[2] = {pl, pl};
    uint32_t *t_order[2] = {nullptr, nullptr}, *t_vstart[2] = {nullptr, nullptr}, *t_vbucket[2] = {nullptr, nullptr};
    if (split_tail) {
        for (uint32_t p = 0; p < 2; ++p) {
            tp[p].total = sc.b_lo[p + 1] - sc.b_lo[p];
            tp[p].vmax = tp[p].total + (pl.vmax - pl.total);
            rc = launch_chunk_order(tp[p], lstart + sc.b_lo[p], lend + sc.b_lo[p], p ? chunk_scratch2 : chunk_scratch, &t_vstart[p], &t_vbucket[p], &t_order[p],
                                    plan_stream);
            if (rc) return rc;
        }
    } else {
        rc = launch_chunk_order(pl, lstart, lend, chunk_scratch, &vstart, &vbucket, &order, plan_stream);
        if (rc) return rc;
    }
    if (plan_stream != s) C12_CUDA(cudaEventRecord(c.msm_ev[BA_MAX_PIPES], plan_stream));
    if (parse_aside) C12_CUDA(cudaEventRecord(c.parse_ev[1], c.front_stream));
    // ---- the points, group by group; round 0 of each group behind them ------------------------------------------------------
    for (uint32_t g = 0; g < groups; ++g) {
        cudaStream_t sg = lane_stream(g);
        if (points_ready) {
            if (groups > 1)
                C12_CUDA(cudaStreamWaitEvent(sg, points_ready[g], 0));
            else
                for (uint32_t q = 0; q < groups_req; ++q) C12_CUDA(cudaStreamWaitEvent(s, points_ready[q], 0));
        }
        if (g == 0 && !parse_aside) C12_CUDA(cudaEventRecord(c.pev[3], s));
        const uint32_t first = g * lp.n_group, last = first + lp.n_group < n ? first + lp.n_group : n;
        if (first < last) {
            k_parse_points<F><<<cdiv(last - first, 128), 128, 0, sg>>>(d_points, first, last, n, pl.parts, pts, flags_word());
            C12_LAUNCHED();
        }
        if (parse_aside) {          // join: the bucket lists.  Phase events: what is left of the parse behind the scatter counts as "bounds_order"
            C12_CUDA(cudaStreamWaitEvent(s, c.parse_ev[1], 0));
            C12_CUDA(cudaEventRecord(c.pev[3], s));
        }
        if (g == 0) {
            C12_CUDA(cudaEventRecord(c.ev[1], s));
            C12_CUDA(cudaEventRecord(c.pev[4], s));
        }
        if (R) {
            BaLaneArgs<F> a;
            a.stream = sg;
            a.first = true;
            a.warps = sc.warps0[g];
            a.g.off_out = off1 + g * lvl_words;
            a.g.refs = ba_refs + sc.region0[g];
            a.g.b_lo = 0;
            a.g.b_hi = B;
            a.g.J = sc.J0[g];
            a.g.out_region = sc.region0[g];
            a.g.scratch_region = sc.region0[g];
            a.io.pts = pts;
            a.io.lists = nullptr;
            a.io.out = E[0];
            a.io.prefix = ba_prefix;
            a.io.pool = ba_pool + (size_t)g * sc.pool_stride;
            a.io.others = ba_others + (size_t)g * sc.pool_stride * 32;
            rc = msm_ba_round_launch<F>(a);
            if (rc) return rc;
        }
    }
    if (groups > 1) {           // join: all level-1 lists and the merged plan
        for (uint32_t g = 1; g < groups; ++g) {
            C12_CUDA(cudaEventRecord(c.side_ev[g], c.side[g - 1]));
            C12_CUDA(cudaStreamWaitEvent(s, c.side_ev[g], 0));
        }
    }
    if (plan_stream != s) C12_CUDA(cudaStreamWaitEvent(s, c.msm_ev[BA_MAX_PIPES], 0));
    // ---- rounds 1 .. R - 1 on the merged lists, pipeline by pipeline ------------------------------------------------------------
    if (R > 1) {
        if (sc.pipes > 1) {
            C12_CUDA(cudaEventRecord(c.side_ev[0], s));
            for (uint32_t p = 1; p < sc.pipes; ++p) C12_CUDA(cudaStreamWaitEvent(c.side[p - 1], c.side_ev[0], 0));
        }
        for (uint32_t r = 1; r < R; ++r) {
            for (uint32_t p = 0; p < sc.pipes; ++p) {
                const bool last = r + 1 == R;
                BaLaneArgs<F> a;
                a.stream = lane_stream(p);
                a.first = false;
                a.warps = sc.warps[r][p];
                a.g.off_out = offm + (size_t)(r - 1) * lvl_words;
                a.g.refs = ba_refs + sc.ref_region[r][p];
                a.g.b_lo = sc.b_lo[p];
                a.g.b_hi = sc.b_lo[p + 1];
                a.g.J = sc.J[r][p];
                a.g.out_region = last ? sc.final_region[p] : sc.region[r & 1][p];
                a.g.scratch_region = sc.scratch_region[p];
                a.io.pts = nullptr;
                a.io.lists = r == 1 ? E[0] : (((r - 1) & 1) ? E[1] : E[3]);
                a.io.out = last ? E[2] : ((r & 1) ? E[1] : E[3]);
                a.io.prefix = ba_prefix;
                a.io.pool = ba_pool + (size_t)p * sc.pool_stride;
                a.io.others = ba_others + (size_t)p * sc.pool_stride * 32;
                rc = msm_ba_round_launch<F>(a);
                if (rc) return rc;
            }
        }
        if (!split_tail)
            for (uint32_t p = 1; p < sc.pipes; ++p) {
                C12_CUDA(cudaEventRecord(c.side_ev[p], c.side[p - 1]));
                C12_CUDA(cudaStreamWaitEvent(s, c.side_ev[p], 0));
            }
    }
    if (split_tail) {
        // Pipeline 1 (the high windows, the smaller share) finishes its rounds first: what is left of its lists, its bucket
        // reduction and its part of the Horner chain - INCLUDING the c * (windows below) doublings that carry it to its place -
        // run on its stream while pipeline 0 is still adding.  Pipeline 0 then has only its own windows' doublings to do.
        const uint32_t wsplit = sc.b_lo[1] / pl.half;
        for (uint32_t p = 0; p < 2; ++p) {
            cudaStream_t st = lane_stream(p);
            const uint32_t b0 = sc.b_lo[p], nb = tp[p].total, w0 = b0 / pl.half, nw = nb / pl.half;
            k_accumulate<F, true><<<cdiv(tp[p].vmax, AccShape<F>::THREADS), AccShape<F>::THREADS, 0, st>>>(tp[p].vmax, pl.chunk, lstart + b0, lend + b0, nullptr, final_lists,
                                                                                                           t_order[p], t_vbucket[p], t_vstart[p], p ? vpartial2 : vpartial);
            C12_LAUNCHED();
            k_fold<F><<<cdiv(nb, 128), 128, 0, st>>>(nb, 1, t_vstart[p], p ? vpartial2 : vpartial, buckets + b0, heavy[p]);
            C12_LAUNCHED();
            k_fold_heavy<F><<<2 * (unsigned)(c.sm_count > 0 ? c.sm_count : 148), 128, 0, st>>>(heavy[p], t_vstart[p], p ? vpartial2 : vpartial, buckets + b0);
            C12_LAUNCHED();
            if (p == 0) {
                C12_CUDA(cudaEventRecord(c.ev[2], s));
                C12_CUDA(cudaEventRecord(c.pev[5], s));
            }
            k_reduce_level0<F><<<dim3(cdiv(pl.segs, 128), nw), 128, 0, st>>>(pl, buckets, partial, w0);
            C12_LAUNCHED();
            if (p == 0) C12_CUDA(cudaEventRecord(c.pev[6], s));
            k_reduce_planes<F><<<dim3(pl.plane_bits + 1, nw, 2 * PlaneShape<F>::SPLITS), PlaneShape<F>::THREADS, 0, st>>>(pl, partial, plane_parts, plane_tickets, wsum, w0);
            C12_LAUNCHED();
            if (p == 0) C12_CUDA(cudaEventRecord(c.pev[7], s));
        }
        k_finish<F><<<1, 256, sizeof(Proj<F>) * (pl.windows - wsplit), c.side[0]>>>(pl, wsum, nullptr, out_mode, wsplit, pl.windows, wsplit, nullptr, hpart);
        C12_LAUNCHED();
        C12_CUDA(cudaEventRecord(c.side_ev[1], c.side[0]));
        C12_CUDA(cudaStreamWaitEvent(s, c.side_ev[1], 0));
        k_finish<F><<<1, 256, sizeof(Proj<F>) * wsplit, s>>>(pl, wsum, d_out, out_mode, 0, wsplit, 0, hpart, nullptr);
        C12_LAUNCHED();
        C12_CUDA(cudaEventRecord(c.pev[8], s));
        C12_CUDA(cudaEventRecord(c.ev[3], s));
        c.stats.window_bits = (int)pl.c;
        c.stats.ba_rounds = (int)sc.rounds;
        c.stats.ba_pipes = (int)sc.pipes;
        c.stats.groups = (int)groups;
        c.stats.bucket_adds = N;
        c.stats.accumulate_ms = -1.0;
        return C12381_OK;
    }
    if (R)
        k_accumulate<F, true><<<cdiv(pl.vmax, AccShape<F>::THREADS), AccShape<F>::THREADS, 0, s>>>(pl.vmax, pl.chunk, lstart, lend, nullptr, final_lists, order, vbucket, vstart, vpartial);
    else
        k_accumulate<F, false><<<cdiv(pl.vmax, AccShape<F>::THREADS), AccShape<F>::THREADS, 0, s>>>(pl.vmax, pl.chunk, start, end, svals, pts, order, vbucket, vstart, vpartial);
    C12_LAUNCHED();
    k_fold<F><<<cdiv(pl.total, 128), 128, 0, s>>>(pl.total, 1, vstart, vpartial, buckets, heavy[0]);
    C12_LAUNCHED();
    k_fold_heavy<F><<<2 * (unsigned)(c.sm_count > 0 ? c.sm_count : 148), 128, 0, s>>>(heavy[0], vstart, vpartial, buckets);
    C12_LAUNCHED();
    C12_CUDA(cudaEventRecord(c.ev[2], s));
    C12_CUDA(cudaEventRecord(c.pev[5], s));
    if (c.split_tail == 2 && pl.windows >= 4) {
        // LATE split (knob 7 = 2): one accumulation, then the reduction and the Horner part of the HIGH half of the windows as a chain
        // of its own on a side stream (which outranks `s`), beside the low half's.  The high chain is still c (W - 1) doublings
        // long; what it gains is that its planes are ready before all windows' would be.
        const uint32_t wsplit = pl.windows / 2, nhi = pl.windows - wsplit;
        C12_CUDA(cudaEventRecord(c.side_ev[0], s));
        C12_CUDA(cudaStreamWaitEvent(c.side[0], c.side_ev[0], 0));
        k_reduce_level0<F><<<dim3(cdiv(pl.segs, 128), nhi), 128, 0, c.side[0]>>>(pl, buckets, partial, wsplit);
        C12_LAUNCHED();
        k_reduce_planes<F><<<dim3(pl.plane_bits + 1, nhi, 2 * PlaneShape<F>::SPLITS), PlaneShape<F>::THREADS, 0, c.side[0]>>>(pl, partial, plane_parts, plane_tickets, wsum, wsplit);
        C12_LAUNCHED();
        k_finish<F><<<1, 256, sizeof(Proj<F>) * nhi, c.side[0]>>>(pl, wsum, nullptr, out_mode, wsplit, pl.windows, wsplit, nullptr, hpart);
        C12_LAUNCHED();
        C12_CUDA(cudaEventRecord(c.side_ev[1], c.side[0]));
        k_reduce_level0<F><<<dim3(cdiv(pl.segs, 128), wsplit), 128, 0, s>>>(pl, buckets, partial, 0);
        C12_LAUNCHED();
        C12_CUDA(cudaEventRecord(c.pev[6], s));
        k_reduce_planes<F><<<dim3(pl.plane_bits + 1, wsplit, 2 * PlaneShape<F>::SPLITS), PlaneShape<F>::THREADS, 0, s>>>(pl, partial, plane_parts, plane_tickets, wsum, 0);
        C12_LAUNCHED();
        C12_CUDA(cudaEventRecord(c.pev[7], s));
        k_finish<F><<<1, 256, sizeof(Proj<F>) * wsplit, s>>>(pl, wsum, nullptr, out_mode, 0, wsplit, 0, nullptr, hpart + 1);      // the low chain, beside the high one
        C12_LAUNCHED();
        C12_CUDA(cudaStreamWaitEvent(s, c.side_ev[1], 0));
        k_combine2<F><<<1, 32, 0, s>>>(hpart, hpart + 1, d_out, out_mode);
        C12_LAUNCHED();
    } else {
        k_reduce_level0<F><<<dim3(cdiv(pl.segs, 128), pl.windows), 128, 0, s>>>(pl, buckets, partial, 0);
        C12_LAUNCHED();
        C12_CUDA(cudaEventRecord(c.pev[6], s));
        k_reduce_planes<F><<<dim3(pl.plane_bits + 1, pl.windows, 2 * PlaneShape<F>::SPLITS), PlaneShape<F>::THREADS, 0, s>>>(pl, partial, plane_parts, plane_tickets, wsum, 0);
        C12_LAUNCHED();
        C12_CUDA(cudaEventRecord(c.pev[7], s));
        k_finish<F><<<1, 256, sizeof(Proj<F>) * pl.windows, s>>>(pl, wsum, d_out, out_mode, 0, pl.windows, 0, nullptr, nullptr);
        C12_LAUNCHED();
    }
    C12_CUDA(cudaEventRecord(c.pev[8], s));
    C12_CUDA(cudaEventRecord(c.ev[3], s));
    c.stats.window_bits = (int)pl.c;
    c.stats.ba_rounds = (int)sc.rounds;
    c.stats.ba_pipes = (int)sc.pipes;
    c.stats.groups = (int)groups;
    c.stats.bucket_adds = N;
    c.stats.accumulate_ms = -1.0;  // resolved lazily by c12381_last_msm_stats
    return C12381_OK;
}

template <class F> int sum_run(const uint8_t* d_points, size_t n, uint8_t* d_out, int out_mode, cudaStream_t s)
{
    Ctx& c = ctx();
    if (n > 0xffffffffull) return set_error(C12381_EARG, "sum: too many points");
    k_sum_points<F><<<1, 256, 0, s>>>(d_points, (uint32_t)n, d_out, out_mode, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}

template <class F>
int scalar_mul_run(const uint8_t* d_points, const uint8_t* d_scalars, size_t n, uint8_t* d_out, cudaStream_t s, int out_mode = OUT_COMPRESSED)
{
    Ctx& c = ctx();
    if (n == 0) return C12381_OK;
    if (n > 0x7fffffffull) return set_error(C12381_EARG, "mul_batch: too many terms");
    k_scalar_mul<F><<<cdiv(n, 128), 128, 0, s>>>(d_points, d_scalars, (uint32_t)n, d_out, out_mode, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}

// per-context device tables (freed by c12381_shutdown)
template <class F> Affine<F>*& fixed_base_table_slot();
template <> inline Affine<Fp>*& fixed_base_table_slot<Fp>() { return reinterpret_cast<Affine<Fp>*&>(ctx().fb_table[0]); }
template <> inline Affine<Fp2>*& fixed_base_table_slot<Fp2>() { return reinterpret_cast<Affine<Fp2>*&>(ctx().fb_table[1]); }

template <class F> int fixed_base_run(const uint8_t* d_scalars, size_t n, uint8_t* d_out, cudaStream_t s)
{
    Ctx& c = ctx();
    if (n == 0) return C12381_OK;
    if (n > 0x7fffffffull) return set_error(C12381_EARG, "fixed_base: too many terms");
    Affine<F>*& table = fixed_base_table_slot<F>();
    if (!table) {   // first use on this context: build the window table once; published only when the build has succeeded
        Affine<F>* fresh = nullptr;
        C12_CUDA(cudaMalloc(&fresh, sizeof(Affine<F>) * FB_WINDOWS * FB_HALF + sizeof(Proj<F>) * FB_WINDOWS));
        Proj<F>* wbase = reinterpret_cast<Proj<F>*>(fresh + FB_WINDOWS * FB_HALF);
        k_fixed_base_windows<F><<<1, 32, 0, s>>>(nullptr, 1, wbase, flags_word());
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) {
            ctx().launches++;
            k_fixed_base_table<F><<<cdiv(FB_WINDOWS * FB_HALF, 128), 128, 0, s>>>(1, wbase, fresh);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) {
            ctx().launches++;
            e = cudaStreamSynchronize(s);   // later calls may come on other streams
        }
        if (e != cudaSuccess) {
            cudaFree(fresh);
            return set_error(C12381_ECUDA, "fixed-base table build", e);
        }
        table = fresh;
    }
    k_fixed_base<F><<<cdiv(n, 128), 128, 0, s>>>(d_scalars, (uint32_t)n, 1, table, d_out, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}

// out[b] = sum_j scalars[b * m + j] * bases[j] for m caller-supplied bases shared by all B instances (window tables are
// built per call in the caller's arena reservation: multi_fixed_scratch<F>(m) bytes)
template <class F> size_t multi_fixed_scratch(size_t m)
{
    return align_up(sizeof(Affine<F>) * m * FB_WINDOWS * FB_HALF) + align_up(sizeof(Proj<F>) * m * FB_WINDOWS) + 4096;
}
template <class F> int multi_fixed_base_run(const uint8_t* d_bases, size_t m, const uint8_t* d_scalars, size_t B, uint8_t* d_out, cudaStream_t s)
{
    Ctx& c = ctx();
    if (B == 0) return C12381_OK;
    if (m == 0 || m > 4096) return set_error(C12381_EARG, "multi_fixed_base: between 1 and 4096 bases");
    if (B > 0x7fffffffull) return set_error(C12381_EARG, "multi_fixed_base: too many instances");
    Affine<F>* table = (Affine<F>*)arena_take(sizeof(Affine<F>) * m * FB_WINDOWS * FB_HALF);
    Proj<F>* wbase = (Proj<F>*)arena_take(sizeof(Proj<F>) * m * FB_WINDOWS);
    if (!wbase) return set_error(C12381_ECUDA, "multi_fixed_base: scratch arena bound too small");
    k_fixed_base_windows<F><<<(unsigned)m, 32, 0, s>>>(d_bases, (uint32_t)m, wbase, flags_word());
    C12_LAUNCHED();
    k_fixed_base_table<F><<<cdiv(m * FB_WINDOWS * FB_HALF, 128), 128, 0, s>>>((uint32_t)m, wbase, table);
    C12_LAUNCHED();
    k_fixed_base<F><<<cdiv(B, 128), 128, 0, s>>>(d_scalars, (uint32_t)B, (uint32_t)m, table, d_out, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}

// verdict[i] = 1 iff points[i] is in the r-torsion subgroup (the identity is not: the reference's convention)
template <class F>
__global__ void __launch_bounds__(128) k_subgroup_check(const uint8_t* __restrict__ in, uint32_t n, uint8_t* __restrict__ out, int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p;
    if (!Wire<F>::parse(p, in + (size_t)Wire<F>::AFFINE * i)) atomicOr(flags, FLAG_BAD_POINT);
    out[i] = subgroup_member(p) ? 1 : 0;
}

template <class F> int subgroup_run(const uint8_t* d_in, size_t n, uint8_t* d_out, cudaStream_t s)
{
    Ctx& c = ctx();
    if (n == 0) return C12381_OK;
    if (n > 0x7fffffffull) return set_error(C12381_EARG, "subgroup_check: too many points");
    k_subgroup_check<F><<<cdiv(n, 128), 128, 0, s>>>(d_in, (uint32_t)n, d_out, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}

template <class F> int convert_run(const uint8_t* d_in, size_t n, uint8_t* d_out, bool decompress, cudaStream_t s)
{
    Ctx& c = ctx();
    if (n == 0) return C12381_OK;
    if (n > 0x7fffffffull) return set_error(C12381_EARG, "convert: too many points");
    if (decompress)
        k_decompress<F><<<cdiv(n, 128), 128, 0, s>>>(d_in, (uint32_t)n, d_out, flags_word());
    else
        k_compress<F><<<cdiv(n, 128), 128, 0, s>>>(d_in, (uint32_t)n, d_out, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}

// ---- host-pointer wrappers: stage in, run, stage out, report the flag word ---------------------------------------
int flags_reset(cudaStream_t s);
int flags_collect(cudaStream_t s);  // synchronises `s`; returns C12381_EINPUT if any input was malformed

// Host-pointer entry: one arena reservation covers staging + pipeline scratch; copies ride the context stream.
// run(d_in[], d_out, stream) enqueues the device pipeline.
struct HostScope {      // a host-pointer entry is running: its kernels flag malformed input in the host word (common.cuh)
    HostScope() { ctx().host_depth++; }
    ~HostScope() { ctx().host_depth--; }
};

template <class Fn>
int with_staged(const void* const* host_in, const size_t* in_bytes, int n_in, void* host_out, size_t out_bytes, size_t scratch, Fn&& run)
{
    Ctx& c = ctx();
    HostScope scope;
    cudaStream_t s = c.stream;
    size_t total = scratch + align_up(out_bytes) + 4096;
    for (int i = 0; i < n_in; ++i) total += align_up(in_bytes[i]);
    int rc = arena_begin(total, s);
    if (rc) return rc;
    uint8_t* d_in[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < n_in; ++i) {
        d_in[i] = (uint8_t*)arena_take(in_bytes[i] ? in_bytes[i] : 4);
        if (in_bytes[i]) C12_CUDA(cudaMemcpyAsync(d_in[i], host_in[i], in_bytes[i], cudaMemcpyHostToDevice, s));
    }
    uint8_t* d_out = (uint8_t*)arena_take(out_bytes ? out_bytes : 4);
    rc = flags_reset(s);
    if (rc) return rc;
    c.arena_depth++;
    rc = run(d_in, d_out, s);
    c.arena_depth--;
    if (rc) {
        cudaStreamSynchronize(s);
        return rc;
    }
    if (out_bytes) C12_CUDA(cudaMemcpyAsync(host_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    return flags_collect(s);
}

} // namespace c12

// ---- entry-point bodies shared by the G1 and G2 translation units --------------------------------------------
namespace c12 {

// `_dev` entries run on exactly the stream they are given; NULL is CUDA's default stream (what torch uses unless
// told otherwise), NOT the context's private stream — the caller's other work on that stream stays ordered with ours.
static inline cudaStream_t pick_stream(void* stream) { return (cudaStream_t)stream; }

template <class F> int entry_msm_dev(const uint8_t* d_points, const uint8_t* d_scalars, size_t n, uint8_t* d_out, int out_mode, void* stream)
{
    C12_REQUIRE_CTX();
    if (!d_out || (n && (!d_points || !d_scalars))) return set_error(C12381_EARG, "msm: null pointer");
    cudaStream_t s = pick_stream(stream);
    int rc = arena_begin(msm_scratch_for<F>(n), s);
    if (rc) return rc;
    return msm_run<F>(d_points, d_scalars, n, d_out, out_mode, s);
}

template <class F> int entry_msm_host(const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out, int out_mode = OUT_COMPRESSED)
{
    C12_REQUIRE_CTX();
    if (!out || (n && (!points || !scalars))) return set_error(C12381_EARG, "msm: null pointer");
    HostScope scope;
    const size_t out_bytes = out_mode == OUT_AFFINE ? Wire<F>::AFFINE : Wire<F>::COMPRESSED;
    // scalars go up on the context stream (the first stages need only them); the points - three quarters of the bytes -
    // follow on the copy stream in upload groups, each awaited by the pipeline that adds up that group's bucket lists
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    const size_t pb = n * Wire<F>::AFFINE, sb = n * 32;
    uint32_t groups = (uint32_t)c.upload_groups;
    {
        MsmPlan pl, lp;
        if (!msm_plans<F>(n, groups, pl, lp)) groups = 1; else groups = lp.groups;
    }
    int rc = arena_begin(msm_scratch_for<F>(n, groups) + align_up(pb) + align_up(sb) + 8192, s);
    if (rc) return rc;
    uint8_t* d_pts = (uint8_t*)arena_take(pb ? pb : 4);
    uint8_t* d_sc = (uint8_t*)arena_take(sb ? sb : 4);
    uint8_t* d_out = (uint8_t*)arena_take(out_bytes);
    rc = flags_reset(s);
    if (rc) return rc;
    cudaEvent_t ready_s[BA_MAX_PIPES] = {}, ready_p[BA_MAX_PIPES] = {};
    if (n) {
        // copy stream, behind the point of `s` from which the arena is ours: scalars and points of group 0, of group 1, ...
        C12_CUDA(cudaEventRecord(c.copy_ev[0], s));
        C12_CUDA(cudaStreamWaitEvent(c.copy_stream, c.copy_ev[0], 0));
        const size_t per = (n + groups - 1) / groups;
        for (uint32_t g = 0; g < groups; ++g) {
            const size_t first = (size_t)g * per, last = first + per < n ? first + per : n;
            if (first < last) C12_CUDA(cudaMemcpyAsync(d_sc + first * 32, scalars + first * 32, (last - first) * 32, cudaMemcpyHostToDevice, c.copy_stream));
            C12_CUDA(cudaEventRecord(c.sgroup_ev[g], c.copy_stream));
            ready_s[g] = c.sgroup_ev[g];
            if (first < last)
                C12_CUDA(cudaMemcpyAsync(d_pts + first * Wire<F>::AFFINE, points + first * Wire<F>::AFFINE, (last - first) * Wire<F>::AFFINE,
                                         cudaMemcpyHostToDevice, c.copy_stream));
            C12_CUDA(cudaEventRecord(c.group_ev[g], c.copy_stream));
            ready_p[g] = c.group_ev[g];
        }
    }
    rc = msm_run<F>(d_pts, d_sc, n, d_out, out_mode, s, groups, n ? ready_s : nullptr, n ? ready_p : nullptr);
    if (rc) {
        cudaStreamSynchronize(c.copy_stream);
        cudaStreamSynchronize(s);
        return rc;
    }
    C12_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    return flags_collect(s);
}

template <class F> int entry_sum_dev(const uint8_t* d_points, size_t n, uint8_t* d_out, void* stream)
{
    C12_REQUIRE_CTX();
    if (!d_out || (n && !d_points)) return set_error(C12381_EARG, "sum: null pointer");
    return sum_run<F>(d_points, n, d_out, OUT_COMPRESSED, pick_stream(stream));
}

// A batch of ONE (the reference's multiply(point&, big) called term by term) is a one-term sum of products: the MSM pipeline's
// tail runs on lane-cooperative point arithmetic, a single thread's double-and-add does not (G1 2.48 -> 0.86 ms, G2 4.58 ->
// 1.4 ms per call, profiles/r03c_latency_probe.txt).  Same bytes: both leave as the canonical compressed encoding.
template <class F> int entry_mul_dev(const uint8_t* d_points, const uint8_t* d_scalars, size_t n, uint8_t* d_out, void* stream)
{
    C12_REQUIRE_CTX();
    if (n && (!d_out || !d_points || !d_scalars)) return set_error(C12381_EARG, "mul_batch: null pointer");
    if (n == 1) {
        cudaStream_t s = pick_stream(stream);
        int rc = arena_begin(msm_scratch_for<F>(1), s);
        if (rc) return rc;
        return msm_run<F>(d_points, d_scalars, 1, d_out, OUT_COMPRESSED, s);
    }
    return scalar_mul_run<F>(d_points, d_scalars, n, d_out, pick_stream(stream));
}

template <class F> int entry_mul_host(const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out)
{
    C12_REQUIRE_CTX();
    if (n && (!out || !points || !scalars)) return set_error(C12381_EARG, "mul_batch: null pointer");
    const void* in[2] = {points, scalars};
    size_t sz[2] = {n * Wire<F>::AFFINE, n * 32};
    return with_staged(in, sz, 2, out, n * Wire<F>::COMPRESSED, n == 1 ? msm_scratch_for<F>(1) : 0, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        if (n == 1) return msm_run<F>(d_in[0], d_in[1], 1, d_out, OUT_COMPRESSED, s);      // see entry_mul_dev
        return scalar_mul_run<F>(d_in[0], d_in[1], n, d_out, s);
    });
}

template <class F> int entry_fixed_dev(const uint8_t* d_scalars, size_t n, uint8_t* d_out, void* stream)
{
    C12_REQUIRE_CTX();
    if (n && (!d_out || !d_scalars)) return set_error(C12381_EARG, "fixed_base: null pointer");
    return fixed_base_run<F>(d_scalars, n, d_out, pick_stream(stream));
}

template <class F> int entry_fixed_host(const uint8_t* scalars, size_t n, uint8_t* out)
{
    C12_REQUIRE_CTX();
    if (n && (!out || !scalars)) return set_error(C12381_EARG, "fixed_base: null pointer");
    const void* in[1] = {scalars};
    size_t sz[1] = {n * 32};
    return with_staged(in, sz, 1, out, n * Wire<F>::AFFINE, 0, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        return fixed_base_run<F>(d_in[0], n, d_out, s);
    });
}

template <class F> int entry_multi_fixed_dev(const uint8_t* d_bases, size_t m, const uint8_t* d_scalars, size_t B, uint8_t* d_out, void* stream)
{
    C12_REQUIRE_CTX();
    if (B && (!d_out || !d_scalars || !d_bases)) return set_error(C12381_EARG, "multi_fixed_base: null pointer");
    cudaStream_t s = pick_stream(stream);
    int rc = arena_begin(multi_fixed_scratch<F>(m), s);
    if (rc) return rc;
    return multi_fixed_base_run<F>(d_bases, m, d_scalars, B, d_out, s);
}

template <class F> int entry_multi_fixed_host(const uint8_t* bases, size_t m, const uint8_t* scalars, size_t B, uint8_t* out)
{
    C12_REQUIRE_CTX();
    if (B && (!out || !scalars || !bases)) return set_error(C12381_EARG, "multi_fixed_base: null pointer");
    const void* in[2] = {bases, scalars};
    size_t sz[2] = {m * Wire<F>::AFFINE, B * m * 32};
    return with_staged(in, sz, 2, out, B * Wire<F>::AFFINE, multi_fixed_scratch<F>(m), [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        return multi_fixed_base_run<F>(d_in[0], m, d_in[1], B, d_out, s);
    });
}

template <class F> int entry_convert_dev(const uint8_t* d_in, size_t n, uint8_t* d_out, bool decompress, void* stream)
{
    C12_REQUIRE_CTX();
    if (n && (!d_in || !d_out)) return set_error(C12381_EARG, "convert: null pointer");
    return convert_run<F>(d_in, n, d_out, decompress, pick_stream(stream));
}

template <class F> int entry_convert_host(const uint8_t* in_bytes, size_t n, uint8_t* out, bool decompress)
{
    C12_REQUIRE_CTX();
    if (n && (!in_bytes || !out)) return set_error(C12381_EARG, "convert: null pointer");
    const void* in[1] = {in_bytes};
    size_t sz[1] = {n * (decompress ? Wire<F>::COMPRESSED : Wire<F>::AFFINE)};
    return with_staged(in, sz, 1, out, n * (decompress ? Wire<F>::AFFINE : Wire<F>::COMPRESSED), 0, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        return convert_run<F>(d_in[0], n, d_out, decompress, s);
    });
}

template <class F> int entry_subgroup_dev(const uint8_t* d_in, size_t n, uint8_t* d_out, void* stream)
{
    C12_REQUIRE_CTX();
    if (n && (!d_in || !d_out)) return set_error(C12381_EARG, "subgroup_check: null pointer");
    return subgroup_run<F>(d_in, n, d_out, pick_stream(stream));
}

template <class F> int entry_subgroup_host(const uint8_t* in_bytes, size_t n, uint8_t* out)
{
    C12_REQUIRE_CTX();
    if (n && (!in_bytes || !out)) return set_error(C12381_EARG, "subgroup_check: null pointer");
    const void* in[1] = {in_bytes};
    size_t sz[1] = {n * Wire<F>::AFFINE};
    return with_staged(in, sz, 1, out, n, 0, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) { return subgroup_run<F>(d_in[0], n, d_out, s); });
}

} // namespace c12
