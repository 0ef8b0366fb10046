// Deterministic LSD radix sort of (key, value) u32 pairs, 8 bits per pass, stable.
// Used to group the MSM's (bucket, term) pairs by bucket without atomically accumulating curve points:
// the per-bucket lists come out in term order, so the whole MSM is reproducible run to run.
// (The reference adds terms in input order on one CPU thread, ecp_BLS12381.cpp:1131-1137; there is nothing
// to port — this is the B200 replacement for that loop's bucket indexing.)
//
// The sort is SEGMENTED: the recode kernel writes window w's pairs at [w*n, (w+1)*n) and keys are window-local
// bucket indices, so each window is sorted independently (blockIdx.y = segment) and the window bits never
// have to be sorted at all.  The histogram matrix is laid out [segment][digit][tile], so ONE exclusive scan
// over the whole matrix yields global output positions.
//
// Per pass: k_radix_hist (per-tile digit histogram, digit-major matrix) -> exclusive scan of the matrix ->
// k_radix_scatter (re-reads the tile, ranks every item stably with warp match/ballot + per-warp counters).
// Shared-memory integer atomics are used only for COUNTING (order-independent); ranks never depend on them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace c12 {

constexpr int SORT_THREADS = 256;
#ifndef C12_SORT_ITEMS
#define C12_SORT_ITEMS 8
#endif
constexpr int SORT_ITEMS = C12_SORT_ITEMS;   // items per thread (tile = 256 x this).  Measured at n = 2^20: 16 -> 0.55 ms, 8 -> 0.43, 6 -> 0.42, 4 -> 0.47 (profiles/r01ay)
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
constexpr int SORT_WARPS = SORT_THREADS / 32;

__global__ void __launch_bounds__(SORT_THREADS) k_radix_hist(const uint32_t* __restrict__ keys, uint32_t n, int shift,
                                                               uint32_t* __restrict__ hist, uint32_t nblk)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t* seg_keys = keys + (uint64_t)blockIdx.y * n;
    uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        uint64_t idx = base + (uint64_t)i * SORT_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&sh[(seg_keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[((uint64_t)blockIdx.y * 256 + threadIdx.x) * nblk + blockIdx.x] = sh[threadIdx.x];
}

// ---- exclusive scan over m u32 values (3 kernels) ----------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* total)
{
    __shared__ uint32_t warp_sums[8];
    __shared__ uint32_t tot;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 8; ++i) {
            uint32_t t = warp_sums[i];
            warp_sums[i] = run;
            run += t;
        }
        tot = run;
    }
    __syncthreads();
    uint32_t r = x - v + warp_sums[w];
    *total = tot;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t* __restrict__ data, uint64_t m,
                                                                   uint32_t* __restrict__ tile_sums)
{
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < m) s += data[base + i];
    uint32_t total;
    block_exclusive_scan_256(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_top(uint32_t* __restrict__ tile_sums, uint32_t ntiles)
{
    uint32_t carry = 0;
    for (uint32_t base = 0; base < ntiles; base += SCAN_THREADS) {
        uint32_t idx = base + threadIdx.x;
        uint32_t v = idx < ntiles ? tile_sums[idx] : 0;
        uint32_t total;
        uint32_t e = block_exclusive_scan_256(v, &total);
        if (idx < ntiles) tile_sums[idx] = e + carry;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(uint32_t* __restrict__ data, uint64_t m,
                                                               const uint32_t* __restrict__ tile_sums)
{
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < m) ? data[base + i] : 0;
        s += v[i];
    }
    uint32_t total;
    uint32_t e = block_exclusive_scan_256(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < m) data[base + i] = e;
        e += v[i];
    }
}

// ---- stable scatter ------------------------------------------------------------------------------------------
// Ranks every item of the tile stably (warp match/ballot + per-warp counters), then STAGES the tile in shared memory in
// digit order before writing: a tile of 2,048 items over 256 digits holds ~8 items per digit, so consecutive threads
// write runs of ~32 contiguous bytes (one sector) instead of isolated 4-byte words (the direct scatter moved 8x the sectors).
// Smaller tiles rank faster (more blocks in flight per SM) than they lose in run length: see SORT_ITEMS.
__global__ void __launch_bounds__(SORT_THREADS) k_radix_scatter(const uint32_t* __restrict__ keys_in,
                                                                  const uint32_t* __restrict__ vals_in,
                                                                  uint32_t* __restrict__ keys_out,
                                                                  uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                                  const uint32_t* __restrict__ hist_scanned, uint32_t nblk)
{
    __shared__ uint32_t cnt[SORT_WARPS][256];     // per-warp digit counts, then per-warp tile-local offsets
    __shared__ uint32_t tile_off[257];            // tile-local start of each digit's run
    __shared__ uint32_t gbase[256];               // global position of each digit's run of this tile
    __shared__ uint32_t skey[SORT_TILE], sval[SORT_TILE];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    // warp w owns the contiguous sub-tile [base, base + 32*ITEMS); item (i, lane) sits at base + 32 i + lane
    keys_in += (uint64_t)blockIdx.y * n;
    vals_in += (uint64_t)blockIdx.y * n;
    const uint64_t base = (uint64_t)blockIdx.x * SORT_TILE + (uint64_t)w * (32 * SORT_ITEMS);
    uint32_t k[SORT_ITEMS], v[SORT_ITEMS], rank[SORT_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        uint64_t idx = base + (uint64_t)i * 32 + lane;
        bool act = idx < n;
        k[i] = act ? keys_in[idx] : 0xffffffffu;
        v[i] = act ? vals_in[idx] : 0u;
        uint32_t d = act ? ((k[i] >> shift) & 255u) : 256u;  // 256: inactive lanes group together, never counted
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t before = act ? cnt[w][d] : 0u;
        rank[i] = before + __popc(peers & lt);
        __syncwarp();
        if (act && (peers & lt) == 0) cnt[w][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {
        // thread d: digit d's count in this tile, per-warp exclusive offsets inside the digit's run
        const uint32_t d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < SORT_WARPS; ++ww) {
            uint32_t c = cnt[ww][d];
            cnt[ww][d] = run;
            run += c;
        }
        gbase[d] = hist_scanned[((uint64_t)blockIdx.y * 256 + d) * nblk + blockIdx.x];
        // exclusive scan of the 256 digit counts -> tile-local run starts
        uint32_t total;
        uint32_t e = block_exclusive_scan_256(run, &total);
        tile_off[d] = e;
        if (d == 255) tile_off[256] = total;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        uint64_t idx = base + (uint64_t)i * 32 + lane;
        if (idx < n) {
            uint32_t d = (k[i] >> shift) & 255u;
            uint32_t slot = tile_off[d] + cnt[w][d] + rank[i];
            skey[slot] = k[i];
            sval[slot] = v[i];
        }
    }
    __syncthreads();
    const uint32_t count = tile_off[256];
    for (uint32_t slot = threadIdx.x; slot < count; slot += SORT_THREADS) {
        uint32_t kk = skey[slot];
        uint32_t d = (kk >> shift) & 255u;
        uint32_t pos = gbase[d] + (slot - tile_off[d]);
        keys_out[pos] = kk;
        vals_out[pos] = sval[slot];
    }
}

// bucket boundaries in the sorted key array of segment (window) blockIdx.y: start[w*half + k] = first GLOBAL index
// with key k, end[...] = one past the last; keys >= half (zero digits) own no bucket.  start/end are pre-zeroed.
__global__ void k_bucket_bounds(const uint32_t* __restrict__ keys, uint32_t n, uint32_t half, uint32_t* __restrict__ start,
                                uint32_t* __restrict__ end, uint32_t seg0)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t off = (uint64_t)(seg0 + blockIdx.y) * n;       // segments seg0 .. seg0 + gridDim.y - 1 of the whole array
    const uint32_t* seg_keys = keys + off;
    uint32_t k = seg_keys[i];
    if (k >= half) return;
    uint64_t b = (uint64_t)(seg0 + blockIdx.y) * half + k;
    if (i == 0 || seg_keys[i - 1] != k) start[b] = (uint32_t)(off + i);
    if (i + 1 == n || seg_keys[i + 1] != k) end[b] = (uint32_t)(off + i + 1);
}

} // namespace c12
