// Field-independent parts of the MSM pipeline: scalar recoding, the segmented radix sort driver, bucket bounds,
// and the input-validation flag word.  (Kernels here are launched only through the host functions below, so the
// per-field translation units never reference a __global__ symbol of another TU.)
#include "msm_impl.cuh"
#include "sort.cuh"

namespace c12 {

__global__ void __launch_bounds__(128) k_recode(MsmPlan pl, const uint8_t* __restrict__ scalars, uint32_t* __restrict__ keys,
                                                uint32_t* __restrict__ vals, int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.n_in) return;
    if (!scalar_is_canonical(scalar_from_be32(scalars + 32ull * i))) atomicOr(flags, FLAG_BAD_SCALAR);
    msm_recode_body(pl, i, scalars, keys, vals);
}

int launch_recode(const MsmPlan& pl, const uint8_t* d_scalars, uint32_t* keys, uint32_t* vals, int* flags, cudaStream_t s)
{
    k_recode<<<cdiv(pl.n_in, 128), 128, 0, s>>>(pl, d_scalars, keys, vals, flags);
    C12_LAUNCHED();
    return C12381_OK;
}

int launch_bucket_bounds(const MsmPlan& pl, const uint32_t* keys, uint32_t* start, uint32_t* end, cudaStream_t s)
{
    C12_CUDA(cudaMemsetAsync(start, 0, 4 * (size_t)pl.total, s));
    C12_CUDA(cudaMemsetAsync(end, 0, 4 * (size_t)pl.total, s));
    k_bucket_bounds<<<dim3(cdiv(pl.n, 256), pl.windows), 256, 0, s>>>(keys, pl.n, pl.half, start, end);
    C12_LAUNCHED();
    return C12381_OK;
}

__global__ void k_bucket_size_keys(uint32_t total, const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ ids)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total) return;
    uint32_t sz = end[b] - start[b];
    keys[b] = 255u - (sz > 255u ? 255u : sz);   // ascending key = descending size
    ids[b] = b;
}

__global__ void k_ba_counts(uint32_t total, const uint32_t* __restrict__ start, const uint32_t* __restrict__ end, uint32_t* __restrict__ out)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < total) out[b] = ba_pairs0(end[b] - start[b]);
}

size_t ba_offsets_tile_words(const MsmPlan& pl) { return cdiv(pl.total, SCAN_TILE) + 2; }

// o0[b] = sum of the round-0 pair counts of the buckets before b
int launch_ba_offsets(const MsmPlan& pl, const uint32_t* start, const uint32_t* end, uint32_t* o0, uint32_t* tile_sums, cudaStream_t s)
{
    k_ba_counts<<<cdiv(pl.total, 256), 256, 0, s>>>(pl.total, start, end, o0);
    C12_LAUNCHED();
    const uint32_t ntiles = cdiv(pl.total, SCAN_TILE);
    k_scan_tile_sums<<<ntiles, SCAN_THREADS, 0, s>>>(o0, pl.total, tile_sums);
    C12_LAUNCHED();
    k_scan_top<<<1, SCAN_THREADS, 0, s>>>(tile_sums, ntiles);
    C12_LAUNCHED();
    k_scan_apply<<<ntiles, SCAN_THREADS, 0, s>>>(o0, pl.total, tile_sums);
    C12_LAUNCHED();
    return C12381_OK;
}

size_t bucket_order_scratch_words(const MsmPlan& pl)
{
    size_t tile_words = 0;
    size_t hist = sort_scratch_words(pl.total, 1, &tile_words);
    return 4 * (size_t)pl.total + hist + tile_words + 64;
}

int launch_bucket_order(const MsmPlan& pl, const uint32_t* start, const uint32_t* end, uint32_t* scratch, uint32_t** order, cudaStream_t s)
{
    size_t tile_words = 0;
    size_t hist_words = sort_scratch_words(pl.total, 1, &tile_words);
    uint32_t* keys = scratch;
    uint32_t* ids = keys + pl.total;
    uint32_t* keys2 = ids + pl.total;
    uint32_t* ids2 = keys2 + pl.total;
    uint32_t* hist = ids2 + pl.total;
    uint32_t* tiles = hist + hist_words;
    k_bucket_size_keys<<<cdiv(pl.total, 256), 256, 0, s>>>(pl.total, start, end, keys, ids);
    C12_LAUNCHED();
    int rc = sort_pairs_segmented(keys, ids, keys2, ids2, pl.total, 1, 8, hist, tiles, s);
    if (rc) return rc;
    *order = ids;
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
