// Context management, error reporting and the integer-multiply roofline probes of libc12381_cuda.so.
#include "common.cuh"
#include "fp2.cuh"

namespace c12 {

static Ctx g_ctx;
Ctx& ctx() { return g_ctx; }

int set_error(int code, const char* what, cudaError_t e)
{
    char buf[512];
    if (e != cudaSuccess)
        snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    else
        snprintf(buf, sizeof buf, "%s", what);
    g_ctx.err = buf;
    return code;
}

int arena_begin(size_t total, cudaStream_t s)
{
    Ctx& c = g_ctx;
    total = align_up(total + 4096);
    if (c.arena_depth > 0) {   // an entry point called from inside a host entry: its scratch was reserved by the outer call
        if (c.arena_used + total > c.arena_bytes) return set_error(C12381_ECUDA, "nested scratch reservation too small");
        return C12381_OK;
    }
    // the previous call carved from the same arena on another stream: order this call's stream behind everything enqueued there
    bool ordered = !c.arena_stream_valid || c.arena_stream == s;
    if (!ordered) {
        if (cudaEventRecord(c.arena_ev, c.arena_stream) == cudaSuccess) {
            C12_CUDA(cudaStreamWaitEvent(s, c.arena_ev, 0));
            ordered = true;
        } else {
            cudaGetLastError();            // that stream no longer exists: whatever ran on it has been synchronised by its destruction
            c.arena_stream_valid = false;
        }
    }
    if (total > c.arena_bytes) {
        // growing: earlier work may still read the old arena
        if (c.arena_stream_valid && c.arena_stream != s && ordered) C12_CUDA(cudaEventSynchronize(c.arena_ev));
        C12_CUDA(cudaStreamSynchronize(s));
        if (s != c.stream) C12_CUDA(cudaStreamSynchronize(c.stream));
        if (c.copy_stream) C12_CUDA(cudaStreamSynchronize(c.copy_stream));
        if (c.arena) C12_CUDA(cudaFree(c.arena));
        c.arena = nullptr;
        c.arena_bytes = 0;
        size_t want = total + total / 8;
        C12_CUDA(cudaMalloc(&c.arena, want));
        c.arena_bytes = want;
    }
    c.arena_used = 0;
    c.arena_stream = s;
    c.arena_stream_valid = true;
    return C12381_OK;
}

void* arena_take(size_t bytes)
{
    Ctx& c = g_ctx;
    size_t off = c.arena_used;
    c.arena_used = align_up(off + bytes);
    if (c.arena_used > c.arena_bytes) return nullptr;  // arena_begin bound was wrong: caller checks
    return c.arena + off;
}

// ---- probes ----------------------------------------------------------------------------------------------
constexpr int PROBE_THREADS = 256;

// kind 0: 8 independent mad.lo.u32 chains
__global__ void __launch_bounds__(PROBE_THREADS) k_probe_imad(uint32_t* out, int iters, uint32_t seed)
{
    uint32_t a0 = threadIdx.x + seed, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    uint32_t m = seed | 1u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            asm volatile("mad.lo.u32 %0, %0, %8, %1;\n\t"
                         "mad.lo.u32 %1, %1, %8, %2;\n\t"
                         "mad.lo.u32 %2, %2, %8, %3;\n\t"
                         "mad.lo.u32 %3, %3, %8, %4;\n\t"
                         "mad.lo.u32 %4, %4, %8, %5;\n\t"
                         "mad.lo.u32 %5, %5, %8, %6;\n\t"
                         "mad.lo.u32 %6, %6, %8, %7;\n\t"
                         "mad.lo.u32 %7, %7, %8, %0;\n\t"
                         : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7)
                         : "r"(m));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

// kind 1: four independent carry chains of (mad.lo.cc, madc.hi.cc) pairs, the shape of a Montgomery row
__global__ void __launch_bounds__(PROBE_THREADS) k_probe_madc(uint32_t* out, int iters, uint32_t seed)
{
    uint32_t a0 = threadIdx.x + seed, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    uint32_t m = seed | 1u, q = seed * 3u + 5u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\t"
                         "madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                         "madc.lo.cc.u32 %2, %8, %9, %2;\n\t"
                         "madc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                         "madc.lo.cc.u32 %4, %8, %9, %4;\n\t"
                         "madc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                         "madc.lo.cc.u32 %6, %8, %9, %6;\n\t"
                         "madc.hi.u32 %7, %8, %9, %7;\n\t"
                         : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7)
                         : "r"(m), "r"(q));
            m += a7;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

// kind 2: 8 independent mad.wide.u32 chains (32x32 -> 64 + 64).  Every chain multiplies the low word of ITS OWN accumulator: with
// one shared pair of multiplicands (the probe until r03u) ptxas computes the product once per step and the eight "mad.wide" turn into
// 64-bit additions on the ALU pipe - ncu showed fmaheavy 27 %, ALU 76 % and only 41 IMAD.WIDE per 320 in the SASS
// (profiles/r03u_k_probe_wide_full.md): that "13.6 T/s" was not a multiplier rate.
__global__ void __launch_bounds__(PROBE_THREADS) k_probe_wide(uint32_t* out, int iters, uint32_t seed)
{
    unsigned long long a0 = threadIdx.x + seed, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    uint32_t q = seed * 3u + 5u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            asm volatile("{\n\t"
                         ".reg .u32 t0, t1, t2, t3, t4, t5, t6, t7;\n\t"
                         "cvt.u32.u64 t0, %0;\n\t"
                         "cvt.u32.u64 t1, %1;\n\t"
                         "cvt.u32.u64 t2, %2;\n\t"
                         "cvt.u32.u64 t3, %3;\n\t"
                         "cvt.u32.u64 t4, %4;\n\t"
                         "cvt.u32.u64 t5, %5;\n\t"
                         "cvt.u32.u64 t6, %6;\n\t"
                         "cvt.u32.u64 t7, %7;\n\t"
                         "mad.wide.u32 %0, t0, %8, %0;\n\t"
                         "mad.wide.u32 %1, t1, %8, %1;\n\t"
                         "mad.wide.u32 %2, t2, %8, %2;\n\t"
                         "mad.wide.u32 %3, t3, %8, %3;\n\t"
                         "mad.wide.u32 %4, t4, %8, %4;\n\t"
                         "mad.wide.u32 %5, t5, %8, %5;\n\t"
                         "mad.wide.u32 %6, t6, %8, %6;\n\t"
                         "mad.wide.u32 %7, t7, %8, %7;\n\t"
                         "}\n\t"
                         : "+l"(a0), "+l"(a1), "+l"(a2), "+l"(a3), "+l"(a4), "+l"(a5), "+l"(a6), "+l"(a7)
                         : "r"(q));
        }
    }
    unsigned long long x = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)x ^ (uint32_t)(x >> 32);
}

// kind 3 / 4: chains of full Montgomery products / squarings (two independent chains per thread)
__global__ void __launch_bounds__(PROBE_THREADS) k_probe_fp(uint32_t* out, int iters, uint32_t seed, int square)
{
    Fp a = fp_one(), b = fp_r2();
    a.v[0] ^= (threadIdx.x + seed) & 0xffu;
    b.v[1] ^= (blockIdx.x + seed) & 0xffu;
    Fp c = fp_add(a, b), d = fp_sub(a, b);
    if (square) {
        for (int i = 0; i < iters; ++i) {
            c = fp_sqr_inl(c);
            d = fp_sqr_inl(d);
        }
    } else {
        for (int i = 0; i < iters; ++i) {
            c = fp_mul_inl(c, a);
            d = fp_mul_inl(d, b);
        }
    }
    uint32_t x = 0;
    for (int i = 0; i < 12; ++i) x ^= c.v[i] ^ d.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

} // namespace c12

using namespace c12;

extern "C" {

int c12381_init(int device)
{
    Ctx& c = ctx();
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(C12381_ENODEV, "no CUDA device visible: this library has no CPU fallback", e);
    }
    if (device < 0 || device >= count) return set_error(C12381_EARG, "device index out of range");
    if (c.device == device) return C12381_OK;
    if (c.device >= 0) c12381_shutdown();
    C12_CUDA(cudaSetDevice(device));
    C12_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    C12_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    for (auto& ev : c.copy_ev) C12_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    C12_CUDA(cudaEventCreateWithFlags(&c.arena_ev, cudaEventDisableTiming));
    for (auto& ev : c.msm_ev) C12_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c.side_ev) C12_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c.group_ev) C12_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c.sgroup_ev) C12_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c.parse_ev) C12_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    {
        int least = 0, greatest = 0;
        C12_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        C12_CUDA(cudaStreamCreateWithPriority(&c.front_stream, cudaStreamNonBlocking, greatest));
        C12_CUDA(cudaStreamCreateWithPriority(&c.plan_stream, cudaStreamNonBlocking, greatest));
        // the side streams outrank the caller's (default = lowest priority): pipeline 1 - the high windows of the split tail - gets its
        // blocks placed first, finishes its rounds early, and its reduction runs under pipeline 0's last rounds
        for (auto& st : c.side) C12_CUDA(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, greatest));
    }
    C12_CUDA(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, device));
    C12_CUDA(cudaMallocHost(&c.h_flags, 64 * sizeof(int)));
    C12_CUDA(cudaMalloc(&c.d_flags, 64 * sizeof(int)));
    C12_CUDA(cudaMemset(c.d_flags, 0, 64 * sizeof(int)));
    for (auto& ev : c.ev) C12_CUDA(cudaEventCreate(&ev));
    for (auto& ev : c.pev) C12_CUDA(cudaEventCreate(&ev));
    c.device = device;
    c.launches = 0;
    return C12381_OK;
}

void c12381_shutdown(void)
{
    Ctx& c = ctx();
    if (c.device < 0) return;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();       // `_dev` calls may still run on the caller's streams
    if (c.arena) cudaFree(c.arena);
    for (auto& t : c.fb_table)
        if (t) cudaFree(t);
    if (c.d_flags) cudaFree(c.d_flags);
    if (c.h_flags) cudaFreeHost(c.h_flags);
    for (auto& ev : c.ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : c.pev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : c.copy_ev)
        if (ev) cudaEventDestroy(ev);
    if (c.arena_ev) cudaEventDestroy(c.arena_ev);
    for (auto& ev : c.side_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : c.group_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : c.parse_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : c.sgroup_ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& st : c.side)
        if (st) cudaStreamDestroy(st);
    if (c.plan_stream) cudaStreamDestroy(c.plan_stream);
    if (c.front_stream) cudaStreamDestroy(c.front_stream);
    for (auto& ev : c.msm_ev)
        if (ev) cudaEventDestroy(ev);
    if (c.copy_stream) {
        cudaStreamSynchronize(c.copy_stream);
        cudaStreamDestroy(c.copy_stream);
    }
    cudaStreamDestroy(c.stream);
    c = Ctx();
}

const char* c12381_last_error(void) { return ctx().err.c_str(); }
int c12381_device(void) { return ctx().device; }
void c12381_set_msm_window(int c) { ctx().forced_window = c; }
void c12381_set_msm_batch_affine(int rounds) { ctx().ba_rounds = rounds < 0 ? -1 : (rounds > 16 ? 16 : rounds); }
void c12381_set_knob(int id, int value)
{
    if (id >= 0 && id < 4) ctx().knob[id] = value;
    if (id == 4) ctx().upload_groups = value < 1 ? 1 : (value > 8 ? 8 : value);
    if (id == 5) ctx().front_end = value ? 1 : 0;
    if (id == 6) ctx().parse_aside = value ? 1 : 0;
    if (id == 7) ctx().split_tail = value < 0 || value > 2 ? 0 : value;
    if (id == 8) ctx().ba_fill_pct = value < 10 ? 10 : (value > 800 ? 800 : value);
}
void c12381_set_msm_pipelines(int pipes) { ctx().ba_pipes = pipes < 1 ? 1 : (pipes > 4 ? 4 : pipes); }
unsigned long long c12381_launch_count(void) { return ctx().launches; }

int c12381_last_msm_stats(double* accumulate_ms, double* total_ms, unsigned long long* bucket_adds, int* window_bits)
{
    MsmStats& s = ctx().stats;
    if (s.accumulate_ms < 0 && ctx().device >= 0) {
        float a = 0, t = 0;
        C12_CUDA(cudaEventSynchronize(ctx().ev[3]));
        C12_CUDA(cudaEventElapsedTime(&a, ctx().ev[1], ctx().ev[2]));
        C12_CUDA(cudaEventElapsedTime(&t, ctx().ev[0], ctx().ev[3]));
        s.accumulate_ms = a;
        s.total_ms = t;
        for (int i = 0; i < 8; ++i) {
            float p = 0;
            C12_CUDA(cudaEventElapsedTime(&p, ctx().pev[i], ctx().pev[i + 1]));
            s.phase_ms[i] = p;
        }
    }
    if (accumulate_ms) *accumulate_ms = s.accumulate_ms;
    if (total_ms) *total_ms = s.total_ms;
    if (bucket_adds) *bucket_adds = s.bucket_adds;
    if (window_bits) *window_bits = s.window_bits;
    return C12381_OK;
}

int c12381_last_msm_shape(int* ba_rounds, int* ba_pipelines, int* upload_groups)
{
    if (ba_rounds) *ba_rounds = ctx().stats.ba_rounds;
    if (ba_pipelines) *ba_pipelines = ctx().stats.ba_pipes;
    if (upload_groups) *upload_groups = ctx().stats.groups;
    return C12381_OK;
}

int c12381_last_msm_phases(double* phase_ms8)
{
    int rc = c12381_last_msm_stats(nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    if (!phase_ms8) return set_error(C12381_EARG, "last_msm_phases: null pointer");
    for (int i = 0; i < 8; ++i) phase_ms8[i] = ctx().stats.phase_ms[i];
    return C12381_OK;
}

int c12381_probe(int kind, int iters, double* out_gops, double* out_ms)
{
    C12_REQUIRE_CTX();
    Ctx& c = ctx();
    if (kind < 0 || kind > 4 || iters <= 0 || !out_gops) return set_error(C12381_EARG, "c12381_probe: bad argument");
    int sms = 0;
    C12_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device));
    const int blocks = sms * 8;
    if (int rc = arena_begin((size_t)blocks * PROBE_THREADS * 4, c.stream)) return rc;
    uint32_t* out = (uint32_t*)arena_take((size_t)blocks * PROBE_THREADS * 4);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {  // first repetition is the warm-up
        C12_CUDA(cudaEventRecord(c.ev[0], c.stream));
        switch (kind) {
        case 0: k_probe_imad<<<blocks, PROBE_THREADS, 0, c.stream>>>(out, iters, 12381u); break;
        case 1: k_probe_madc<<<blocks, PROBE_THREADS, 0, c.stream>>>(out, iters, 12381u); break;
        case 2: k_probe_wide<<<blocks, PROBE_THREADS, 0, c.stream>>>(out, iters, 12381u); break;
        default: k_probe_fp<<<blocks, PROBE_THREADS, 0, c.stream>>>(out, iters, 12381u, kind == 4); break;
        }
        C12_LAUNCHED();
        C12_CUDA(cudaEventRecord(c.ev[1], c.stream));
        C12_CUDA(cudaEventSynchronize(c.ev[1]));
        C12_CUDA(cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]));
    }
    double threads = (double)blocks * PROBE_THREADS;
    double ops = kind <= 2 ? threads * iters * 64.0 : threads * iters * 2.0;
    *out_gops = ops / (ms * 1e-3) / 1e9;
    if (out_ms) *out_ms = ms;
    return C12381_OK;
}

} // extern "C"
