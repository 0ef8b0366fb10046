// Host-side context shared by the translation units of libc12381_cuda.so: the bound device, its stream, a
// grow-only device scratch arena, the launch counter and the last error string.  One context per process
// (one process per GPU); not re-entrant (SURVEY §8b "Threading").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/c12381_cuda.h"

namespace c12 {

struct MsmStats {
    double accumulate_ms = 0, total_ms = 0;
    double phase_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // recode, sort, bounds+order, parse, accumulate, segment sums, plane sums, finish
    unsigned long long bucket_adds = 0;
    int window_bits = 0;
    int ba_rounds = 0, ba_pipes = 0, groups = 1;     // batch-affine halving rounds / pipelines / upload groups the last MSM ran with
};

struct Ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host entries: input uploads that overlap the first pipeline stages
    cudaEvent_t copy_ev[3] = {nullptr, nullptr, nullptr};
    std::string err;
    unsigned long long launches = 0;
    // scratch arena: one allocation, bump-allocated per call, grown (after a stream sync) when too small
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0;
    size_t arena_used = 0;
    int arena_depth = 0;   // > 0 while a host entry runs its pipeline: nested arena_begin calls carve from the same reservation
    // stream ordering of the ONE arena: the stream of the call that carved from it last.  A call arriving on another stream first
    // waits (on the device, cudaStreamWaitEvent) for everything enqueued on that stream so far - the previous call's kernels still
    // read the scratch the new call is about to overwrite; before the arena is freed to grow it, that event is waited for on the host.
    cudaStream_t arena_stream = nullptr;
    bool arena_stream_valid = false;
    cudaEvent_t arena_ev = nullptr;
    int host_depth = 0;    // > 0 inside a host-pointer entry: its kernels report malformed input in a flag word of their own
    bool pc_configured = false;   // cooperative pairing kernels: dynamic shared memory opt-in done on THIS device
    int sm_count = 0;
    // pinned staging for small results / flags
    // d_flags[0]: the `_dev` entries' word, collected by c12381_sync_status; d_flags[HOST_FLAG_WORD]: the host entries' word,
    // reset and collected inside each call - so a host call can neither wipe nor inherit what an earlier `_dev` call flagged
    int* h_flags = nullptr;
    int* d_flags = nullptr;
    int forced_window = 0;
    int ba_rounds = -1;     // batch-affine halving rounds in front of the XYZZ accumulation: -1 = chosen from the bucket load, 0 = none, k = k rounds
    int ba_pipes = 2;       // independent round pipelines (groups of windows on their own streams: one's inversion kernel hides behind the other's additions)
    cudaStream_t front_stream = nullptr;   // highest priority.  One group: the scalar-only stages run here, the parse of the points beside them on the caller's stream (the default priority is the LOWEST there is, so it is the latency-bound chain that gets the high one; the parse fills what it leaves idle)
    cudaEvent_t parse_ev[2] = {nullptr, nullptr};   // fork, join of that
    int parse_aside = 1;
    int ba_fill_pct = 100;  // knob 8 (A/B): share of the resident warps (3 blocks per SM counted) one lane's round is sized for, in percent
    int split_tail = 2;     // knob 7: 0 = one common tail; 1 = EARLY split (the high windows as the smaller of two pipelines, their whole tail under the low windows' last rounds: tail 1.24 -> 1.02 ms but the uneven rounds give it back, profiles/r03g); 2 (default) = LATE split (one accumulation, then the high half's reduction + Horner chain on a side stream beside the low half's: G1 n = 2^20 -0.7 %, n = 2^16 -5 %, n = 2^10 -8 %, G2 unchanged, profiles/r04c)
    int front_end = 0;      // bucket lists: 0 = by counting (atomic ranks + scan + scatter), 1 = segmented radix sort + bounds search (stable; the A/B twin)
    int upload_groups = 4;  // host-pointer MSM entries: the points go up in this many groups, each in front of its own pipeline of halving rounds
    cudaEvent_t group_ev[8] = {}, sgroup_ev[8] = {};   // a group's points / scalars have arrived
    int knob[4] = {1, 32, 3, 0};   // c12381_set_knob: [0] waves a pipeline round should span, [1] largest J, [2] halvings left to the XYZZ accumulation, [3] threads the segment running sums of the bucket reduction should fill (0 = default)
    cudaStream_t side[7] = {};      // streams of lanes 1 .. 7 (lane 0 runs on the caller's stream)
    cudaStream_t plan_stream = nullptr;                      // several upload groups: the merged plan, beside the groups' own stages
    cudaEvent_t msm_ev[9] = {};   // a group's bucket bounds are in place (0 .. 7); the merged plan is (8)
    cudaEvent_t side_ev[8] = {};   // fork + one join per side stream
    void* fb_table[2] = {nullptr, nullptr};   // fixed-base window tables (G1, G2), built on first use
    MsmStats stats;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // MSM phase boundaries
};

Ctx& ctx();
constexpr int HOST_FLAG_WORD = 16;
inline int* flags_word() { Ctx& c = ctx(); return c.d_flags + (c.host_depth > 0 ? HOST_FLAG_WORD : 0); }
int set_error(int code, const char* what, cudaError_t e = cudaSuccess);

// Reserve `bytes` of scratch for the CURRENT call.  Call arena_begin(total) once per entry point with an upper
// bound, then carve with arena_take.
int arena_begin(size_t total_bytes, cudaStream_t s);
void* arena_take(size_t bytes);

#define C12_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) return c12::set_error(C12381_ECUDA, #call, e__);            \
    } while (0)

#define C12_LAUNCHED()                                                                      \
    do {                                                                                    \
        c12::ctx().launches++;                                                              \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) return c12::set_error(C12381_ECUDA, "kernel launch", e__);  \
    } while (0)

// every entry re-asserts the context's device: the caller (or torch) may have changed the thread's current device
#define C12_REQUIRE_CTX()                                                                   \
    do {                                                                                    \
        if (c12::ctx().device < 0) return c12::set_error(C12381_ENODEV, "c12381_init was not called or no CUDA device"); \
        cudaError_t e__ = cudaSetDevice(c12::ctx().device);                                 \
        if (e__ != cudaSuccess) return c12::set_error(C12381_ECUDA, "cudaSetDevice", e__);  \
    } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

} // namespace c12
