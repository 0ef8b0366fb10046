// Batched pairings, pairing products and GT helpers: __global__ wrappers + C-ABI entries (include/c12381_cuda.h).
// One instance (a k-pair product sharing its squarings and ONE final exponentiation) per thread; the bodies are in
// pairing.cuh.  Replaces the one-at-a-time bridge calls pair_ate / pair_double_ate / pair_final_exponentiation /
// multiply(fp12&) / pow(fp12&) (reference: src/miracl_core_interface.cpp:251-289).
#include <stdlib.h>
#define C12_PAIRING_TU 1      // fp.cuh: scope of the -DC12_EXP_LAZY_ADDS timing experiment

// Measured defaults of this translation unit (profiles/r01ah_pairing_lockstep_ab.txt; -DC12_PAIR_NO_LOCKSTEP restores the old build):
//  * block-wide barriers keep the warps of a block at the same place of the straight-line code (pairing.cuh, C12_BLOCK_ALIGN):
//    ncu had shown instruction-fetch stalls growing with every extra resident warp (sm__icc hit rate 93 % -> 74 % from 8 to 12
//    warps per SM);
//  * with the instruction cache shared that way, the three Montgomery products of an Fp2 product can be inlined so that their
//    carry chains interleave (slower without the barriers: three times the code).
// Together +25 % on the thread-per-instance kernels (2.06 -> 2.57 M pairings/s at full waves), +4 % on the cooperative ones.
#if !defined(C12_PAIR_NO_LOCKSTEP)
#if !defined(C12_PAIR_LOCKSTEP)
#define C12_PAIR_LOCKSTEP 1
#endif
#if !defined(C12_FP2_INLINE_MULS) && !defined(C12_PAIR_NO_FP2_INLINE)
#define C12_FP2_INLINE_MULS 1
#endif
#endif

#include "msm_impl.cuh"
#include "pairing.cuh"
#include "pairing_coop.cuh"

namespace c12 {

#ifndef C12_PAIR_THREADS
#define C12_PAIR_THREADS 128      // two blocks per SM at 255 registers
#endif
constexpr int PAIR_THREADS = C12_PAIR_THREADS;
#ifndef C12_PAIR_MIN_BLOCKS
#define C12_PAIR_MIN_BLOCKS 1      // blocks per SM the register allocation is bounded for (A/B knob, profiles/)
#endif

// mode 0: raw Miller product (576 B); mode 1: final-exponentiated GT value (576 B); mode 2: verdict byte (GT == 1)
__global__ void __launch_bounds__(PAIR_THREADS, C12_PAIR_MIN_BLOCKS) k_pairing(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2, uint32_t B,
                                                          uint32_t k, int mode, uint8_t* __restrict__ out, int* flags)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
#if defined(C12_PAIR_LOCKSTEP)
    // every thread of the block goes through the bodies (they hold block-wide barriers): the tail threads redo the last instance
    const bool tail = b >= B;
    if (tail) b = B - 1;
    uint8_t buf[576];
    bool ok = pairing_product_body(g1 + 96ull * b * k, g2 + 192ull * b * k, k, mode == 2 ? 1 : mode, buf);
    if (tail) return;
    if (!ok) atomicOr(flags, FLAG_BAD_POINT);
    if (mode == 2) {
        uint32_t acc = buf[575] ^ 1u;
        for (int i = 0; i < 575; ++i) acc |= buf[i];
        out[b] = acc == 0 ? 1 : 0;
    } else {
        for (int i = 0; i < 576; ++i) out[576ull * b + i] = buf[i];
    }
#else
    if (b >= B) return;
    if (mode == 2) {
        uint8_t buf[576];
        bool ok = pairing_product_body(g1 + 96ull * b * k, g2 + 192ull * b * k, k, 1, buf);
        if (!ok) atomicOr(flags, FLAG_BAD_POINT);
        // canonical encoding of 1: the last byte (low byte of the real part of a.a) is 1, everything else 0
        uint32_t acc = buf[575] ^ 1u;
        for (int i = 0; i < 575; ++i) acc |= buf[i];
        out[b] = acc == 0 ? 1 : 0;
        return;
    }
    if (!pairing_product_body(g1 + 96ull * b * k, g2 + 192ull * b * k, k, mode, out + 576ull * b)) atomicOr(flags, FLAG_BAD_POINT);
#endif
}

__global__ void __launch_bounds__(PAIR_THREADS) k_final_exp(const uint8_t* __restrict__ in, uint32_t B, uint8_t* __restrict__ out)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
#if defined(C12_PAIR_LOCKSTEP)
    const bool tail = b >= B;
    if (tail) b = B - 1;
    uint8_t buf[576];
    final_exp_body(in + 576ull * b, buf);
    if (tail) return;
    for (int i = 0; i < 576; ++i) out[576ull * b + i] = buf[i];
#else
    if (b >= B) return;
    final_exp_body(in + 576ull * b, out + 576ull * b);
#endif
}

__global__ void __launch_bounds__(PAIR_THREADS) k_gt_mul(const uint8_t* __restrict__ a, const uint8_t* __restrict__ bb, uint32_t B,
                                                         uint8_t* __restrict__ out)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    gt_mul_body(a + 576ull * b, bb + 576ull * b, out + 576ull * b);
}

__global__ void __launch_bounds__(PAIR_THREADS) k_gt_pow(const uint8_t* __restrict__ a, const uint8_t* __restrict__ s32, uint32_t B,
                                                         uint8_t* __restrict__ out, int* flags)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (!scalar_is_canonical(scalar_from_be32(s32 + 32ull * b))) atomicOr(flags, FLAG_BAD_SCALAR);
    gt_pow_body(a + 576ull * b, s32 + 32ull * b, out + 576ull * b);
}

__global__ void __launch_bounds__(PAIR_THREADS) k_gt_pow_gs(const uint8_t* __restrict__ a, const uint8_t* __restrict__ s32, uint32_t B,
                                                            uint8_t* __restrict__ out, int* flags)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (!scalar_is_canonical(scalar_from_be32(s32 + 32ull * b))) atomicOr(flags, FLAG_BAD_SCALAR);
    gt_pow_gs_body(a + 576ull * b, s32 + 32ull * b, out + 576ull * b);
}

// ---- lane-cooperative kernels (pairing_coop.cuh): six lanes per instance, five instances per warp ---------------------
#ifndef C12_PAIR_WARPS
#define C12_PAIR_WARPS 4            // warps per block: 20 instances, 57,600 B of shared memory; three blocks per SM
#endif
#ifndef C12_PAIR_BLOCKS
#define C12_PAIR_BLOCKS 2           // resident blocks per SM the register allocation is bounded for (3 measured 10 % slower)
#endif
constexpr int PC_WARPS = C12_PAIR_WARPS;
constexpr int PC_THREADS = PC_WARPS * 32;
constexpr int PC_INST = PC_WARPS * pc::INST_PER_WARP;
constexpr size_t PC_SMEM = (size_t)PC_INST * sizeof(pc::InstSmem);

__device__ __forceinline__ pc::Lane pc_lane(unsigned char* smem, uint32_t B, uint32_t& b)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool lane_ok = lane < pc::GROUP * pc::INST_PER_WARP;
    const int gi = lane_ok ? lane / pc::GROUP : 0;
    b = blockIdx.x * PC_INST + warp * pc::INST_PER_WARP + gi;
    pc::Lane L;
    L.r = lane % pc::GROUP;
    L.on = lane_ok && b < B;
    L.s = reinterpret_cast<pc::InstSmem*>(smem) + warp * pc::INST_PER_WARP + gi;
    return L;
}

// lane r writes coefficient r in the wire order / verdict of the whole group
__device__ __forceinline__ void pc_emit(pc::Lane L, int mode, uint32_t b, uint8_t* out)
{
    Fp2 c = pc::ld(L.s->f + L.r);
    if (mode == 2) {
        bool ok = eq(c, L.r == 0 ? fp2_one() : fp2_zero());
        const unsigned base = (threadIdx.x & 31) / pc::GROUP * pc::GROUP;
        unsigned m = (__ballot_sync(0xffffffffu, ok) >> base) & 0x3fu;
        if (L.on && L.r == 0) out[b] = m == 0x3fu ? 1 : 0;
        return;
    }
    if (L.on) fp2_to_bytes(out + 576ull * b + pc::wire_offset(L.r), c);
}

// mode 0: raw Miller product; 1: final-exponentiated GT value; 2: verdict byte (GT == 1)
__global__ void __launch_bounds__(PC_THREADS, C12_PAIR_BLOCKS) k_pairing_coop(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2, uint32_t B, uint32_t k,
                                                              int mode, uint8_t* __restrict__ out, pc::PairIn* __restrict__ pin_all,
                                                              Fp2* __restrict__ gscratch, int* flags)
{
    extern __shared__ __align__(16) unsigned char pc_smem[];
    uint32_t b;
    pc::Lane L = pc_lane(pc_smem, B, b);
    pc::PairIn* pin = pin_all + (size_t)(L.on ? b : 0) * k;
    Fp2* g = gscratch + (size_t)(L.on ? b : 0) * 6;
    // parse: lane r < 4 takes pairs r, r + 4 (the pairs it will own in the Miller passes)
    if (L.on && L.r < pc::CHUNK) {
        for (uint32_t j = L.r; j < k; j += pc::CHUNK) {
            pc::PairIn in;
            bool ok = g1_from_bytes96(in.P, g1 + 96ull * ((size_t)b * k + j));
            ok = g2_from_bytes192(in.Q, g2 + 192ull * ((size_t)b * k + j)) && ok;
            if (!ok) atomicOr(flags, FLAG_BAD_POINT);
            pin[j] = in;
        }
    }
    __syncwarp();
    for (uint32_t c0 = 0; c0 < k; c0 += pc::CHUNK) {
        const int kk = (int)(k - c0 < (uint32_t)pc::CHUNK ? k - c0 : (uint32_t)pc::CHUNK);
        if (c0) {   // park the product so far
            if (L.on) g[L.r] = L.s->f[L.r];
            __syncwarp();
        }
        pc::miller_coop(L, pin + c0, kk);
        if (c0) {
            pc::st(L.s->t + L.r, L.on ? g[L.r] : pc::ld(L.s->t + L.r), L.on);
            __syncwarp();
            pc::full_mul(L, L.s->f, L.s->f, L.s->t);
        }
    }
    if (mode >= 1) pc::final_exp_coop(L, g);
    pc_emit(L, mode, b, out);
}

__device__ __forceinline__ void pc_load_wire(pc::Lane L, Fp2* dst, const uint8_t* in576)
{
    pc::st(dst + L.r, L.on ? fp2_from_bytes(in576 + pc::wire_offset(L.r)) : fp2_zero(), L.on);
    __syncwarp();
}

__global__ void __launch_bounds__(PC_THREADS, C12_PAIR_BLOCKS) k_final_exp_coop(const uint8_t* __restrict__ in, uint32_t B, uint8_t* __restrict__ out, Fp2* __restrict__ gscratch)
{
    extern __shared__ __align__(16) unsigned char pc_smem[];
    uint32_t b;
    pc::Lane L = pc_lane(pc_smem, B, b);
    pc_load_wire(L, L.s->f, in + 576ull * (L.on ? b : 0));
    pc::final_exp_coop(L, gscratch + (size_t)(L.on ? b : 0) * 6);
    pc_emit(L, 1, b, out);
}

__global__ void __launch_bounds__(PC_THREADS) k_gt_mul_coop(const uint8_t* __restrict__ a, const uint8_t* __restrict__ bb, uint32_t B, uint8_t* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char pc_smem[];
    uint32_t b;
    pc::Lane L = pc_lane(pc_smem, B, b);
    pc_load_wire(L, L.s->f, a + 576ull * (L.on ? b : 0));
    pc_load_wire(L, L.s->t, bb + 576ull * (L.on ? b : 0));
    pc::full_mul(L, L.s->f, L.s->f, L.s->t);
    pc_emit(L, 1, b, out);
}

// a^k for unitary a (FP12_pow, fp12_BLS12381.cpp:736-777): square-and-multiply with Granger-Scott squarings
__global__ void __launch_bounds__(PC_THREADS, C12_PAIR_BLOCKS) k_gt_pow_coop(const uint8_t* __restrict__ a, const uint8_t* __restrict__ s32, uint32_t B,
                                                             uint8_t* __restrict__ out, int* flags)
{
    extern __shared__ __align__(16) unsigned char pc_smem[];
    uint32_t b;
    pc::Lane L = pc_lane(pc_smem, B, b);
    Scalar256 e = scalar_from_be32(s32 + 32ull * (L.on ? b : 0));
    if (L.on && L.r == 0 && !scalar_is_canonical(e)) atomicOr(flags, FLAG_BAD_SCALAR);
    if (!L.on)
        for (int i = 0; i < 8; ++i) e.v[i] = 0;
    pc_load_wire(L, L.s->t, a + 576ull * (L.on ? b : 0));           // base
    pc::st(L.s->f + L.r, L.r == 0 ? fp2_one() : fp2_zero(), L.on);   // accumulator
    __syncwarp();
    // the exponents of the instances sharing a warp differ: every step runs both operations and keeps what applies
    bool started = false;
#pragma unroll 1
    for (int i = 255; i >= 0; --i) {
        const bool bit = (e.v[i >> 5] >> (i & 31)) & 1u;
        if (__any_sync(0xffffffffu, started)) {
            Fp2 keep = pc::ld(L.s->f + L.r);
            pc::cyclo_sqr(L, L.s->f);
            if (!started) pc::st(L.s->f + L.r, keep, L.on);
            __syncwarp();
        }
        if (__any_sync(0xffffffffu, bit)) {
            Fp2 keep = pc::ld(L.s->f + L.r);
            pc::full_mul(L, L.s->f, L.s->f, L.s->t);
            if (!bit) pc::st(L.s->f + L.r, keep, L.on);
            __syncwarp();
        }
        started = started || bit;
    }
    pc_emit(L, 1, b, out);
}

// Two implementations of every entry: thread-per-instance (pairing.cuh bodies) and six-lanes-per-instance
// (pairing_coop.cuh).  Measured (profiles/r01ac, r01t): the cooperative kernels run at ~1.8 M pairings/s from 2^14 four-pair
// instances up and have six times the parallelism per instance (a single product: 9 ms against 47 ms); the
// thread-per-instance kernel is latency-bound per thread, ~59 ms per wave of SMs x 256 threads however full the wave is,
// i.e. 2.57 M/s x the fill of its last wave (profiles/r01ah).  So: thread-per-instance when the batch fills its waves to
// >= 76 %, the cooperative kernels otherwise.  C12381_PAIRING=scalar|coop or c12381_set_pairing_kernel force one of them.
static bool scalar_fills_its_waves(size_t B)
{
    const int sms = ctx().sm_count > 0 ? ctx().sm_count : 148;
    const size_t wave = (size_t)sms * 256;
    const size_t waves = (B + wave - 1) / wave;
    return B >= 16384 && (double)B >= 0.76 * (double)(waves * wave);
}

static int g_pairing_kernel = -1;   // 0 automatic, 1 thread-per-instance, 2 cooperative (c12381_set_pairing_kernel)

static bool use_scalar_kernels(size_t B)
{
    if (g_pairing_kernel < 0) {
        const char* e = getenv("C12381_PAIRING");
        g_pairing_kernel = !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 'c' ? 2 : 0));
    }
    if (g_pairing_kernel == 1) return true;
    if (g_pairing_kernel == 2) return false;
    return scalar_fills_its_waves(B);
}

static int pc_configure()
{
    bool& done = ctx().pc_configured;      // per context: the opt-in is per device and c12381_init may rebind
    if (done) return C12381_OK;
    C12_CUDA(cudaFuncSetAttribute(k_pairing_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC_SMEM));
    C12_CUDA(cudaFuncSetAttribute(k_final_exp_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC_SMEM));
    C12_CUDA(cudaFuncSetAttribute(k_gt_mul_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC_SMEM));
    C12_CUDA(cudaFuncSetAttribute(k_gt_pow_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC_SMEM));
    done = true;
    return C12381_OK;
}

static size_t pairing_scratch(size_t B, int k) { return align_up(B * (size_t)k * sizeof(pc::PairIn)) + align_up(B * 6 * sizeof(Fp2)) + 4096; }
static size_t gt_scratch(size_t B) { return align_up(B * 6 * sizeof(Fp2)) + 4096; }

// the caller has arena_begin()'d at least pairing_scratch(B, k) beyond what it took itself
static int pairing_run(const uint8_t* d_g1, const uint8_t* d_g2, size_t B, int k, int mode, uint8_t* d_out, cudaStream_t s)
{
    if (k < 1 || k > C12381_MAX_PAIRS) return set_error(C12381_EARG, "pairing: k must be in [1, C12381_MAX_PAIRS]");
    if (B == 0) return C12381_OK;
    if (B > 0x7fffffffull) return set_error(C12381_EARG, "pairing: too many instances");
    if (use_scalar_kernels(B)) {
        k_pairing<<<cdiv(B, PAIR_THREADS), PAIR_THREADS, 0, s>>>(d_g1, d_g2, (uint32_t)B, (uint32_t)k, mode, d_out, flags_word());
        C12_LAUNCHED();
        return C12381_OK;
    }
    int rc = pc_configure();
    if (rc) return rc;
    pc::PairIn* pin = (pc::PairIn*)arena_take(B * (size_t)k * sizeof(pc::PairIn));
    Fp2* g = (Fp2*)arena_take(B * 6 * sizeof(Fp2));
    if (!g) return set_error(C12381_ECUDA, "pairing: scratch arena bound too small");
    k_pairing_coop<<<cdiv(B, PC_INST), PC_THREADS, PC_SMEM, s>>>(d_g1, d_g2, (uint32_t)B, (uint32_t)k, mode, d_out, pin, g, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}

static int pairing_host(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, int mode, uint8_t* out)
{
    C12_REQUIRE_CTX();
    if (k < 1 || k > C12381_MAX_PAIRS) return set_error(C12381_EARG, "pairing: k must be in [1, C12381_MAX_PAIRS]");
    if (B && (!g1s || !g2s || !out)) return set_error(C12381_EARG, "pairing: null pointer");
    const void* in[2] = {g1s, g2s};
    size_t sz[2] = {B * k * 96, B * k * 192};
    return with_staged(in, sz, 2, out, mode == 2 ? B : B * 576, pairing_scratch(B, k), [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        return pairing_run(d_in[0], d_in[1], B, k, mode, d_out, s);
    });
}

static int pairing_dev(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, int mode, uint8_t* out, void* stream)
{
    C12_REQUIRE_CTX();
    if (k < 1 || k > C12381_MAX_PAIRS) return set_error(C12381_EARG, "pairing: k must be in [1, C12381_MAX_PAIRS]");
    if (B && (!g1s || !g2s || !out)) return set_error(C12381_EARG, "pairing: null pointer");
    int rc = arena_begin(pairing_scratch(B, k), pick_stream(stream));
    if (rc) return rc;
    return pairing_run(g1s, g2s, B, k, mode, out, pick_stream(stream));
}

// callers have arena_begin()'d at least gt_scratch(B) beyond what they took themselves
static int final_exp_run(const uint8_t* d_in, size_t B, uint8_t* d_out, cudaStream_t s)
{
    if (B == 0) return C12381_OK;
    if (B > 0x7fffffffull) return set_error(C12381_EARG, "final_exp: too many instances");
    if (use_scalar_kernels(B)) {
        k_final_exp<<<cdiv(B, PAIR_THREADS), PAIR_THREADS, 0, s>>>(d_in, (uint32_t)B, d_out);
        C12_LAUNCHED();
        return C12381_OK;
    }
    int rc = pc_configure();
    if (rc) return rc;
    Fp2* g = (Fp2*)arena_take(B * 6 * sizeof(Fp2));
    if (!g) return set_error(C12381_ECUDA, "final_exp: scratch arena bound too small");
    k_final_exp_coop<<<cdiv(B, PC_INST), PC_THREADS, PC_SMEM, s>>>(d_in, (uint32_t)B, d_out, g);
    C12_LAUNCHED();
    return C12381_OK;
}
static int gt_mul_run(const uint8_t* a, const uint8_t* b, size_t B, uint8_t* d_out, cudaStream_t s)
{
    if (B == 0) return C12381_OK;
    if (B > 0x7fffffffull) return set_error(C12381_EARG, "gt_mul: too many instances");
    if (use_scalar_kernels(B)) {
        k_gt_mul<<<cdiv(B, PAIR_THREADS), PAIR_THREADS, 0, s>>>(a, b, (uint32_t)B, d_out);
        C12_LAUNCHED();
        return C12381_OK;
    }
    int rc = pc_configure();
    if (rc) return rc;
    k_gt_mul_coop<<<cdiv(B, PC_INST), PC_THREADS, PC_SMEM, s>>>(a, b, (uint32_t)B, d_out);
    C12_LAUNCHED();
    return C12381_OK;
}
static int gt_pow_run(const uint8_t* a, const uint8_t* sc, size_t B, uint8_t* d_out, cudaStream_t s, bool gs = false)
{
    if (B == 0) return C12381_OK;
    if (B > 0x7fffffffull) return set_error(C12381_EARG, "gt_pow: too many instances");
    if (gs) {
        k_gt_pow_gs<<<cdiv(B, PAIR_THREADS), PAIR_THREADS, 0, s>>>(a, sc, (uint32_t)B, d_out, flags_word());
        C12_LAUNCHED();
        return C12381_OK;
    }
    if (use_scalar_kernels(B)) {
        k_gt_pow<<<cdiv(B, PAIR_THREADS), PAIR_THREADS, 0, s>>>(a, sc, (uint32_t)B, d_out, flags_word());
        C12_LAUNCHED();
        return C12381_OK;
    }
    int rc = pc_configure();
    if (rc) return rc;
    k_gt_pow_coop<<<cdiv(B, PC_INST), PC_THREADS, PC_SMEM, s>>>(a, sc, (uint32_t)B, d_out, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}

} // namespace c12

using namespace c12;

extern "C" {
void c12381_set_pairing_kernel(int mode) { g_pairing_kernel = mode == 1 || mode == 2 ? mode : 0; }
int c12381_miller_batch(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, uint8_t* out576) { return pairing_host(g1s, g2s, B, k, 0, out576); }
int c12381_pairing_product_batch(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, uint8_t* out576) { return pairing_host(g1s, g2s, B, k, 1, out576); }
int c12381_pairing_check_batch(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, uint8_t* verdicts) { return pairing_host(g1s, g2s, B, k, 2, verdicts); }
int c12381_miller_batch_dev(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, uint8_t* o, void* st) { return pairing_dev(g1s, g2s, B, k, 0, o, st); }
int c12381_pairing_product_batch_dev(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, uint8_t* o, void* st) { return pairing_dev(g1s, g2s, B, k, 1, o, st); }
int c12381_pairing_check_batch_dev(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, uint8_t* o, void* st) { return pairing_dev(g1s, g2s, B, k, 2, o, st); }

int c12381_final_exp_batch(const uint8_t* in576, size_t B, uint8_t* out576)
{
    C12_REQUIRE_CTX();
    if (B && (!in576 || !out576)) return set_error(C12381_EARG, "final_exp: null pointer");
    const void* in[1] = {in576};
    size_t sz[1] = {B * 576};
    return with_staged(in, sz, 1, out576, B * 576, gt_scratch(B), [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) { return final_exp_run(d_in[0], B, d_out, s); });
}
int c12381_final_exp_batch_dev(const uint8_t* d_in576, size_t B, uint8_t* d_out576, void* stream)
{
    C12_REQUIRE_CTX();
    if (B && (!d_in576 || !d_out576)) return set_error(C12381_EARG, "final_exp: null pointer");
    int rc = arena_begin(gt_scratch(B), pick_stream(stream));
    if (rc) return rc;
    return final_exp_run(d_in576, B, d_out576, pick_stream(stream));
}

int c12381_gt_mul_batch(const uint8_t* a576, const uint8_t* b576, size_t B, uint8_t* out576)
{
    C12_REQUIRE_CTX();
    if (B && (!a576 || !b576 || !out576)) return set_error(C12381_EARG, "gt_mul: null pointer");
    const void* in[2] = {a576, b576};
    size_t sz[2] = {B * 576, B * 576};
    return with_staged(in, sz, 2, out576, B * 576, 0, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) { return gt_mul_run(d_in[0], d_in[1], B, d_out, s); });
}
int c12381_gt_mul_batch_dev(const uint8_t* a, const uint8_t* b, size_t B, uint8_t* o, void* stream)
{
    C12_REQUIRE_CTX();
    if (B && (!a || !b || !o)) return set_error(C12381_EARG, "gt_mul: null pointer");
    return gt_mul_run(a, b, B, o, pick_stream(stream));
}
int c12381_gt_pow_batch(const uint8_t* a576, const uint8_t* scalars32, size_t B, uint8_t* out576)
{
    C12_REQUIRE_CTX();
    if (B && (!a576 || !scalars32 || !out576)) return set_error(C12381_EARG, "gt_pow: null pointer");
    const void* in[2] = {a576, scalars32};
    size_t sz[2] = {B * 576, B * 32};
    return with_staged(in, sz, 2, out576, B * 576, 0, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) { return gt_pow_run(d_in[0], d_in[1], B, d_out, s); });
}
int c12381_gt_pow_gs_batch(const uint8_t* a576, const uint8_t* scalars32, size_t B, uint8_t* out576)
{
    C12_REQUIRE_CTX();
    if (B && (!a576 || !scalars32 || !out576)) return set_error(C12381_EARG, "gt_pow_gs: null pointer");
    const void* in[2] = {a576, scalars32};
    size_t sz[2] = {B * 576, B * 32};
    return with_staged(in, sz, 2, out576, B * 576, 0, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) { return gt_pow_run(d_in[0], d_in[1], B, d_out, s, true); });
}
int c12381_gt_pow_gs_batch_dev(const uint8_t* a, const uint8_t* sc, size_t B, uint8_t* o, void* stream)
{
    C12_REQUIRE_CTX();
    if (B && (!a || !sc || !o)) return set_error(C12381_EARG, "gt_pow_gs: null pointer");
    return gt_pow_run(a, sc, B, o, pick_stream(stream), true);
}
int c12381_gt_pow_batch_dev(const uint8_t* a, const uint8_t* sc, size_t B, uint8_t* o, void* stream)
{
    C12_REQUIRE_CTX();
    if (B && (!a || !sc || !o)) return set_error(C12381_EARG, "gt_pow: null pointer");
    return gt_pow_run(a, sc, B, o, pick_stream(stream));
}
}
