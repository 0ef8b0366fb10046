// Batched pairings, pairing products and GT helpers: __global__ wrappers + C-ABI entries (include/c12381_cuda.h).
// One instance (a k-pair product sharing its squarings and ONE final exponentiation) per thread; the bodies are in
// pairing.cuh.  Replaces the one-at-a-time bridge calls pair_ate / pair_double_ate / pair_final_exponentiation /
// multiply(fp12&) / pow(fp12&) (reference: src/miracl_core_interface.cpp:251-289).
#include "msm_impl.cuh"
#include "pairing.cuh"

namespace c12 {

constexpr int PAIR_THREADS = 64;
#ifndef C12_PAIR_MIN_BLOCKS
#define C12_PAIR_MIN_BLOCKS 1      // blocks per SM the register allocation is bounded for (A/B knob, profiles/)
#endif

// mode 0: raw Miller product (576 B); mode 1: final-exponentiated GT value (576 B); mode 2: verdict byte (GT == 1)
__global__ void __launch_bounds__(PAIR_THREADS, C12_PAIR_MIN_BLOCKS) k_pairing(const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2, uint32_t B,
                                                          uint32_t k, int mode, uint8_t* __restrict__ out, int* flags)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
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
}

__global__ void __launch_bounds__(PAIR_THREADS) k_final_exp(const uint8_t* __restrict__ in, uint32_t B, uint8_t* __restrict__ out)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    final_exp_body(in + 576ull * b, out + 576ull * b);
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

static int pairing_run(const uint8_t* d_g1, const uint8_t* d_g2, size_t B, int k, int mode, uint8_t* d_out, cudaStream_t s)
{
    if (k < 1 || k > C12381_MAX_PAIRS) return set_error(C12381_EARG, "pairing: k must be in [1, C12381_MAX_PAIRS]");
    if (B == 0) return C12381_OK;
    if (B > 0x7fffffffull) return set_error(C12381_EARG, "pairing: too many instances");
    k_pairing<<<cdiv(B, PAIR_THREADS), PAIR_THREADS, 0, s>>>(d_g1, d_g2, (uint32_t)B, (uint32_t)k, mode, d_out, ctx().d_flags);
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
    return with_staged(in, sz, 2, out, mode == 2 ? B : B * 576, 0, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        return pairing_run(d_in[0], d_in[1], B, k, mode, d_out, s);
    });
}

static int pairing_dev(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, int mode, uint8_t* out, void* stream)
{
    C12_REQUIRE_CTX();
    if (B && (!g1s || !g2s || !out)) return set_error(C12381_EARG, "pairing: null pointer");
    return pairing_run(g1s, g2s, B, k, mode, out, pick_stream(stream));
}

static int final_exp_run(const uint8_t* d_in, size_t B, uint8_t* d_out, cudaStream_t s)
{
    if (B == 0) return C12381_OK;
    k_final_exp<<<cdiv(B, PAIR_THREADS), PAIR_THREADS, 0, s>>>(d_in, (uint32_t)B, d_out);
    C12_LAUNCHED();
    return C12381_OK;
}
static int gt_mul_run(const uint8_t* a, const uint8_t* b, size_t B, uint8_t* d_out, cudaStream_t s)
{
    if (B == 0) return C12381_OK;
    k_gt_mul<<<cdiv(B, PAIR_THREADS), PAIR_THREADS, 0, s>>>(a, b, (uint32_t)B, d_out);
    C12_LAUNCHED();
    return C12381_OK;
}
static int gt_pow_run(const uint8_t* a, const uint8_t* sc, size_t B, uint8_t* d_out, cudaStream_t s)
{
    if (B == 0) return C12381_OK;
    k_gt_pow<<<cdiv(B, PAIR_THREADS), PAIR_THREADS, 0, s>>>(a, sc, (uint32_t)B, d_out, ctx().d_flags);
    C12_LAUNCHED();
    return C12381_OK;
}

} // namespace c12

using namespace c12;

extern "C" {
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
    return with_staged(in, sz, 1, out576, B * 576, 0, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) { return final_exp_run(d_in[0], B, d_out, s); });
}
int c12381_final_exp_batch_dev(const uint8_t* d_in576, size_t B, uint8_t* d_out576, void* stream)
{
    C12_REQUIRE_CTX();
    if (B && (!d_in576 || !d_out576)) return set_error(C12381_EARG, "final_exp: null pointer");
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
int c12381_gt_pow_batch_dev(const uint8_t* a, const uint8_t* sc, size_t B, uint8_t* o, void* stream)
{
    C12_REQUIRE_CTX();
    if (B && (!a || !sc || !o)) return set_error(C12381_EARG, "gt_pow: null pointer");
    return gt_pow_run(a, sc, B, o, pick_stream(stream));
}
}
