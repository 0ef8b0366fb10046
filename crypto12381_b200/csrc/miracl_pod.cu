// Drop-in entries on the reference's in-memory PODs (include/crypto12381/miracl_core_interface.hpp:
// big :32-33, fp :76-79, point1 :80-87 region, fp2/point2 :130-141, fp4/fp12 :172-184).  The forwarding translation
// unit passes its arguments straight through; the structs are uploaded AS THEY ARE and converted
// on the device:
//   big   = int64[7], 58-bit digits (possibly un-normalised)            -> 32-byte big-endian scalar
//   fp    = { big g; int32 xes }: Montgomery residue with R = 2^406, value < xes * p (lazy reduction)
//           -> our residue (R = 2^384) by  x R384 = L * 2^-22 + H * 2^362   with g = L + 2^384 H
//              = mont_mul(2^362, L) + mont_mul(2^746, H)
//   ours -> fp : g = mont_mul(ours, 2^406) (canonical, < p), xes = 1; fp12.type = FP_DENSE (5)
// Projective inputs (X:Y:Z) are normalised to affine on the device and fed to the same byte-format pipelines as the
// batched entries; nothing here computes on the CPU.
#include <string.h>

#include "miracl_pod.cuh"
#include "msm_impl.cuh"

namespace c12 {

enum { FLAG_BAD_POD = 4 };

template <class F> __global__ void __launch_bounds__(128) k_pod_points_to_wire(const uint8_t* __restrict__ pods, uint32_t n, uint8_t* __restrict__ wire, int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!pod_point_to_wire<F>(pods + (size_t)Pod<F>::POINT * i, wire + (size_t)Wire<F>::AFFINE * i)) atomicOr(flags, FLAG_BAD_POD);
}

template <class F> __global__ void __launch_bounds__(128) k_wire_to_pod_points(const uint8_t* __restrict__ wire, uint32_t n, uint8_t* __restrict__ pods, int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!wire_to_pod_point<F>(wire + (size_t)Wire<F>::AFFINE * i, pods + (size_t)Pod<F>::POINT * i)) atomicOr(flags, FLAG_BAD_POINT);
}

__global__ void __launch_bounds__(128) k_pod_bigs_to_scalars(const uint8_t* __restrict__ bigs, uint32_t n, uint8_t* __restrict__ out32, int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!pod_big_to_scalar(bigs + (size_t)POD_BIG * i, out32 + 32ull * i)) atomicOr(flags, FLAG_BAD_SCALAR);
}

__global__ void __launch_bounds__(64) k_pod_fp12_to_wire(const uint8_t* __restrict__ pods, uint32_t n, uint8_t* __restrict__ wire, int* flags)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 12) return;
    uint32_t e = i / 12, j = i % 12;
    if (!pod_fp12_coeff_to_wire(pods + (size_t)POD_FP12 * e, j, wire + 576ull * e)) atomicOr(flags, FLAG_BAD_POD);
}
__global__ void __launch_bounds__(64) k_wire_to_pod_fp12(const uint8_t* __restrict__ wire, uint32_t n, uint8_t* __restrict__ pods)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 12) return;
    uint32_t e = i / 12, j = i % 12;
    wire_to_pod_fp12_coeff(wire + 576ull * e, j, pods + (size_t)POD_FP12 * e);
}

// pairing kernels live in pairing.cu; reached through the _dev entries of the public ABI
} // namespace c12

using namespace c12;

namespace {

template <class F> int pods_to_wire(const uint8_t* d_pods, uint32_t n, uint8_t* d_wire, cudaStream_t s)
{
    k_pod_points_to_wire<F><<<cdiv(n, 128), 128, 0, s>>>(d_pods, n, d_wire, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}
template <class F> int wire_to_pods(const uint8_t* d_wire, uint32_t n, uint8_t* d_pods, cudaStream_t s)
{
    k_wire_to_pod_points<F><<<cdiv(n, 128), 128, 0, s>>>(d_wire, n, d_pods, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}
int bigs_to_scalars(const uint8_t* d_bigs, uint32_t n, uint8_t* d_out, cudaStream_t s)
{
    k_pod_bigs_to_scalars<<<cdiv(n, 128), 128, 0, s>>>(d_bigs, n, d_out, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}
int fp12_to_wire(const uint8_t* d_pods, uint32_t n, uint8_t* d_wire, cudaStream_t s)
{
    k_pod_fp12_to_wire<<<cdiv((size_t)n * 12, 64), 64, 0, s>>>(d_pods, n, d_wire, flags_word());
    C12_LAUNCHED();
    return C12381_OK;
}
int wire_to_fp12(const uint8_t* d_wire, uint32_t n, uint8_t* d_pods, cudaStream_t s)
{
    k_wire_to_pod_fp12<<<cdiv((size_t)n * 12, 64), 64, 0, s>>>(d_wire, n, d_pods);
    C12_LAUNCHED();
    return C12381_OK;
}

// result = sum_i numbers[i] * points[i] on PODs
template <class F> int pod_msm(void* result, int n, const void* points, const void* numbers)
{
    C12_REQUIRE_CTX();
    if (!result || n < 0 || (n && (!points || !numbers))) return set_error(C12381_EARG, "sum_of_products: bad argument");
    const void* in[2] = {points, numbers};
    size_t sz[2] = {(size_t)n * Pod<F>::POINT, (size_t)n * POD_BIG};
    size_t extra = align_up((size_t)n * Wire<F>::AFFINE + 4) + align_up((size_t)n * 32 + 4) + align_up(Wire<F>::AFFINE) + msm_scratch_for<F>(n);
    return with_staged(in, sz, 2, result, Pod<F>::POINT, extra, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        uint8_t* wire = (uint8_t*)arena_take((size_t)n * Wire<F>::AFFINE + 4);
        uint8_t* sc = (uint8_t*)arena_take((size_t)n * 32 + 4);
        uint8_t* res = (uint8_t*)arena_take(Wire<F>::AFFINE);
        int rc = C12381_OK;
        if (n) rc = pods_to_wire<F>(d_in[0], (uint32_t)n, wire, s);
        if (!rc && n) rc = bigs_to_scalars(d_in[1], (uint32_t)n, sc, s);
        if (!rc) rc = msm_run<F>(wire, sc, (size_t)n, res, OUT_AFFINE, s);
        if (!rc) rc = wire_to_pods<F>(res, 1, d_out, s);
        return rc;
    });
}

// object = value * object
template <class F> int pod_mul(void* object, const void* value)
{
    C12_REQUIRE_CTX();
    if (!object || !value) return set_error(C12381_EARG, "multiply: null pointer");
    const void* in[2] = {object, value};
    size_t sz[2] = {(size_t)Pod<F>::POINT, POD_BIG};
    // a one-term sum of products: the MSM pipeline's lane-cooperative tail beats one thread's double-and-add (entry_mul_dev)
    return with_staged(in, sz, 2, object, Pod<F>::POINT, 4096 + msm_scratch_for<F>(1), [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        uint8_t* wire = (uint8_t*)arena_take(Wire<F>::AFFINE);
        uint8_t* sc = (uint8_t*)arena_take(32);
        uint8_t* res = (uint8_t*)arena_take(Wire<F>::AFFINE);
        int rc = pods_to_wire<F>(d_in[0], 1, wire, s);
        if (!rc) rc = bigs_to_scalars(d_in[1], 1, sc, s);
        if (!rc) rc = msm_run<F>(wire, sc, 1, res, OUT_AFFINE, s);
        if (!rc) rc = wire_to_pods<F>(res, 1, d_out, s);
        return rc;
    });
}

// k-pair Miller product on PODs (k = 1: pair_ate, k = 2: pair_double_ate, up to C12381_MAX_PAIRS for the n-pairing entry)
int pod_miller(void* result_fp12, const void* const* p2s, const void* const* p1s, int k)
{
    C12_REQUIRE_CTX();
    if (k < 1 || k > C12381_MAX_PAIRS) return set_error(C12381_EARG, "pair_multi_ate: between 1 and C12381_MAX_PAIRS pairs");
    uint8_t h1[C12381_MAX_PAIRS * POD_P1], h2[C12381_MAX_PAIRS * POD_P2];
    for (int j = 0; j < k; ++j) {
        if (!p1s[j] || !p2s[j] || !result_fp12) return set_error(C12381_EARG, "pair_ate: null pointer");
        memcpy(h1 + (size_t)POD_P1 * j, p1s[j], POD_P1);
        memcpy(h2 + (size_t)POD_P2 * j, p2s[j], POD_P2);
    }
    const void* in[2] = {h1, h2};
    size_t sz[2] = {(size_t)POD_P1 * k, (size_t)POD_P2 * k};
    return with_staged(in, sz, 2, result_fp12, POD_FP12, 65536, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        uint8_t* w1 = (uint8_t*)arena_take(96 * C12381_MAX_PAIRS);
        uint8_t* w2 = (uint8_t*)arena_take(192 * C12381_MAX_PAIRS);
        uint8_t* f = (uint8_t*)arena_take(576);
        int rc = pods_to_wire<Fp>(d_in[0], (uint32_t)k, w1, s);
        if (!rc) rc = pods_to_wire<Fp2>(d_in[1], (uint32_t)k, w2, s);
        if (!rc) rc = c12381_miller_batch_dev(w1, w2, 1, k, f, s);
        if (!rc) rc = wire_to_fp12(f, 1, d_out, s);
        return rc;
    });
}

} // namespace

extern "C" {

int c12381_sum_of_products_miracl(void* result_point1, int n, const void* points_point1, const void* numbers_big)
{
    return pod_msm<Fp>(result_point1, n, points_point1, numbers_big);
}
int c12381_sum_of_products2_miracl(void* result_point2, int n, const void* points_point2, const void* numbers_big)
{
    return pod_msm<Fp2>(result_point2, n, points_point2, numbers_big);
}
int c12381_multiply_point1_miracl(void* object_point1, const void* value_big) { return pod_mul<Fp>(object_point1, value_big); }
int c12381_multiply_point2_miracl(void* object_point2, const void* value_big) { return pod_mul<Fp2>(object_point2, value_big); }

int c12381_double_multiply_miracl(void* p1, const void* p2, const void* v1, const void* v2)
{
    if (!p1 || !p2 || !v1 || !v2) return set_error(C12381_EARG, "double_multiply: null pointer");
    uint8_t pts[2 * POD_P1], nums[2 * POD_BIG];
    memcpy(pts, p1, POD_P1);
    memcpy(pts + POD_P1, p2, POD_P1);
    memcpy(nums, v1, POD_BIG);
    memcpy(nums + POD_BIG, v2, POD_BIG);
    return pod_msm<Fp>(p1, 2, pts, nums);
}

int c12381_pair_ate_miracl(void* result_fp12, const void* p2_point2, const void* p1_point1)
{
    const void* a2[1] = {p2_point2};
    const void* a1[1] = {p1_point1};
    return pod_miller(result_fp12, a2, a1, 1);
}
int c12381_pair_double_ate_miracl(void* result_fp12, const void* p2, const void* p1, const void* q2, const void* q1)
{
    const void* a2[2] = {p2, q2};
    const void* a1[2] = {p1, q1};
    return pod_miller(result_fp12, a2, a1, 2);
}

int c12381_pair_multi_ate_miracl(void* result_fp12, int n, const void* p2s_point2, const void* p1s_point1)
{
    if (n < 1 || n > C12381_MAX_PAIRS || !p2s_point2 || !p1s_point1) return set_error(C12381_EARG, "pair_multi_ate: between 1 and C12381_MAX_PAIRS pairs");
    const void *a2[C12381_MAX_PAIRS], *a1[C12381_MAX_PAIRS];
    for (int j = 0; j < n; ++j) {
        a2[j] = (const uint8_t*)p2s_point2 + (size_t)POD_P2 * j;
        a1[j] = (const uint8_t*)p1s_point1 + (size_t)POD_P1 * j;
    }
    return pod_miller(result_fp12, a2, a1, n);
}

int c12381_pair_final_exponentiation_miracl(void* object_fp12)
{
    C12_REQUIRE_CTX();
    if (!object_fp12) return set_error(C12381_EARG, "pair_final_exponentiation: null pointer");
    const void* in[1] = {object_fp12};
    size_t sz[1] = {POD_FP12};
    return with_staged(in, sz, 1, object_fp12, POD_FP12, 65536, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        uint8_t* f = (uint8_t*)arena_take(576);
        uint8_t* g = (uint8_t*)arena_take(576);
        int rc = fp12_to_wire(d_in[0], 1, f, s);
        if (!rc) rc = c12381_final_exp_batch_dev(f, 1, g, s);
        if (!rc) rc = wire_to_fp12(g, 1, d_out, s);
        return rc;
    });
}

int c12381_fp12_multiply_miracl(void* result_fp12, const void* value_fp12)
{
    C12_REQUIRE_CTX();
    if (!result_fp12 || !value_fp12) return set_error(C12381_EARG, "multiply(fp12): null pointer");
    const void* in[2] = {result_fp12, value_fp12};
    size_t sz[2] = {POD_FP12, POD_FP12};
    return with_staged(in, sz, 2, result_fp12, POD_FP12, 65536, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        uint8_t* a = (uint8_t*)arena_take(576);
        uint8_t* b = (uint8_t*)arena_take(576);
        uint8_t* g = (uint8_t*)arena_take(576);
        int rc = fp12_to_wire(d_in[0], 1, a, s);
        if (!rc) rc = fp12_to_wire(d_in[1], 1, b, s);
        if (!rc) rc = c12381_gt_mul_batch_dev(a, b, 1, g, s);
        if (!rc) rc = wire_to_fp12(g, 1, d_out, s);
        return rc;
    });
}

int c12381_fp12_pow_miracl(void* result_fp12, const void* base_fp12, const void* exponent_big)
{
    C12_REQUIRE_CTX();
    if (!result_fp12 || !base_fp12 || !exponent_big) return set_error(C12381_EARG, "pow(fp12): null pointer");
    const void* in[2] = {base_fp12, exponent_big};
    size_t sz[2] = {POD_FP12, POD_BIG};
    return with_staged(in, sz, 2, result_fp12, POD_FP12, 65536, [&](uint8_t** d_in, uint8_t* d_out, cudaStream_t s) {
        uint8_t* a = (uint8_t*)arena_take(576);
        uint8_t* e = (uint8_t*)arena_take(32);
        uint8_t* g = (uint8_t*)arena_take(576);
        int rc = fp12_to_wire(d_in[0], 1, a, s);
        if (!rc) rc = bigs_to_scalars(d_in[1], 1, e, s);
        if (!rc) rc = c12381_gt_pow_batch_dev(a, e, 1, g, s);
        if (!rc) rc = wire_to_fp12(g, 1, d_out, s);
        return rc;
    });
}

} // extern "C"
