// G2 instantiation of the MSM / scalar-multiplication pipeline and its C-ABI entries (include/c12381_cuda.h).
// In this translation unit the Montgomery products inside an Fp2 product are inlined so that ptxas interleaves their carry
// chains: measured +9 % on the G2 bucket accumulation (profiles/r01ab_fp2_inline_ab.txt); the register-capped cooperative
// pairing kernels lose 20 % with the same setting, so pairing.cu keeps the calls.
#define C12_FP2_INLINE_MULS 1
#include "msm_impl.cuh"
using namespace c12;

extern "C" {
int c12381_g2_msm(const uint8_t* points192, const uint8_t* scalars32, size_t n, uint8_t out97[97]) { return entry_msm_host<Fp2>(points192, scalars32, n, out97); }
int c12381_g2_msm_partial(const uint8_t* p, const uint8_t* s, size_t n, uint8_t* o) { return entry_msm_host<Fp2>(p, s, n, o, OUT_AFFINE); }
int c12381_g2_msm_dev(const uint8_t* p, const uint8_t* s, size_t n, uint8_t* o, void* st) { return entry_msm_dev<Fp2>(p, s, n, o, OUT_COMPRESSED, st); }
int c12381_g2_msm_partial_dev(const uint8_t* p, const uint8_t* s, size_t n, uint8_t* o, void* st) { return entry_msm_dev<Fp2>(p, s, n, o, OUT_AFFINE, st); }
int c12381_g2_sum_dev(const uint8_t* p, size_t n, uint8_t* o, void* st) { return entry_sum_dev<Fp2>(p, n, o, st); }
int c12381_g2_mul_batch(const uint8_t* p, const uint8_t* s, size_t n, uint8_t* o) { return entry_mul_host<Fp2>(p, s, n, o); }
int c12381_g2_mul_batch_dev(const uint8_t* p, const uint8_t* s, size_t n, uint8_t* o, void* st) { return entry_mul_dev<Fp2>(p, s, n, o, st); }
int c12381_g2_fixed_base_mul_batch(const uint8_t* s, size_t n, uint8_t* o) { return entry_fixed_host<Fp2>(s, n, o); }
int c12381_g2_fixed_base_mul_batch_dev(const uint8_t* s, size_t n, uint8_t* o, void* st) { return entry_fixed_dev<Fp2>(s, n, o, st); }
int c12381_g2_multi_fixed_base_batch(const uint8_t* bases, size_t m, const uint8_t* s, size_t B, uint8_t* o) { return entry_multi_fixed_host<Fp2>(bases, m, s, B, o); }
int c12381_g2_multi_fixed_base_batch_dev(const uint8_t* bases, size_t m, const uint8_t* s, size_t B, uint8_t* o, void* st) { return entry_multi_fixed_dev<Fp2>(bases, m, s, B, o, st); }
int c12381_g2_decompress_batch(const uint8_t* in, size_t n, uint8_t* o) { return entry_convert_host<Fp2>(in, n, o, true); }
int c12381_g2_decompress_batch_dev(const uint8_t* in, size_t n, uint8_t* o, void* st) { return entry_convert_dev<Fp2>(in, n, o, true, st); }
int c12381_g2_compress_batch(const uint8_t* in, size_t n, uint8_t* o) { return entry_convert_host<Fp2>(in, n, o, false); }
int c12381_g2_compress_batch_dev(const uint8_t* in, size_t n, uint8_t* o, void* st) { return entry_convert_dev<Fp2>(in, n, o, false, st); }
int c12381_g2_subgroup_check_batch(const uint8_t* in, size_t n, uint8_t* o) { return entry_subgroup_host<Fp2>(in, n, o); }
int c12381_g2_subgroup_check_batch_dev(const uint8_t* in, size_t n, uint8_t* o, void* st) { return entry_subgroup_dev<Fp2>(in, n, o, st); }
}
