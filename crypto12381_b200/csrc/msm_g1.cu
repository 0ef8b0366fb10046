// G1 instantiation of the MSM / scalar-multiplication pipeline and its C-ABI entries (include/c12381_cuda.h).
#include "msm_impl.cuh"
using namespace c12;

extern "C" {
int c12381_g1_msm(const uint8_t* points96, const uint8_t* scalars32, size_t n, uint8_t out49[49]) { return entry_msm_host<Fp>(points96, scalars32, n, out49); }
int c12381_g1_msm_partial(const uint8_t* p, const uint8_t* s, size_t n, uint8_t* o) { return entry_msm_host<Fp>(p, s, n, o, OUT_AFFINE); }
int c12381_g1_msm_dev(const uint8_t* p, const uint8_t* s, size_t n, uint8_t* o, void* st) { return entry_msm_dev<Fp>(p, s, n, o, OUT_COMPRESSED, st); }
int c12381_g1_msm_partial_dev(const uint8_t* p, const uint8_t* s, size_t n, uint8_t* o, void* st) { return entry_msm_dev<Fp>(p, s, n, o, OUT_AFFINE, st); }
int c12381_g1_sum_dev(const uint8_t* p, size_t n, uint8_t* o, void* st) { return entry_sum_dev<Fp>(p, n, o, st); }
int c12381_g1_mul_batch(const uint8_t* p, const uint8_t* s, size_t n, uint8_t* o) { return entry_mul_host<Fp>(p, s, n, o); }
int c12381_g1_mul_batch_dev(const uint8_t* p, const uint8_t* s, size_t n, uint8_t* o, void* st) { return entry_mul_dev<Fp>(p, s, n, o, st); }
int c12381_g1_fixed_base_mul_batch(const uint8_t* s, size_t n, uint8_t* o) { return entry_fixed_host<Fp>(s, n, o); }
int c12381_g1_fixed_base_mul_batch_dev(const uint8_t* s, size_t n, uint8_t* o, void* st) { return entry_fixed_dev<Fp>(s, n, o, st); }
int c12381_g1_multi_fixed_base_batch(const uint8_t* bases, size_t m, const uint8_t* s, size_t B, uint8_t* o) { return entry_multi_fixed_host<Fp>(bases, m, s, B, o); }
int c12381_g1_multi_fixed_base_batch_dev(const uint8_t* bases, size_t m, const uint8_t* s, size_t B, uint8_t* o, void* st) { return entry_multi_fixed_dev<Fp>(bases, m, s, B, o, st); }
int c12381_g1_decompress_batch(const uint8_t* in, size_t n, uint8_t* o) { return entry_convert_host<Fp>(in, n, o, true); }
int c12381_g1_decompress_batch_dev(const uint8_t* in, size_t n, uint8_t* o, void* st) { return entry_convert_dev<Fp>(in, n, o, true, st); }
int c12381_g1_compress_batch(const uint8_t* in, size_t n, uint8_t* o) { return entry_convert_host<Fp>(in, n, o, false); }
int c12381_g1_compress_batch_dev(const uint8_t* in, size_t n, uint8_t* o, void* st) { return entry_convert_dev<Fp>(in, n, o, false, st); }
int c12381_g1_subgroup_check_batch(const uint8_t* in, size_t n, uint8_t* o) { return entry_subgroup_host<Fp>(in, n, o); }
int c12381_g1_subgroup_check_batch_dev(const uint8_t* in, size_t n, uint8_t* o, void* st) { return entry_subgroup_dev<Fp>(in, n, o, st); }
}
