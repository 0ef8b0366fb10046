/* c12381_cuda.h — C ABI of libc12381_cuda.so: the B200 (sm_100a) implementation of crypto12381's data-parallel
 * hot path.  These entry points are what a replacement of the reference bridge
 * (`crypto12381::detail::miracl_core`, include/crypto12381/miracl_core_interface.hpp, defined in
 * src/miracl_core_interface.cpp) binds to; each declaration cites the reference interface it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative C12381_E* code otherwise; nothing throws; the caller owns
 *     every buffer; there is NO CPU fallback: without a CUDA device every compute entry returns C12381_ENODEV.
 *   - "host" entries take host pointers and include the host<->device copies (on the context's own stream, and
 *     return when the result is in the caller's buffer); "_dev" entries take device pointers plus a cudaStream_t
 *     (passed as void*; NULL = CUDA's default stream) and enqueue work on exactly that stream.
 *   - canonical byte formats (the reference's own wire formats):
 *       scalar      32 B big-endian integer in [0, r)
 *       G1 affine   96 B  = x || y, 48 B big-endian each; identity = 96 zero bytes
 *       G2 affine   192 B = x.b || x.a || y.b || y.a (imaginary part first, FP2_toBytes order); identity = zeros
 *       G1 compressed 49 B = (0x02 | parity(y)) || x   (ECP_toOctet, ecp_BLS12381.cpp:445-491); identity = zeros
 *       G2 compressed 97 B = (0x02 | FP2_sign(y)) || x.b || x.a (ECP2_toOctet, ecp2_BLS12381.cpp:184-222)
 *       GT          576 B FP12_toOctet order (fp12_BLS12381.cpp:923-929)
 *   - "_miracl" entries take the reference's in-memory PODs (big = int64[7] 58-bit limbs; fp = big + int32 xes,
 *     Montgomery residue with R = 2^406; point1 192 B, point2 384 B, fp12 776 B) so the forwarding
 *     translation unit can pass its arguments through unchanged.
 *   - inputs are expected in the r-torsion subgroups (as produced by the reference), scalars reduced mod r.
 *   - one context per process (one process per GPU), one host thread at a time.  Every entry carves its scratch from the
 *     context's single arena; a call that arrives on a different stream than the previous one is ordered behind it on the
 *     device (cudaStreamWaitEvent on everything enqueued on the previous call's stream), so "_dev" calls on several
 *     streams and the host entries (the context's own stream, blocking) may be mixed freely - they serialise, they do not
 *     corrupt each other.  Malformed-input reports of the "_dev" entries and of the host entries use separate flag words.
 */
#ifndef C12381_CUDA_H
#define C12381_CUDA_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define C12381_API __attribute__((visibility("default")))
#else
#define C12381_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define C12381_OK 0
#define C12381_ENODEV (-1)   /* no CUDA device / context not initialised */
#define C12381_ECUDA (-2)    /* a CUDA runtime call failed (see c12381_last_error) */
#define C12381_EINPUT (-3)   /* malformed input: non-canonical coordinate, point off curve, scalar >= r */
#define C12381_EARG (-4)     /* bad argument (null pointer, k out of range, ...) */

#define C12381_MAX_PAIRS 8   /* pairs per pairing product instance */

/* ---- context ------------------------------------------------------------------------------------------- */
/* Bind this process to CUDA device `device` (one process per GPU) and create the stream + scratch pool. */
C12381_API int c12381_init(int device);
C12381_API void c12381_shutdown(void);
C12381_API const char* c12381_last_error(void);
C12381_API int c12381_device(void);              /* bound device or -1 */
/* The "_dev" entries only enqueue work, so malformed input (C12381_EINPUT) cannot be reported by their return value:
 * this call synchronises `stream` (NULL = the default stream) and returns C12381_EINPUT if any kernel enqueued since
 * the last check flagged a non-canonical coordinate, an off-curve point or a scalar >= r (then clears the flag).
 * The host-pointer entries do this themselves.  The reference reports parse failures the same way: by status
 * (from_bytes returns 0, src/miracl_core_interface.cpp:109-112), never by a different result. */
C12381_API int c12381_sync_status(void* stream);
/* testing/tuning knob: force the MSM window width (0 = automatic) */
C12381_API void c12381_set_msm_window(int c);
/* testing/tuning knob: batch-affine halving rounds of the bucket lists in front of the XYZZ accumulation
 * (-1 = chosen from the bucket load, the default; 0 = none; k = k rounds; same results) */
C12381_API void c12381_set_msm_batch_affine(int rounds);
/* testing/tuning knob: independent pipelines (window groups on their own streams) the halving rounds are split into (1 .. 4) */
C12381_API void c12381_set_msm_pipelines(int pipes);
/* measurement knobs of the halving rounds (A/B runs; same results): id 0 = waves of resident warps a pipeline's round should
 * span (slots per lane J follows), 1 = largest J, 2 = halvings left to the XYZZ accumulation by the automatic round count,
 * 3 = threads the segment running sums of the bucket reduction should fill (sets the segment length; 0 = default),
 * 4 = upload groups of the host-pointer MSM entries (1 .. 8; four measured best, profiles/r03n: the points are uploaded in that many pieces, each in front of its
 * own pipeline of halving rounds; default 4),
 * 5 = how the bucket lists are made: 0 (default) by counting - atomic ranks, one scan, one scatter - or 1 by the stable segmented
 * radix sort and a bounds search (same results; the order inside a bucket's list is irrelevant to its sum),
 * 6 = device-resident single-group calls: 1 (default) runs the scalar-only stages on a high-priority stream beside the parse of
 * the points, 0 runs everything in line,
 * 7 = the tail (bucket reduction + Horner combination): 0 = one chain over all windows; 1 = early split (two pipelines of halving
 * rounds, the high windows the smaller one, their whole tail under the low windows' last rounds); 2 (default) = late split (one
 * accumulation, then the high half of the windows reduced and combined on a side stream beside the low half),
 * 8 = percent of the resident warps (BA kernels' blocks per SM x 4 warps x SMs) one lane's halving round is sized for (default 100) */
C12381_API void c12381_set_knob(int id, int value);

/* ---- multi-scalar multiplication ------------------------------------------------------------------------- */
/* out = sum_i scalars[i] * points[i] over G1.
 * Replaces sum_of_products(point1&, int, point1*, const big*) -> ECP_muln
 * (miracl_core_interface.hpp:102; src/miracl_core_interface.cpp:134-137) and the live DSL MSM loop
 * product(type_identity<G1Pow>, range) of double_multiply calls (g1_point.hpp:371-404). */
C12381_API int c12381_g1_msm(const uint8_t* points96, const uint8_t* scalars32, size_t n, uint8_t out49[49]);
C12381_API int c12381_g1_msm_dev(const uint8_t* d_points96, const uint8_t* d_scalars32, size_t n, uint8_t* d_out49, void* stream);
/* same sum, result as AFFINE 96 B (the per-rank partial of a sharded MSM); host-pointer and device-pointer forms */
C12381_API int c12381_g1_msm_partial(const uint8_t* points96, const uint8_t* scalars32, size_t n, uint8_t out96[96]);
C12381_API int c12381_g1_msm_partial_dev(const uint8_t* d_points96, const uint8_t* d_scalars32, size_t n, uint8_t* d_out96, void* stream);
/* out = sum_i points[i] (combining the all-gathered per-rank partials in rank order); compressed result.
 * Replaces the add(point1&, point1&) loop (src/miracl_core_interface.cpp:129-132) used to merge partials. */
C12381_API int c12381_g1_sum_dev(const uint8_t* d_points96, size_t n, uint8_t* d_out49, void* stream);

/* out = sum_i scalars[i] * points[i] over G2.  The reference has no G2 MSM entry: it replaces the per-term
 * multiply(point2&, const big&) -> PAIR_G2mul + add(point2&, point2&) loop (g2_point.hpp:202-236;
 * src/miracl_core_interface.cpp:202-205,212-215). */
C12381_API int c12381_g2_msm(const uint8_t* points192, const uint8_t* scalars32, size_t n, uint8_t out97[97]);
C12381_API int c12381_g2_msm_dev(const uint8_t* d_points192, const uint8_t* d_scalars32, size_t n, uint8_t* d_out97, void* stream);
C12381_API int c12381_g2_msm_partial(const uint8_t* points192, const uint8_t* scalars32, size_t n, uint8_t out192[192]);
C12381_API int c12381_g2_msm_partial_dev(const uint8_t* d_points192, const uint8_t* d_scalars32, size_t n, uint8_t* d_out192, void* stream);
C12381_API int c12381_g2_sum_dev(const uint8_t* d_points192, size_t n, uint8_t* d_out97, void* stream);

/* ---- batched scalar multiplication ------------------------------------------------------------------------ */
/* out[i] = scalars[i] * points[i], compressed.  Replaces multiply(point1&, const big&) -> PAIR_G1mul
 * (src/miracl_core_interface.cpp:174-177) / multiply(point2&, const big&) -> PAIR_G2mul (:202-205). */
C12381_API int c12381_g1_mul_batch(const uint8_t* points96, const uint8_t* scalars32, size_t n, uint8_t* out49);
C12381_API int c12381_g2_mul_batch(const uint8_t* points192, const uint8_t* scalars32, size_t n, uint8_t* out97);
C12381_API int c12381_g1_mul_batch_dev(const uint8_t* d_points96, const uint8_t* d_scalars32, size_t n, uint8_t* d_out49, void* stream);
C12381_API int c12381_g2_mul_batch_dev(const uint8_t* d_points192, const uint8_t* d_scalars32, size_t n, uint8_t* d_out97, void* stream);
/* out[i] = scalars[i] * generator, AFFINE.  Replaces the `select` path g^x: get_default_generator + multiply
 * (g1_point.hpp:355-369, g2_point.hpp:129-143; src/miracl_core_interface.cpp:169-177,233-236). */
C12381_API int c12381_g1_fixed_base_mul_batch(const uint8_t* scalars32, size_t n, uint8_t* out96);
C12381_API int c12381_g2_fixed_base_mul_batch(const uint8_t* scalars32, size_t n, uint8_t* out192);
C12381_API int c12381_g1_fixed_base_mul_batch_dev(const uint8_t* d_scalars32, size_t n, uint8_t* d_out96, void* stream);
C12381_API int c12381_g2_fixed_base_mul_batch_dev(const uint8_t* d_scalars32, size_t n, uint8_t* d_out192, void* stream);

/* out[b] = sum_j scalars[b * m + j] * bases[j], AFFINE: B independent small sums over m bases shared by all instances
 * (window tables for the bases are built inside the call).  This is the shape of the per-signature products in the
 * examples - g1 * h0^r * Π[n](h[i]^m[i]) over the public parameters (examples/bbs-plus/src/bbs+.cpp:53,72; ps.cpp:81,98) -
 * when many signatures are processed at once; it replaces one chain of double_multiply / multiply / add calls
 * (src/miracl_core_interface.cpp:129-132,174-182,202-215) per instance. */
C12381_API int c12381_g1_multi_fixed_base_batch(const uint8_t* bases96, size_t m, const uint8_t* scalars32, size_t B, uint8_t* out96);
C12381_API int c12381_g2_multi_fixed_base_batch(const uint8_t* bases192, size_t m, const uint8_t* scalars32, size_t B, uint8_t* out192);
C12381_API int c12381_g1_multi_fixed_base_batch_dev(const uint8_t* d_bases96, size_t m, const uint8_t* d_scalars32, size_t B, uint8_t* d_out96, void* stream);
C12381_API int c12381_g2_multi_fixed_base_batch_dev(const uint8_t* d_bases192, size_t m, const uint8_t* d_scalars32, size_t B, uint8_t* d_out192, void* stream);

/* ---- wire-format conversions in batch ------------------------------------------------------------- */
/* decompress: 49 / 97-byte encodings -> affine 96 / 192 B.  Replaces from_bytes(point1&, bytes_view&) -> ECP_fromOctet and
 * from_bytes(point2&, ...) -> ECP2_fromOctet (src/miracl_core_interface.cpp:109-112,187-190): y = sqrt(x^3 + b) with the
 * sign from the tag byte; all-zero input = identity (g1_point.hpp:87-106); a bad tag, x >= p or x not on the curve is
 * reported as C12381_EINPUT (the reference returns status 0).  compress: the inverse (to_bytes, :114-117,192-195). */
C12381_API int c12381_g1_decompress_batch(const uint8_t* in49, size_t n, uint8_t* out96);
C12381_API int c12381_g2_decompress_batch(const uint8_t* in97, size_t n, uint8_t* out192);
C12381_API int c12381_g1_compress_batch(const uint8_t* in96, size_t n, uint8_t* out49);
C12381_API int c12381_g2_compress_batch(const uint8_t* in192, size_t n, uint8_t* out97);
C12381_API int c12381_g1_decompress_batch_dev(const uint8_t* d_in49, size_t n, uint8_t* d_out96, void* stream);
C12381_API int c12381_g2_decompress_batch_dev(const uint8_t* d_in97, size_t n, uint8_t* d_out192, void* stream);
C12381_API int c12381_g1_compress_batch_dev(const uint8_t* d_in96, size_t n, uint8_t* d_out49, void* stream);
C12381_API int c12381_g2_compress_batch_dev(const uint8_t* d_in192, size_t n, uint8_t* d_out97, void* stream);

/* ---- subgroup membership --------------------------------------------------------------------------- */
/* verdicts[i] = 1 iff the (on-curve) point i lies in the r-torsion subgroup.  Same test and conventions as the reference's
 * unbridged PAIR_G1member / PAIR_G2member (3rd-party/miracl-core/pair_BLS12381.cpp:1034-1130): the endomorphism test, the
 * identity is NOT a member.  Off-curve / non-canonical input is C12381_EINPUT. */
C12381_API int c12381_g1_subgroup_check_batch(const uint8_t* points96, size_t n, uint8_t* verdicts);
C12381_API int c12381_g2_subgroup_check_batch(const uint8_t* points192, size_t n, uint8_t* verdicts);
C12381_API int c12381_g1_subgroup_check_batch_dev(const uint8_t* d_points96, size_t n, uint8_t* d_verdicts, void* stream);
C12381_API int c12381_g2_subgroup_check_batch_dev(const uint8_t* d_points192, size_t n, uint8_t* d_verdicts, void* stream);

/* ---- Fiat-Shamir hashing in batch -------------------------------------------------------------------- */
/* B messages of msg_len bytes each, back to back.  sha3_512: the 64-byte SHA3-512 digests - what hash_state yields for the
 * serialised bytes it absorbed (include/crypto12381/set.hpp:317-392 -> sha3_init / sha3_process / sha3_hash,
 * src/miracl_core_interface.cpp:12-25 -> 3rd-party/miracl-core/hash.cpp:480-554).  hash_to_zp: that digest read as a
 * big-endian 512-bit integer reduced mod r, 32 B big-endian (Zp's from_hash, zp_number.hpp:538-547). */
C12381_API int c12381_sha3_512_batch(const uint8_t* msgs, size_t msg_len, size_t B, uint8_t* out64);
C12381_API int c12381_hash_to_zp_batch(const uint8_t* msgs, size_t msg_len, size_t B, uint8_t* out32);
C12381_API int c12381_sha3_512_batch_dev(const uint8_t* d_msgs, size_t msg_len, size_t B, uint8_t* d_out64, void* stream);
C12381_API int c12381_hash_to_zp_batch_dev(const uint8_t* d_msgs, size_t msg_len, size_t B, uint8_t* d_out32, void* stream);
/* hash_to_g1: what `hash(...) -> G1` yields for each message - G1Point::from_hash (include/crypto12381/g1_point.hpp:219-234): the
 * SHA3-512 digest reduced mod p (from_bytes(big2&) + fixed_time_mod, src/miracl_core_interface.cpp:91-99), map_to_point (:154-157 ->
 * ECP_map2point, 3rd-party/miracl-core/ecp_BLS12381.cpp:1276,1493-1627: simplified SWU on the 11-isogenous curve, Z = 11, then the
 * isogeny) and multiply_cofactor (:159-162 -> ECP_cfp, times 1 - x).  49 B compressed per message.  map_to_g1: the last two steps for
 * B field elements given as 48 B big-endian integers < p (else C12381_EINPUT); the two inputs with Z^2 u^4 + Z u^2 = 0 give the
 * identity, as in the reference. */
C12381_API int c12381_hash_to_g1_batch(const uint8_t* msgs, size_t msg_len, size_t B, uint8_t* out49);
C12381_API int c12381_map_to_g1_batch(const uint8_t* u48, size_t B, uint8_t* out49);
C12381_API int c12381_hash_to_g1_batch_dev(const uint8_t* d_msgs, size_t msg_len, size_t B, uint8_t* d_out49, void* stream);
C12381_API int c12381_map_to_g1_batch_dev(const uint8_t* d_u48, size_t B, uint8_t* d_out49, void* stream);

/* ---- pairings --------------------------------------------------------------------------------------------- */
/* B instances, k pairs each (1 <= k <= C12381_MAX_PAIRS): g1s = B*k*96 B, g2s = B*k*192 B, instance-major.
 * miller: out[b] = conj-adjusted product of Miller loops, NOT exponentiated (576 B raw Fp12).
 *   Replaces pair_ate -> PAIR_ate (src/miracl_core_interface.cpp:276-279), pair_double_ate -> PAIR_double_ate
 *   (:286-289) and their products via multiply(fp12&, fp12&) (:256-259; liner_pair.hpp:219-230,291-303).
 * final_exp: out[b] = in[b]^(3 (p^12-1)/r).  Replaces pair_final_exponentiation -> PAIR_fexp (:281-284).
 * pairing_product = final_exp(miller).  pairing_check: verdict[b] = 1 iff the product is the GT identity
 *   (operator== on pair/Miller operands, liner_pair.hpp:336-357). */
C12381_API int c12381_miller_batch(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, uint8_t* out576);
C12381_API int c12381_final_exp_batch(const uint8_t* in576, size_t B, uint8_t* out576);
C12381_API int c12381_pairing_product_batch(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, uint8_t* out576);
C12381_API int c12381_pairing_check_batch(const uint8_t* g1s, const uint8_t* g2s, size_t B, int k, uint8_t* verdicts);
C12381_API int c12381_miller_batch_dev(const uint8_t* d_g1s, const uint8_t* d_g2s, size_t B, int k, uint8_t* d_out576, void* stream);
C12381_API int c12381_final_exp_batch_dev(const uint8_t* d_in576, size_t B, uint8_t* d_out576, void* stream);
C12381_API int c12381_pairing_product_batch_dev(const uint8_t* d_g1s, const uint8_t* d_g2s, size_t B, int k, uint8_t* d_out576, void* stream);
C12381_API int c12381_pairing_check_batch_dev(const uint8_t* d_g1s, const uint8_t* d_g2s, size_t B, int k, uint8_t* d_verdicts, void* stream);

/* testing/tuning knob: which kernels serve the pairing / GT entries.  0 = automatic (by batch size), 1 = one thread per
 * instance, 2 = six cooperating lanes per instance with the state in shared memory.  Same results either way. */
C12381_API void c12381_set_pairing_kernel(int mode);

/* ---- GT helpers ------------------------------------------------------------------------------------------- */
/* out[b] = a[b] * b[b].  Replaces multiply(fp12&, fp12&) -> FP12_mul (src/miracl_core_interface.cpp:256-259). */
C12381_API int c12381_gt_mul_batch(const uint8_t* a576, const uint8_t* b576, size_t B, uint8_t* out576);
/* out[b] = a[b]^scalars[b] for unitary a.  Replaces pow(fp12&, fp12&, const big&) -> FP12_pow (:261-264). */
C12381_API int c12381_gt_pow_batch(const uint8_t* a576, const uint8_t* scalars32, size_t B, uint8_t* out576);
C12381_API int c12381_gt_mul_batch_dev(const uint8_t* d_a576, const uint8_t* d_b576, size_t B, uint8_t* d_out576, void* stream);
C12381_API int c12381_gt_pow_batch_dev(const uint8_t* d_a576, const uint8_t* d_scalars32, size_t B, uint8_t* d_out576, void* stream);
/* out[b] = a[b]^scalars[b] for a in GT (order r: pairing values and their products / powers) through the Galbraith-Scott
 * split of the exponent - the value of the reference's unbridged PAIR_GTpow built with USE_GS_GT
 * (3rd-party/miracl-core/pair_BLS12381.cpp:985-1026): four 63-bit exponents over a, a^p, a^(p^2), a^(p^3), 62 cyclotomic
 * squarings instead of 254.  Equal to c12381_gt_pow_batch on GT; NOT valid for other unitary elements. */
C12381_API int c12381_gt_pow_gs_batch(const uint8_t* a576, const uint8_t* scalars32, size_t B, uint8_t* out576);
C12381_API int c12381_gt_pow_gs_batch_dev(const uint8_t* d_a576, const uint8_t* d_scalars32, size_t B, uint8_t* d_out576, void* stream);

/* ---- drop-in entries on the reference's PODs (batch of 1 per call; host pointers) --------------------------- */
/* void sum_of_products(point1& result, int n, point1* points, const big* numbers)  (miracl_core_interface.hpp:102) */
C12381_API int c12381_sum_of_products_miracl(void* result_point1, int n, const void* points_point1, const void* numbers_big);
/* void multiply(point1& object, const big& value)  (:122) */
C12381_API int c12381_multiply_point1_miracl(void* object_point1, const void* value_big);
/* void double_multiply(point1& p1, point1& p2, big& v1, big& v2): p1 = v1*p1 + v2*p2  (:125) */
C12381_API int c12381_double_multiply_miracl(void* p1_point1, const void* p2_point1, const void* v1_big, const void* v2_big);
/* void multiply(point2& object, const big& value)  (:151; src/miracl_core_interface.cpp:202) */
C12381_API int c12381_multiply_point2_miracl(void* object_point2, const void* value_big);
/* batched G2 sum of products on PODs (new entry the lazy G2Pow of SURVEY "next" N2 would call) */
C12381_API int c12381_sum_of_products2_miracl(void* result_point2, int n, const void* points_point2, const void* numbers_big);
/* void pair_ate(fp12& result, point2& p2, point1& p1)  (:200) */
C12381_API int c12381_pair_ate_miracl(void* result_fp12, const void* p2_point2, const void* p1_point1);
/* void pair_double_ate(fp12& result, point2& p2, point1& p1, point2& q2, point1& q1)  (:204) */
C12381_API int c12381_pair_double_ate_miracl(void* result_fp12, const void* p2, const void* p1, const void* q2, const void* q1);
/* ABI-additive: the product of n <= C12381_MAX_PAIRS Miller loops with shared squarings, p2s = point2[n], p1s = point1[n];
 * what a `pair * pair * ...` chain (liner_pair.hpp:219-230,291-303: pair_double_ate two at a time + multiply(fp12&, fp12&)) folds to,
 * and the value of MIRACL's unbridged PAIR_initmp / PAIR_another / PAIR_miller (pair_BLS12381.cpp:181-207,352-422). */
C12381_API int c12381_pair_multi_ate_miracl(void* result_fp12, int n, const void* p2s_point2, const void* p1s_point1);
/* void pair_final_exponentiation(fp12& object)  (:202) */
C12381_API int c12381_pair_final_exponentiation_miracl(void* object_fp12);
/* void multiply(fp12& result, fp12& value): result *= value  (:192) */
C12381_API int c12381_fp12_multiply_miracl(void* result_fp12, const void* value_fp12);
/* void pow(fp12& result, fp12& base, const big& exponent)  (:194) */
C12381_API int c12381_fp12_pow_miracl(void* result_fp12, const void* base_fp12, const void* exponent_big);

/* ---- measurement helpers ---------------------------------------------------------------------------------- */
/* Integer-multiply roofline probes: run `iters` dependent-chain iterations of the named
 * instruction mix on every SM and return achieved giga-ops/s in *out_gops (ops as documented per kind):
 *   kind 0: mad.lo.u32 (IMAD)            ops = 32-bit multiply-adds
 *   kind 1: mad.lo.cc / madc.hi.cc pairs ops = 32-bit multiply-add instructions
 *   kind 2: mad.wide.u32 (IMAD.WIDE)     ops = 32x32->64 multiply-adds, every chain multiplying a value of its own (until round 2's
 *           last visit all chains shared their operands and ptxas hoisted the product: that rate was not a multiplier rate)
 *   kind 3: full Fp Montgomery product   ops = Fp multiplications
 *   kind 4: full Fp Montgomery squaring  ops = Fp squarings */
C12381_API int c12381_probe(int kind, int iters, double* out_gops, double* out_ms);
/* number of kernels launched by this library since init (bench.py's gpu_launches) */
C12381_API unsigned long long c12381_launch_count(void);
/* CUDA-event time (ms) of the dominant MSM kernel (bucket accumulation) in the most recent MSM call, and the
 * number of bucket additions it performed */
C12381_API int c12381_last_msm_stats(double* accumulate_ms, double* total_ms, unsigned long long* bucket_adds, int* window_bits);
/* CUDA-event times (ms) of the eight phases of the most recent MSM call: recode, sort, bucket bounds + order, parse
 * (includes waiting for the point upload in the host entry), accumulate, reduction levels, reduce-2, finish */
C12381_API int c12381_last_msm_phases(double* phase_ms8);
/* how the last MSM on this context formed its bucket sums: batch-affine halving rounds (0 = XYZZ additions only), the pipelines
 * they ran in, the upload groups of a host-pointer call */
C12381_API int c12381_last_msm_shape(int* ba_rounds, int* ba_pipelines, int* upload_groups);

#ifdef __cplusplus
}
#endif
#endif
