"""Host-side mirror of the reference bridge `crypto12381::detail::miracl_core` for the hot path
(reference: include/crypto12381/miracl_core_interface.hpp:143-205, src/miracl_core_interface.cpp:134-289).

Same names, argument meaning and error behaviour as the reference functions they stand for; values travel in the
reference's own wire formats instead of MIRACL structs:

    scalar 32 B big-endian < r | G1 affine 96 B (x||y) | G2 affine 192 B (x.b||x.a||y.b||y.a) | identity = zeros
    G1 compressed 49 B | G2 compressed 97 B | GT 576 B

Every call runs on the GPU through libc12381_cuda.so (host pointers in, host pointers out; the copies are part
of the call).  Malformed input raises C12381Error(EINPUT) — the reference reports it by status too
(from_bytes -> 0).  The `*_pod` variants take the reference's in-memory structs (ctypes buffers holding
point1 / point2 / fp12 / big) exactly as the forwarding translation unit of INTEGRATION.md passes them."""
from __future__ import annotations

import ctypes

from . import _lib
from ._lib import check, ensure_init, lib

G1_AFFINE, G2_AFFINE, G1_COMPRESSED, G2_COMPRESSED, GT_BYTES, SCALAR = 96, 192, 49, 97, 576, 32


def _count(buf: bytes, size: int, what: str) -> int:
    if len(buf) % size:
        raise ValueError(f"{what}: length {len(buf)} is not a multiple of {size}")
    return len(buf) // size


def _out(n: int):
    return ctypes.create_string_buffer(max(n, 1))


# ---- sum_of_products(point1& result, int n, point1* points, const big* numbers) -> ECP_muln (:134-137) ----------
def sum_of_products(points: bytes, numbers: bytes) -> bytes:
    """Σ numbers[i]·points[i] over G1 (the DSL's Π[n](h[i]^m[i]), g1_point.hpp:371-404); 49-byte compressed result."""
    ensure_init()
    n = _count(numbers, SCALAR, "numbers")
    if _count(points, G1_AFFINE, "points") != n:
        raise ValueError("sum_of_products: points and numbers differ in length")
    out = _out(G1_COMPRESSED)
    check(lib().c12381_g1_msm(points, numbers, n, out))
    return out.raw[:G1_COMPRESSED]


def sum_of_products2(points: bytes, numbers: bytes) -> bytes:
    """Σ numbers[i]·points[i] over G2 (replaces the per-term PAIR_G2mul + ECP2_add loop, g2_point.hpp:202-236)."""
    ensure_init()
    n = _count(numbers, SCALAR, "numbers")
    if _count(points, G2_AFFINE, "points") != n:
        raise ValueError("sum_of_products2: points and numbers differ in length")
    out = _out(G2_COMPRESSED)
    check(lib().c12381_g2_msm(points, numbers, n, out))
    return out.raw[:G2_COMPRESSED]


# ---- multiply(point1& object, const big& value) -> PAIR_G1mul (:174-177), batched -----------------------------------
def multiply(points: bytes, values: bytes) -> bytes:
    """out[i] = values[i]·points[i]; G1 when points are 96-byte records. Compressed results, concatenated."""
    ensure_init()
    n = _count(values, SCALAR, "values")
    out = _out(G1_COMPRESSED * n)
    if _count(points, G1_AFFINE, "points") != n:
        raise ValueError("multiply: points and values differ in length")
    check(lib().c12381_g1_mul_batch(points, values, n, out))
    return out.raw[:G1_COMPRESSED * n]


def multiply2(points: bytes, values: bytes) -> bytes:
    """multiply(point2&, const big&) -> PAIR_G2mul (:202-205), batched."""
    ensure_init()
    n = _count(values, SCALAR, "values")
    if _count(points, G2_AFFINE, "points") != n:
        raise ValueError("multiply2: points and values differ in length")
    out = _out(G2_COMPRESSED * n)
    check(lib().c12381_g2_mul_batch(points, values, n, out))
    return out.raw[:G2_COMPRESSED * n]


def double_multiply(p1: bytes, p2: bytes, v1: bytes, v2: bytes) -> bytes:
    """v1·p1 + v2·p2 (double_multiply -> ECP_mul2, :179-182): a two-term sum of products."""
    return sum_of_products(p1 + p2, v1 + v2)


def generator_power(values: bytes) -> bytes:
    """g^x for the default G1 generator (select path, g1_point.hpp:355-369): affine 96-byte results."""
    ensure_init()
    n = _count(values, SCALAR, "values")
    out = _out(G1_AFFINE * n)
    check(lib().c12381_g1_fixed_base_mul_batch(values, n, out))
    return out.raw[:G1_AFFINE * n]


def generator_power2(values: bytes) -> bytes:
    """g2^x for the default G2 generator (g2_point.hpp:129-143): affine 192-byte results."""
    ensure_init()
    n = _count(values, SCALAR, "values")
    out = _out(G2_AFFINE * n)
    check(lib().c12381_g2_fixed_base_mul_batch(values, n, out))
    return out.raw[:G2_AFFINE * n]


def products_over_bases(bases: bytes, values: bytes) -> bytes:
    """out[b] = Σ_j values[b][j]·bases[j] over G1 for m shared bases: the per-signature products of the examples,
    g1 * h0^r * Π[n](h[i]^m[i]) (examples/bbs-plus/src/bbs+.cpp:53,72), for many signatures at once.  Affine 96 B each."""
    ensure_init()
    m = _count(bases, G1_AFFINE, "bases")
    B = _count(values, SCALAR * m, "values") if m else 0
    out = _out(G1_AFFINE * B)
    check(lib().c12381_g1_multi_fixed_base_batch(bases, m, values, B, out))
    return out.raw[:G1_AFFINE * B]


def products_over_bases2(bases: bytes, values: bytes) -> bytes:
    """Same over G2 (e.g. w * g2^x, bbs+.cpp:72): affine 192 B each."""
    ensure_init()
    m = _count(bases, G2_AFFINE, "bases")
    B = _count(values, SCALAR * m, "values") if m else 0
    out = _out(G2_AFFINE * B)
    check(lib().c12381_g2_multi_fixed_base_batch(bases, m, values, B, out))
    return out.raw[:G2_AFFINE * B]


# ---- from_bytes / to_bytes (point1, point2) -> ECP_fromOctet / ECP_toOctet, ECP2_* (:109-117,187-195), batched ----------
def from_bytes(encoded: bytes) -> bytes:
    """49-byte G1 encodings -> affine 96 B; raises C12381Error(EINPUT) where the reference's from_bytes returns 0."""
    ensure_init()
    n = _count(encoded, G1_COMPRESSED, "encoded")
    out = _out(G1_AFFINE * n)
    check(lib().c12381_g1_decompress_batch(encoded, n, out))
    return out.raw[:G1_AFFINE * n]


def from_bytes2(encoded: bytes) -> bytes:
    ensure_init()
    n = _count(encoded, G2_COMPRESSED, "encoded")
    out = _out(G2_AFFINE * n)
    check(lib().c12381_g2_decompress_batch(encoded, n, out))
    return out.raw[:G2_AFFINE * n]


def to_bytes(points: bytes) -> bytes:
    ensure_init()
    n = _count(points, G1_AFFINE, "points")
    out = _out(G1_COMPRESSED * n)
    check(lib().c12381_g1_compress_batch(points, n, out))
    return out.raw[:G1_COMPRESSED * n]


def to_bytes2(points: bytes) -> bytes:
    ensure_init()
    n = _count(points, G2_AFFINE, "points")
    out = _out(G2_COMPRESSED * n)
    check(lib().c12381_g2_compress_batch(points, n, out))
    return out.raw[:G2_COMPRESSED * n]


def is_member(points: bytes) -> bytes:
    """PAIR_G1member for every 96-byte point (one verdict byte each; the identity is not a member, as in the reference)."""
    ensure_init()
    n = _count(points, G1_AFFINE, "points")
    out = _out(n)
    check(lib().c12381_g1_subgroup_check_batch(points, n, out))
    return out.raw[:n]


def is_member2(points: bytes) -> bytes:
    """PAIR_G2member for every 192-byte point."""
    ensure_init()
    n = _count(points, G2_AFFINE, "points")
    out = _out(n)
    check(lib().c12381_g2_subgroup_check_batch(points, n, out))
    return out.raw[:n]


# ---- hash(...) : SHA3-512 over serialised bytes, and hash(...) -> Zp (set.hpp:317-460, zp_number.hpp:538-547), batched ------
def sha3_512(messages: bytes, message_len: int) -> bytes:
    ensure_init()
    B = _count(messages, message_len, "messages") if message_len else 0
    out = _out(64 * B)
    check(lib().c12381_sha3_512_batch(messages, message_len, B, out))
    return out.raw[:64 * B]


def hash_to_zp(messages: bytes, message_len: int) -> bytes:
    ensure_init()
    B = _count(messages, message_len, "messages") if message_len else 0
    out = _out(32 * B)
    check(lib().c12381_hash_to_zp_batch(messages, message_len, B, out))
    return out.raw[:32 * B]


def hash_to_g1(messages: bytes, message_len: int) -> bytes:
    """`hash(...) -> G1` per message: G1Point::from_hash (g1_point.hpp:219-234); 49 B compressed each."""
    ensure_init()
    B = _count(messages, message_len, "messages") if message_len else 0
    out = _out(G1_COMPRESSED * B)
    check(lib().c12381_hash_to_g1_batch(messages, message_len, B, out))
    return out.raw[:G1_COMPRESSED * B]


def map_to_g1(elements: bytes) -> bytes:
    """map_to_point + multiply_cofactor (src/miracl_core_interface.cpp:154-162) of field elements, 48 B big-endian each."""
    ensure_init()
    B = _count(elements, 48, "elements")
    out = _out(G1_COMPRESSED * B)
    check(lib().c12381_map_to_g1_batch(elements, B, out))
    return out.raw[:G1_COMPRESSED * B]


# ---- pairings ------------------------------------------------------------------------------------------------------
def _pairs(g1s: bytes, g2s: bytes, k: int) -> int:
    if not 1 <= k <= _lib.MAX_PAIRS:
        raise ValueError(f"k must be in [1, {_lib.MAX_PAIRS}]")
    b = _count(g1s, G1_AFFINE * k, "g1s")
    if _count(g2s, G2_AFFINE * k, "g2s") != b:
        raise ValueError("g1s and g2s differ in length")
    return b


def pair_ate(p2: bytes, p1: bytes) -> bytes:
    """pair_ate(fp12& result, point2& p2, point1& p1) -> PAIR_ate (:276-279): un-exponentiated Miller value."""
    return miller_batch(p1, p2, 1)


def pair_double_ate(p2: bytes, p1: bytes, q2: bytes, q1: bytes) -> bytes:
    """pair_double_ate -> PAIR_double_ate (:286-289): product of two Miller loops sharing their squarings."""
    return miller_batch(p1 + q1, p2 + q2, 2)


def miller_batch(g1s: bytes, g2s: bytes, k: int) -> bytes:
    ensure_init()
    b = _pairs(g1s, g2s, k)
    out = _out(GT_BYTES * b)
    check(lib().c12381_miller_batch(g1s, g2s, b, k, out))
    return out.raw[:GT_BYTES * b]


def pair_final_exponentiation(values: bytes) -> bytes:
    """pair_final_exponentiation -> PAIR_fexp (:281-284), exponent 3(p^12-1)/r; batched over 576-byte records."""
    ensure_init()
    b = _count(values, GT_BYTES, "values")
    out = _out(GT_BYTES * b)
    check(lib().c12381_final_exp_batch(values, b, out))
    return out.raw[:GT_BYTES * b]


def pairing_product_batch(g1s: bytes, g2s: bytes, k: int) -> bytes:
    """out[b] = fexp(Π_j miller(Q_bj, P_bj)): k pairs per instance, one shared final exponentiation."""
    ensure_init()
    b = _pairs(g1s, g2s, k)
    out = _out(GT_BYTES * b)
    check(lib().c12381_pairing_product_batch(g1s, g2s, b, k, out))
    return out.raw[:GT_BYTES * b]


def pairing_check_batch(g1s: bytes, g2s: bytes, k: int) -> bytes:
    """verdict[b] = 1 iff the k-pair product is the GT identity (operator== on pairs, liner_pair.hpp:336-357)."""
    ensure_init()
    b = _pairs(g1s, g2s, k)
    out = _out(b)
    check(lib().c12381_pairing_check_batch(g1s, g2s, b, k, out))
    return out.raw[:b]


def gt_multiply(a: bytes, b: bytes) -> bytes:
    """multiply(fp12& result, fp12& value) -> FP12_mul (:256-259), batched."""
    ensure_init()
    n = _count(a, GT_BYTES, "a")
    if _count(b, GT_BYTES, "b") != n:
        raise ValueError("gt_multiply: operands differ in length")
    out = _out(GT_BYTES * n)
    check(lib().c12381_gt_mul_batch(a, b, n, out))
    return out.raw[:GT_BYTES * n]


def gt_pow(bases: bytes, exponents: bytes) -> bytes:
    """pow(fp12& result, fp12& base, const big& exponent) -> FP12_pow (:261-264), batched, unitary bases."""
    ensure_init()
    n = _count(bases, GT_BYTES, "bases")
    if _count(exponents, SCALAR, "exponents") != n:
        raise ValueError("gt_pow: bases and exponents differ in length")
    out = _out(GT_BYTES * n)
    check(lib().c12381_gt_pow_batch(bases, exponents, n, out))
    return out.raw[:GT_BYTES * n]


def gt_pow_gs(bases: bytes, exponents: bytes) -> bytes:
    """PAIR_GTpow with the Galbraith-Scott split (pair_BLS12381.cpp:985-1026), batched; bases must lie in GT (order r)."""
    ensure_init()
    n = _count(bases, GT_BYTES, "bases")
    if _count(exponents, SCALAR, "exponents") != n:
        raise ValueError("gt_pow_gs: bases and exponents differ in length")
    out = _out(GT_BYTES * n)
    check(lib().c12381_gt_pow_gs_batch(bases, exponents, n, out))
    return out.raw[:GT_BYTES * n]


# ---- the reference's PODs, passed through unchanged (what the forwarding TU does) ------------------------------------
def sum_of_products_pod(result_point1, n: int, points_point1, numbers_big) -> None:
    ensure_init()
    check(lib().c12381_sum_of_products_miracl(result_point1, n, points_point1, numbers_big))


def sum_of_products2_pod(result_point2, n: int, points_point2, numbers_big) -> None:
    ensure_init()
    check(lib().c12381_sum_of_products2_miracl(result_point2, n, points_point2, numbers_big))


def multiply_pod(object_point1, value_big) -> None:
    ensure_init()
    check(lib().c12381_multiply_point1_miracl(object_point1, value_big))


def multiply2_pod(object_point2, value_big) -> None:
    ensure_init()
    check(lib().c12381_multiply_point2_miracl(object_point2, value_big))


def double_multiply_pod(p1_point1, p2_point1, v1_big, v2_big) -> None:
    ensure_init()
    check(lib().c12381_double_multiply_miracl(p1_point1, p2_point1, v1_big, v2_big))


def pair_ate_pod(result_fp12, p2_point2, p1_point1) -> None:
    ensure_init()
    check(lib().c12381_pair_ate_miracl(result_fp12, p2_point2, p1_point1))


def pair_double_ate_pod(result_fp12, p2, p1, q2, q1) -> None:
    ensure_init()
    check(lib().c12381_pair_double_ate_miracl(result_fp12, p2, p1, q2, q1))


def pair_final_exponentiation_pod(object_fp12) -> None:
    ensure_init()
    check(lib().c12381_pair_final_exponentiation_miracl(object_fp12))


def gt_multiply_pod(result_fp12, value_fp12) -> None:
    ensure_init()
    check(lib().c12381_fp12_multiply_miracl(result_fp12, value_fp12))


def gt_pow_pod(result_fp12, base_fp12, exponent_big) -> None:
    ensure_init()
    check(lib().c12381_fp12_pow_miracl(result_fp12, base_fp12, exponent_big))
