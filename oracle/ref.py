"""ctypes loader for oracle/_ref/libref12381.so — the UNMODIFIED reference (bridge + MIRACL-core) behind
the extern "C" shim oracle/ref_shim.cpp.  TEST INFRASTRUCTURE — NOT PRODUCT CODE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref12381.so")

_u8p = ctypes.c_char_p
_sz = ctypes.c_size_t
_int = ctypes.c_int


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle` where /root/reference exists")
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def hardware_threads() -> int:
    return int(lib().ref_hardware_threads())


def struct_sizes():
    out = (ctypes.c_int * 5)()
    lib().ref_struct_sizes(out)
    return dict(zip(("big", "fp", "point1", "point2", "fp12"), list(out)))


def random_scalars(seed: str | bytes, n: int) -> bytes:
    seed = seed.encode() if isinstance(seed, str) else seed
    out = ctypes.create_string_buffer(32 * n)
    lib().ref_random_scalars(seed, _int(len(seed)), _sz(n), out)
    return out.raw


def g1_generator() -> bytes:
    out = ctypes.create_string_buffer(96)
    lib().ref_g1_generator(out)
    return out.raw


def g2_generator() -> bytes:
    out = ctypes.create_string_buffer(192)
    lib().ref_g2_generator(out)
    return out.raw


def g1_fixed_base_mul(scalars: bytes, threads: int = 1) -> bytes:
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(96 * n)
    lib().ref_g1_fixed_base_mul(scalars, _sz(n), out, _int(threads))
    return out.raw


def g2_fixed_base_mul(scalars: bytes, threads: int = 1) -> bytes:
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(192 * n)
    lib().ref_g2_fixed_base_mul(scalars, _sz(n), out, _int(threads))
    return out.raw


def g1_mul_batch(points: bytes, scalars: bytes, threads: int = 1) -> bytes:
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(49 * n)
    ok = lib().ref_g1_mul_batch(points, scalars, _sz(n), out, _int(threads))
    assert ok == 1
    return out.raw


def g2_mul_batch(points: bytes, scalars: bytes, threads: int = 1) -> bytes:
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(97 * n)
    ok = lib().ref_g2_mul_batch(points, scalars, _sz(n), out, _int(threads))
    assert ok == 1
    return out.raw


def g1_msm(points: bytes, scalars: bytes, algo: int = 0, threads: int = 1) -> bytes:
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(49)
    ok = lib().ref_g1_msm(points, scalars, _sz(n), out, _int(algo), _int(threads))
    assert ok == 1
    return out.raw


def g2_msm(points: bytes, scalars: bytes, threads: int = 1) -> bytes:
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(97)
    ok = lib().ref_g2_msm(points, scalars, _sz(n), out, _int(threads))
    assert ok == 1
    return out.raw


def pairing_product_batch(g1: bytes, g2: bytes, k: int, mode: int = 1, threads: int = 1) -> bytes:
    B = len(g1) // (96 * k)
    out = ctypes.create_string_buffer(576 * B)
    ok = lib().ref_pairing_product_batch(g1, g2, _sz(B), _int(k), _int(mode), out, _int(threads))
    assert ok == 1
    return out.raw


def final_exp_batch(f: bytes, threads: int = 1) -> bytes:
    B = len(f) // 576
    out = ctypes.create_string_buffer(576 * B)
    lib().ref_final_exp_batch(f, _sz(B), out, _int(threads))
    return out.raw


def gt_mul_batch(a: bytes, b: bytes) -> bytes:
    B = len(a) // 576
    out = ctypes.create_string_buffer(576 * B)
    lib().ref_gt_mul_batch(a, b, _sz(B), out)
    return out.raw


def gt_pow_batch(a: bytes, scalars: bytes, threads: int = 1) -> bytes:
    B = len(a) // 576
    out = ctypes.create_string_buffer(576 * B)
    lib().ref_gt_pow_batch(a, scalars, _sz(B), out, _int(threads))
    return out.raw


def g1_decompress(c: bytes):
    n = len(c) // 49
    out = ctypes.create_string_buffer(96 * n)
    ok = lib().ref_g1_decompress(c, _sz(n), out)
    return out.raw, ok == 1


def g2_decompress(c: bytes):
    n = len(c) // 97
    out = ctypes.create_string_buffer(192 * n)
    ok = lib().ref_g2_decompress(c, _sz(n), out)
    return out.raw, ok == 1


def g1_compress(a: bytes) -> bytes:
    n = len(a) // 96
    out = ctypes.create_string_buffer(49 * n)
    lib().ref_g1_compress(a, _sz(n), out)
    return out.raw


def g2_compress(a: bytes) -> bytes:
    n = len(a) // 192
    out = ctypes.create_string_buffer(97 * n)
    lib().ref_g2_compress(a, _sz(n), out)
    return out.raw


# --- raw MIRACL structs (for the *_miracl ABI tests) ------------------------------------------------
def make_point1(a: bytes, unnormalise: bool = False):
    n = len(a) // 96
    buf = ctypes.create_string_buffer(192 * n)
    lib().ref_make_point1(a, _sz(n), buf)
    if unnormalise:
        lib().ref_point1_unnormalise(buf, _sz(n))
    return buf


def make_point2(a: bytes, unnormalise: bool = False):
    n = len(a) // 192
    buf = ctypes.create_string_buffer(384 * n)
    lib().ref_make_point2(a, _sz(n), buf)
    if unnormalise:
        lib().ref_point2_unnormalise(buf, _sz(n))
    return buf


def make_big(s: bytes):
    n = len(s) // 32
    buf = ctypes.create_string_buffer(56 * n)
    lib().ref_make_big(s, _sz(n), buf)
    return buf


def point1_to_c49(buf, n: int) -> bytes:
    out = ctypes.create_string_buffer(49 * n)
    lib().ref_point1_to_c49(buf, _sz(n), out)
    return out.raw


def point2_to_c97(buf, n: int) -> bytes:
    out = ctypes.create_string_buffer(97 * n)
    lib().ref_point2_to_c97(buf, _sz(n), out)
    return out.raw


def fp12_to_bytes(buf, n: int) -> bytes:
    out = ctypes.create_string_buffer(576 * n)
    lib().ref_fp12_to_bytes(buf, _sz(n), out)
    return out.raw


def g1_member(points: bytes) -> bytes:
    n = len(points) // 96
    out = ctypes.create_string_buffer(max(n, 1))
    assert lib().ref_g1_member(points, _sz(n), out) == 1
    return out.raw[:n]


def g2_member(points: bytes) -> bytes:
    n = len(points) // 192
    out = ctypes.create_string_buffer(max(n, 1))
    assert lib().ref_g2_member(points, _sz(n), out) == 1
    return out.raw[:n]


def hash_to_g1(msgs: bytes, msg_len: int, n: int, threads: int = 1) -> bytes:
    """G1Point::from_hash of the SHA3-512 digest of each message; n x 49 B compressed."""
    out = ctypes.create_string_buffer(49 * n)
    lib().ref_hash_to_g1(msgs, _sz(msg_len), _sz(n), out, _int(threads))
    return out.raw


def map_to_g1(u48: bytes, n: int) -> bytes:
    """map_to_point + multiply_cofactor of n field elements (48 B big-endian, < p); n x 49 B compressed."""
    out = ctypes.create_string_buffer(49 * n)
    lib().ref_map_to_g1(u48, _sz(n), out)
    return out.raw


# --- BBS+ (examples/bbs-plus/src/bbs+.cpp:38-73) on the reference's own bridge arithmetic ------------------------------
def bbs_sign_batch(g1: bytes, h0: bytes, h: bytes, gamma32: bytes, rows: bytes, n: int, threads: int = 1, xs: bytes | None = None) -> bytes:
    """A_i = (g1 * h0^r_i * prod h_j^m_ij)^(1 / (gamma + x_i)), compressed 49 B each.  g1, h0: affine 96 B; h: n x 96 B;
    rows: B x (2 + n) x 32 B scalars (1, r_i, m_i0 ..); xs: B x 32 B (required)."""
    assert xs is not None and len(h) == 96 * n
    B = len(rows) // (32 * (2 + n))
    assert len(xs) == 32 * B
    out = ctypes.create_string_buffer(49 * B)
    assert lib().ref_bbs_sign_batch(g1, h0, h, _int(n), gamma32, rows, xs, _sz(B), out, _int(threads)) == 1
    return out.raw


def bbs_verify_batch(g1: bytes, g2: bytes, h0: bytes, h: bytes, w: bytes, n: int, A49: bytes, rows: bytes, xrows: bytes, threads: int = 1) -> bytes:
    """verdict bytes of pair(A, w * g2^x) == pair(g1 * h0^r * prod h_j^m_j, g2); xrows: B x 2 x 32 B rows (1, x_i)."""
    B = len(A49) // 49
    assert len(rows) == 32 * (2 + n) * B and len(xrows) == 64 * B and len(h) == 96 * n
    out = ctypes.create_string_buffer(max(B, 1))
    assert lib().ref_bbs_verify_batch(g1, g2, h0, h, w, _int(n), A49, rows, xrows, _sz(B), out, _int(threads)) == 1
    return out.raw[:B]
