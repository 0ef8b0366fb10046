"""TEST-ONLY: builds (g++) and loads tests/hostmirror/libhostmirror.so — the product's host+device templates
compiled for the CPU so kernel bodies can be checked against the oracle without a GPU."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostmirror", "mirror.cpp")
LIB = os.path.join(HERE, "hostmirror", "libhostmirror.so")
CSRC = os.path.join(os.path.dirname(HERE), "crypto12381_b200", "csrc")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if _stale():
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
        _lib = ctypes.CDLL(LIB)
    return _lib


def fp_op(op, a, b=bytes(48)):
    out = ctypes.create_string_buffer(48)
    lib().hm_fp_op(op, a, b, out)
    return out.raw


def g1_msm(points, scalars, c, seg=4):
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(49)
    rc = lib().hm_g1_msm(points, scalars, n, c, seg, out)
    assert rc == 0
    return out.raw


def g2_msm(points, scalars, c, seg=4):
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(97)
    rc = lib().hm_g2_msm(points, scalars, n, c, seg, out)
    assert rc == 0
    return out.raw


def g1_mul(points, scalars):
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(49 * n)
    assert lib().hm_g1_mul(points, scalars, n, out) == 0
    return out.raw


def g2_mul(points, scalars):
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(97 * n)
    assert lib().hm_g2_mul(points, scalars, n, out) == 0
    return out.raw


def g1_fixed_base(scalars):
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(96 * n)
    lib().hm_g1_fixed_base(scalars, n, out)
    return out.raw


def g2_fixed_base(scalars):
    n = len(scalars) // 32
    out = ctypes.create_string_buffer(192 * n)
    lib().hm_g2_fixed_base(scalars, n, out)
    return out.raw


def pairing_product(g1, g2, k, mode=1):
    B = len(g1) // (96 * k)
    out = ctypes.create_string_buffer(576 * B)
    assert lib().hm_pairing_product(g1, g2, B, k, mode, out) == 0
    return out.raw


def final_exp(f):
    B = len(f) // 576
    out = ctypes.create_string_buffer(576 * B)
    lib().hm_final_exp(f, B, out)
    return out.raw


def gt_mul(a, b):
    B = len(a) // 576
    out = ctypes.create_string_buffer(576 * B)
    lib().hm_gt_mul(a, b, B, out)
    return out.raw


def gt_pow_gs(a, s):
    B = len(a) // 576
    out = ctypes.create_string_buffer(576 * B)
    lib().hm_gt_pow_gs(a, s, B, out)
    return out.raw


def gt_pow(a, s):
    B = len(a) // 576
    out = ctypes.create_string_buffer(576 * B)
    lib().hm_gt_pow(a, s, B, out)
    return out.raw
