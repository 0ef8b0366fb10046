"""ctypes binding of libc12381_cuda.so (the C ABI of include/c12381_cuda.h).

There is NO CPU fallback: if the shared library has not been built, or no CUDA device is visible, every compute
entry raises.  The library is looked up in-tree only (crypto12381_b200/libc12381_cuda.so)."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get("C12381_LIB_VARIANT", "")   # experimental twin builds (crypto12381_b200/build.py), measurements only
LIB_PATH = os.path.join(HERE, f"libc12381_cuda_{_VARIANT}.so" if _VARIANT else "libc12381_cuda.so")

OK, ENODEV, ECUDA, EINPUT, EARG = 0, -1, -2, -3, -4
MAX_PAIRS = 8

_p = ctypes.c_void_p
_sz = ctypes.c_size_t
_i = ctypes.c_int

# name -> (restype, argtypes); this table is also what tests use to check the exported symbol set
SIGNATURES = {
    "c12381_init": (_i, [_i]),
    "c12381_shutdown": (None, []),
    "c12381_last_error": (ctypes.c_char_p, []),
    "c12381_device": (_i, []),
    "c12381_sync_status": (_i, [_p]),
    "c12381_set_msm_window": (None, [_i]),
    "c12381_set_msm_batch_affine": (None, [_i]),
    "c12381_set_msm_pipelines": (None, [_i]),
    "c12381_set_knob": (None, [_i, _i]),
    "c12381_g1_msm": (_i, [_p, _p, _sz, _p]),
    "c12381_g1_msm_partial": (_i, [_p, _p, _sz, _p]),
    "c12381_g2_msm_partial": (_i, [_p, _p, _sz, _p]),
    "c12381_g1_msm_dev": (_i, [_p, _p, _sz, _p, _p]),
    "c12381_g1_msm_partial_dev": (_i, [_p, _p, _sz, _p, _p]),
    "c12381_g1_sum_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_g2_msm": (_i, [_p, _p, _sz, _p]),
    "c12381_g2_msm_dev": (_i, [_p, _p, _sz, _p, _p]),
    "c12381_g2_msm_partial_dev": (_i, [_p, _p, _sz, _p, _p]),
    "c12381_g2_sum_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_g1_mul_batch": (_i, [_p, _p, _sz, _p]),
    "c12381_g2_mul_batch": (_i, [_p, _p, _sz, _p]),
    "c12381_g1_mul_batch_dev": (_i, [_p, _p, _sz, _p, _p]),
    "c12381_g2_mul_batch_dev": (_i, [_p, _p, _sz, _p, _p]),
    "c12381_g1_fixed_base_mul_batch": (_i, [_p, _sz, _p]),
    "c12381_g2_fixed_base_mul_batch": (_i, [_p, _sz, _p]),
    "c12381_g1_fixed_base_mul_batch_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_g2_fixed_base_mul_batch_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_g1_multi_fixed_base_batch": (_i, [_p, _sz, _p, _sz, _p]),
    "c12381_g2_multi_fixed_base_batch": (_i, [_p, _sz, _p, _sz, _p]),
    "c12381_g1_multi_fixed_base_batch_dev": (_i, [_p, _sz, _p, _sz, _p, _p]),
    "c12381_g2_multi_fixed_base_batch_dev": (_i, [_p, _sz, _p, _sz, _p, _p]),
    "c12381_g1_decompress_batch": (_i, [_p, _sz, _p]),
    "c12381_g2_decompress_batch": (_i, [_p, _sz, _p]),
    "c12381_g1_compress_batch": (_i, [_p, _sz, _p]),
    "c12381_g2_compress_batch": (_i, [_p, _sz, _p]),
    "c12381_g1_decompress_batch_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_g2_decompress_batch_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_g1_compress_batch_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_g2_compress_batch_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_g1_subgroup_check_batch": (_i, [_p, _sz, _p]),
    "c12381_g2_subgroup_check_batch": (_i, [_p, _sz, _p]),
    "c12381_g1_subgroup_check_batch_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_g2_subgroup_check_batch_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_sha3_512_batch": (_i, [_p, _sz, _sz, _p]),
    "c12381_hash_to_zp_batch": (_i, [_p, _sz, _sz, _p]),
    "c12381_sha3_512_batch_dev": (_i, [_p, _sz, _sz, _p, _p]),
    "c12381_hash_to_zp_batch_dev": (_i, [_p, _sz, _sz, _p, _p]),
    "c12381_hash_to_g1_batch": (_i, [_p, _sz, _sz, _p]),
    "c12381_map_to_g1_batch": (_i, [_p, _sz, _p]),
    "c12381_hash_to_g1_batch_dev": (_i, [_p, _sz, _sz, _p, _p]),
    "c12381_map_to_g1_batch_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_miller_batch": (_i, [_p, _p, _sz, _i, _p]),
    "c12381_final_exp_batch": (_i, [_p, _sz, _p]),
    "c12381_pairing_product_batch": (_i, [_p, _p, _sz, _i, _p]),
    "c12381_pairing_check_batch": (_i, [_p, _p, _sz, _i, _p]),
    "c12381_miller_batch_dev": (_i, [_p, _p, _sz, _i, _p, _p]),
    "c12381_final_exp_batch_dev": (_i, [_p, _sz, _p, _p]),
    "c12381_pairing_product_batch_dev": (_i, [_p, _p, _sz, _i, _p, _p]),
    "c12381_pairing_check_batch_dev": (_i, [_p, _p, _sz, _i, _p, _p]),
    "c12381_gt_mul_batch": (_i, [_p, _p, _sz, _p]),
    "c12381_gt_pow_batch": (_i, [_p, _p, _sz, _p]),
    "c12381_gt_mul_batch_dev": (_i, [_p, _p, _sz, _p, _p]),
    "c12381_gt_pow_batch_dev": (_i, [_p, _p, _sz, _p, _p]),
    "c12381_gt_pow_gs_batch": (_i, [_p, _p, _sz, _p]),
    "c12381_gt_pow_gs_batch_dev": (_i, [_p, _p, _sz, _p, _p]),
    "c12381_sum_of_products_miracl": (_i, [_p, _i, _p, _p]),
    "c12381_multiply_point1_miracl": (_i, [_p, _p]),
    "c12381_double_multiply_miracl": (_i, [_p, _p, _p, _p]),
    "c12381_multiply_point2_miracl": (_i, [_p, _p]),
    "c12381_sum_of_products2_miracl": (_i, [_p, _i, _p, _p]),
    "c12381_pair_ate_miracl": (_i, [_p, _p, _p]),
    "c12381_pair_double_ate_miracl": (_i, [_p, _p, _p, _p, _p]),
    "c12381_pair_multi_ate_miracl": (_i, [_p, ctypes.c_int, _p, _p]),
    "c12381_pair_final_exponentiation_miracl": (_i, [_p]),
    "c12381_fp12_multiply_miracl": (_i, [_p, _p]),
    "c12381_fp12_pow_miracl": (_i, [_p, _p, _p]),
    "c12381_probe": (_i, [_i, _i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "c12381_launch_count": (ctypes.c_ulonglong, []),
    "c12381_last_msm_stats": (_i, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                   ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(_i)]),
    "c12381_last_msm_phases": (_i, [ctypes.POINTER(ctypes.c_double)]),
    "c12381_last_msm_shape": (_i, [ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "c12381_set_pairing_kernel": (None, [_i]),
}


class C12381Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"c12381 error {code}: {message}")
        self.code = code


_lib = None


def lib() -> ctypes.CDLL:
    """The loaded shared library (no device needed to load it)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m crypto12381_b200.build` "
                               "(there is no CPU fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise C12381Error(rc, lib().c12381_last_error().decode(errors="replace"))


def init(device: int = 0) -> None:
    """Bind this process to one CUDA device (one process per GPU)."""
    check(lib().c12381_init(device))


def ensure_init() -> None:
    if lib().c12381_device() < 0:
        init(int(os.environ.get("LOCAL_RANK", "0")))
