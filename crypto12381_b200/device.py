"""The hot-path operations on CUDA-resident data: torch uint8 tensors in, torch uint8 tensors out, work enqueued on
torch's CURRENT stream (so torch.cuda.Event timing brackets it).  PyTorch is only the owner of device memory and
streams here; every kernel is this library's own (libc12381_cuda.so `_dev` entries).

Malformed input cannot be reported by an enqueue-only call: use `sync_status()` (synchronises, raises
C12381Error(EINPUT) if a kernel flagged anything since the last check)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ensure_init, lib

G1_AFFINE, G2_AFFINE, G1_COMPRESSED, G2_COMPRESSED, GT_BYTES, SCALAR = 96, 192, 49, 97, 576, 32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.uint8 or not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{what}: need a contiguous CUDA uint8 tensor")
    return t


def _new(n: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(max(n, 1), dtype=torch.uint8, device=like.device)[:n]


def _records(t: torch.Tensor, record: int, what: str, n: int | None = None) -> int:
    """`t` is a contiguous CUDA uint8 tensor holding a whole number of `record`-byte items (exactly n of them when n is given);
    returns the count.  Every wrapper sizes its raw pointers through this: a mismatch is a ValueError here, never an
    out-of-bounds access on the device."""
    _chk(t, what)
    if record <= 0 or t.numel() % record != 0:
        raise ValueError(f"{what}: {t.numel()} bytes is not a whole number of {record}-byte records")
    got = t.numel() // record
    if n is not None and got != n:
        raise ValueError(f"{what}: {got} records of {record} bytes, expected {n}")
    return got


def _out(out, nbytes: int, like: torch.Tensor) -> torch.Tensor:
    """the caller's output buffer checked for dtype / device / contiguity / size, or a fresh one"""
    if out is None:
        return _new(nbytes, like)
    _chk(out, "out")
    if out.device != like.device or out.numel() != nbytes:
        raise ValueError(f"out: need {nbytes} bytes on {like.device}, got {out.numel()} on {out.device}")
    return out


def sync_status() -> None:
    ensure_init()
    check(lib().c12381_sync_status(_stream()))


def _msm(fn_name: str, point_bytes: int, out_bytes: int, points: torch.Tensor, scalars: torch.Tensor, out=None):
    ensure_init()
    n = _records(scalars, SCALAR, "scalars")
    _records(points, point_bytes, "points", n)
    out = _out(out, out_bytes, points)
    check(getattr(lib(), fn_name)(points.data_ptr(), scalars.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def g1_msm(points, scalars, out=None):
    """Σ scalars[i]·points[i] over G1 -> 49-byte compressed point."""
    return _msm("c12381_g1_msm_dev", G1_AFFINE, G1_COMPRESSED, points, scalars, out)


def g1_msm_partial(points, scalars, out=None):
    """Same sum as a 96-byte AFFINE point: the per-rank partial of a sharded MSM."""
    return _msm("c12381_g1_msm_partial_dev", G1_AFFINE, G1_AFFINE, points, scalars, out)


def g2_msm(points, scalars, out=None):
    return _msm("c12381_g2_msm_dev", G2_AFFINE, G2_COMPRESSED, points, scalars, out)


def g2_msm_partial(points, scalars, out=None):
    return _msm("c12381_g2_msm_partial_dev", G2_AFFINE, G2_AFFINE, points, scalars, out)


def g1_sum(points, out=None):
    """Σ points[i] (merging all-gathered partials) -> 49-byte compressed point."""
    ensure_init()
    n = _records(points, G1_AFFINE, "points")
    out = _out(out, G1_COMPRESSED, points)
    check(lib().c12381_g1_sum_dev(points.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def g2_sum(points, out=None):
    ensure_init()
    n = _records(points, G2_AFFINE, "points")
    out = _out(out, G2_COMPRESSED, points)
    check(lib().c12381_g2_sum_dev(points.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def g1_mul_batch(points, scalars, out=None):
    ensure_init()
    n = _records(scalars, SCALAR, "scalars")
    _records(points, G1_AFFINE, "points", n)
    out = _out(out, n * G1_COMPRESSED, points)
    check(lib().c12381_g1_mul_batch_dev(points.data_ptr(), scalars.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def g2_mul_batch(points, scalars, out=None):
    ensure_init()
    n = _records(scalars, SCALAR, "scalars")
    _records(points, G2_AFFINE, "points", n)
    out = _out(out, n * G2_COMPRESSED, points)
    check(lib().c12381_g2_mul_batch_dev(points.data_ptr(), scalars.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def g1_fixed_base_mul_batch(scalars, out=None):
    ensure_init()
    n = _records(scalars, SCALAR, "scalars")
    out = _out(out, n * G1_AFFINE, scalars)
    check(lib().c12381_g1_fixed_base_mul_batch_dev(scalars.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def g2_fixed_base_mul_batch(scalars, out=None):
    ensure_init()
    n = _records(scalars, SCALAR, "scalars")
    out = _out(out, n * G2_AFFINE, scalars)
    check(lib().c12381_g2_fixed_base_mul_batch_dev(scalars.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def _pairing(fn_name: str, g1s, g2s, k: int, out_per_instance: int, out=None):
    ensure_init()
    if not 1 <= k <= _lib.MAX_PAIRS:
        raise ValueError("k out of range")
    b = _records(g1s, G1_AFFINE * k, "g1s")
    _records(g2s, G2_AFFINE * k, "g2s", b)
    out = _out(out, b * out_per_instance, g1s)
    check(getattr(lib(), fn_name)(g1s.data_ptr(), g2s.data_ptr(), b, k, out.data_ptr(), _stream()))
    return out


def miller_batch(g1s, g2s, k, out=None):
    return _pairing("c12381_miller_batch_dev", g1s, g2s, k, GT_BYTES, out)


def pairing_product_batch(g1s, g2s, k, out=None):
    return _pairing("c12381_pairing_product_batch_dev", g1s, g2s, k, GT_BYTES, out)


def pairing_check_batch(g1s, g2s, k, out=None):
    return _pairing("c12381_pairing_check_batch_dev", g1s, g2s, k, 1, out)


def final_exp_batch(values, out=None):
    ensure_init()
    b = _records(values, GT_BYTES, "values")
    out = _out(out, b * GT_BYTES, values)
    check(lib().c12381_final_exp_batch_dev(values.data_ptr(), b, out.data_ptr(), _stream()))
    return out


def gt_mul_batch(a, b, out=None):
    ensure_init()
    n = _records(a, GT_BYTES, "a")
    _records(b, GT_BYTES, "b", n)
    out = _out(out, n * GT_BYTES, a)
    check(lib().c12381_gt_mul_batch_dev(a.data_ptr(), b.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def gt_pow_batch(a, scalars, out=None):
    ensure_init()
    n = _records(a, GT_BYTES, "a")
    _records(scalars, SCALAR, "scalars", n)
    out = _out(out, n * GT_BYTES, a)
    check(lib().c12381_gt_pow_batch_dev(a.data_ptr(), scalars.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def gt_pow_gs_batch(a, scalars, out=None):
    """a[b]^scalars[b] for a in GT through the Galbraith-Scott split (c12381_gt_pow_gs_batch_dev)."""
    ensure_init()
    n = _records(a, GT_BYTES, "a")
    _records(scalars, SCALAR, "scalars", n)
    out = _out(out, n * GT_BYTES, a)
    check(lib().c12381_gt_pow_gs_batch_dev(a.data_ptr(), scalars.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def _multi_fixed(fn_name, point_bytes, bases, scalars, out=None):
    ensure_init()
    m = _records(bases, point_bytes, "bases")
    if m == 0:
        raise ValueError("bases: at least one base")
    B = _records(scalars, SCALAR * m, "scalars")
    out = _out(out, B * point_bytes, scalars)
    check(getattr(lib(), fn_name)(bases.data_ptr(), m, scalars.data_ptr(), B, out.data_ptr(), _stream()))
    return out


def g1_multi_fixed_base_batch(bases, scalars, out=None):
    """out[b] = Σ_j scalars[b][j]·bases[j] (m shared G1 bases) -> affine 96 B per instance."""
    return _multi_fixed("c12381_g1_multi_fixed_base_batch_dev", G1_AFFINE, bases, scalars, out)


def g2_multi_fixed_base_batch(bases, scalars, out=None):
    return _multi_fixed("c12381_g2_multi_fixed_base_batch_dev", G2_AFFINE, bases, scalars, out)


def _convert(fn_name, in_bytes, out_bytes, data, out=None):
    ensure_init()
    n = _records(data, in_bytes, "data")
    out = _out(out, n * out_bytes, data)
    check(getattr(lib(), fn_name)(data.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def g1_decompress_batch(data, out=None):
    return _convert("c12381_g1_decompress_batch_dev", G1_COMPRESSED, G1_AFFINE, data, out)


def g2_decompress_batch(data, out=None):
    return _convert("c12381_g2_decompress_batch_dev", G2_COMPRESSED, G2_AFFINE, data, out)


def g1_compress_batch(data, out=None):
    return _convert("c12381_g1_compress_batch_dev", G1_AFFINE, G1_COMPRESSED, data, out)


def g2_compress_batch(data, out=None):
    return _convert("c12381_g2_compress_batch_dev", G2_AFFINE, G2_COMPRESSED, data, out)


def _hash(fn_name, out_bytes, msgs, msg_len, out=None):
    ensure_init()
    if msg_len <= 0:
        raise ValueError("msg_len must be positive")
    n = _records(msgs, msg_len, "msgs")
    out = _out(out, n * out_bytes, msgs)
    check(getattr(lib(), fn_name)(msgs.data_ptr(), msg_len, n, out.data_ptr(), _stream()))
    return out


def sha3_512_batch(msgs, msg_len, out=None):
    """SHA3-512 of every msg_len-byte message (c12381_sha3_512_batch_dev); n x 64 B."""
    return _hash("c12381_sha3_512_batch_dev", 64, msgs, msg_len, out)


def hash_to_zp_batch(msgs, msg_len, out=None):
    """`hash(...) -> Zp` per message (c12381_hash_to_zp_batch_dev); n x 32 B big-endian."""
    return _hash("c12381_hash_to_zp_batch_dev", 32, msgs, msg_len, out)


def hash_to_g1_batch(msgs, msg_len, out=None):
    """`hash(...) -> G1` per message (c12381_hash_to_g1_batch_dev); n x 49 B compressed."""
    return _hash("c12381_hash_to_g1_batch_dev", G1_COMPRESSED, msgs, msg_len, out)


def launch_count() -> int:
    return int(lib().c12381_launch_count())


def last_msm_stats() -> dict:
    import ctypes
    a, t = ctypes.c_double(), ctypes.c_double()
    adds, c = ctypes.c_ulonglong(), ctypes.c_int()
    check(lib().c12381_last_msm_stats(ctypes.byref(a), ctypes.byref(t), ctypes.byref(adds), ctypes.byref(c)))
    ph = (ctypes.c_double * 8)()
    check(lib().c12381_last_msm_phases(ph))
    names = ["recode", "sort", "bounds_order", "parse", "accumulate", "reduce1", "reduce2", "finish"]
    r, p, g = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(lib().c12381_last_msm_shape(ctypes.byref(r), ctypes.byref(p), ctypes.byref(g)))
    return {"accumulate_ms": a.value, "total_ms": t.value, "bucket_adds": adds.value, "window_bits": c.value,
            "ba_rounds": r.value, "ba_pipelines": p.value, "upload_groups": g.value, "phases_ms": dict(zip(names, list(ph)))}


def probe(kind: int, iters: int = 2000) -> dict:
    """Integer-multiply roofline probes; see include/c12381_cuda.h for the kinds."""
    import ctypes
    ensure_init()
    g, ms = ctypes.c_double(), ctypes.c_double()
    check(lib().c12381_probe(kind, iters, ctypes.byref(g), ctypes.byref(ms)))
    return {"gops": g.value, "ms": ms.value}
