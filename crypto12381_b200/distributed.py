"""Sharding of the hot path over one-process-per-GPU ranks.

  * MSM: contiguous n/G slices of (point, scalar) per rank; each rank runs the whole single-GPU pipeline on its
    slice and produces ONE affine partial point; the partials are all-gathered (96 B / 192 B per rank — latency
    only) and every rank adds them in rank order, so all ranks hold the same bit-exact result.
  * Batched scalar multiplication / pairings / verification: independent instances, B/G per rank, no data-path
    collective; results stay on the owning rank (gather them only if the caller needs them in one place).

The compute callables are injected so the host logic (slice bounds, the collective, merge order) is testable on
CPU ranks with the `gloo` backend; the defaults are the CUDA entries of `crypto12381_b200.device`."""
from __future__ import annotations

from typing import Callable, Tuple


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n items for `rank` (the first n % world ranks get one extra)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _all_gather_flat(local, group=None):
    """all_gather_into_tensor of equal-sized uint8 shards.  NCCL moves CUDA tensors directly; a gloo group (CPU ranks, or
    several ranks sharing one GPU in tests) has no CUDA all-gather, so CUDA shards bounce through the host - these messages
    are a few hundred bytes."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    local = local.contiguous()
    if local.is_cuda and dist.get_backend(group) == "gloo":
        out = torch.empty(world * local.numel(), dtype=local.dtype)
        dist.all_gather_into_tensor(out, local.cpu(), group=group)
        return out.to(local.device)
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def sharded_msm(points, scalars, point_bytes: int, partial_fn: Callable, sum_fn: Callable, group=None):
    """`points` / `scalars` are THIS RANK's slice (uint8 tensors).  Returns the compressed total on every rank."""
    import torch
    import torch.distributed as dist

    partial = partial_fn(points, scalars)  # point_bytes-long affine encoding of this rank's partial sum
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sum_fn(partial)
    return sum_fn(_all_gather_flat(partial, group))


def _single(group) -> bool:
    import torch.distributed as dist
    return not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1


def g1_msm_sharded(points, scalars, group=None):
    from . import device
    if _single(group):   # nothing to merge: one pipeline, one normalisation
        return device.g1_msm(points, scalars)
    return sharded_msm(points, scalars, device.G1_AFFINE, device.g1_msm_partial, device.g1_sum, group)


def g2_msm_sharded(points, scalars, group=None):
    from . import device
    if _single(group):
        return device.g2_msm(points, scalars)
    return sharded_msm(points, scalars, device.G2_AFFINE, device.g2_msm_partial, device.g2_sum, group)


def _msm_sharded_host(h_points, h_scalars, group, host_fn: str, partial_fn: str, point_bytes: int, out_bytes: int, sum_fn) -> bytes:
    """HOST buffers in, host bytes out: this rank's slice goes through the host-pointer C-ABI entry (the upload overlaps the
    scalar-only stages inside it); with several ranks the affine partials are all-gathered over NCCL and merged on the device.
    `h_points` / `h_scalars`: contiguous CPU uint8 tensors (pin them: pageable memory makes the copies synchronous)."""
    import torch

    from . import _lib
    _lib.ensure_init()
    n = h_scalars.numel() // 32
    if h_scalars.is_cuda or h_points.is_cuda or h_scalars.numel() != 32 * n or h_points.numel() != point_bytes * n:
        raise ValueError("need CPU uint8 tensors of n x %d and n x 32 bytes" % point_bytes)
    if _single(group):
        out = torch.empty(out_bytes, dtype=torch.uint8)
        _lib.check(getattr(_lib.lib(), host_fn)(h_points.data_ptr(), h_scalars.data_ptr(), n, out.data_ptr()))
        return bytes(out.numpy())
    import torch.distributed as dist
    part = torch.empty(point_bytes, dtype=torch.uint8).pin_memory()
    _lib.check(getattr(_lib.lib(), partial_fn)(h_points.data_ptr(), h_scalars.data_ptr(), n, part.data_ptr()))
    dev = torch.device("cuda", _lib.lib().c12381_device())
    return bytes(sum_fn(_all_gather_flat(part.to(dev, non_blocking=True), group)).cpu().numpy())


def g1_msm_sharded_host(h_points, h_scalars, group=None) -> bytes:
    """The sharded G1 sum end to end from host memory: 49-byte compressed total on every rank."""
    from . import device
    return _msm_sharded_host(h_points, h_scalars, group, "c12381_g1_msm", "c12381_g1_msm_partial", device.G1_AFFINE, device.G1_COMPRESSED, device.g1_sum)


def g2_msm_sharded_host(h_points, h_scalars, group=None) -> bytes:
    from . import device
    return _msm_sharded_host(h_points, h_scalars, group, "c12381_g2_msm", "c12381_g2_msm_partial", device.G2_AFFINE, device.G2_COMPRESSED, device.g2_sum)


def gather_results(local, group=None):
    """Final gather of per-instance results (equal-sized shards) — the only collective of the batched-instance
    paths, and only when the caller wants them on every rank."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    return _all_gather_flat(local, group)
