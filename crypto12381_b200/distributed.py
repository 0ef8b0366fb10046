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


def sharded_msm(points, scalars, point_bytes: int, partial_fn: Callable, sum_fn: Callable, group=None):
    """`points` / `scalars` are THIS RANK's slice (uint8 tensors).  Returns the compressed total on every rank."""
    import torch
    import torch.distributed as dist

    partial = partial_fn(points, scalars)  # point_bytes-long affine encoding of this rank's partial sum
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sum_fn(partial)
    world = dist.get_world_size(group)
    gathered = torch.empty(world * point_bytes, dtype=torch.uint8, device=partial.device)
    dist.all_gather_into_tensor(gathered, partial.contiguous(), group=group)
    return sum_fn(gathered)


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


def gather_results(local, group=None):
    """Final gather of per-instance results (equal-sized shards) — the only collective of the batched-instance
    paths, and only when the caller wants them on every rank."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out
