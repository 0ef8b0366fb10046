"""crypto12381_b200 — B200 (sm_100a) implementation of crypto12381's data-parallel hot path: G1/G2 multi-scalar
sums, batched scalar multiplication, batched pairings / pairing products, behind the reference's bridge API.

  bridge       host-side mirror of crypto12381::detail::miracl_core for this path (bytes in, bytes out)
  device       the same operations on CUDA-resident torch uint8 tensors (no host copies)
  distributed  sharding over one-process-per-GPU ranks (torch.distributed)
"""
from . import _lib  # noqa: F401
from ._lib import C12381Error, init  # noqa: F401
