"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous.

* q-phase (``ik_only``): clips are independent (reference ``stac.py:425-440``), so they are block-
  partitioned across ranks and no data-path collective is needed; results are all-gathered only to
  hand every rank the reference's full ``StacData``.
* m-phase: the closed form needs ``s = sum_t R^T z`` and ``z2`` over all sampled frames
  (``stac_core.py:157-160``); frames are sharded and the 3K+2 numbers are all-reduced.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of n units for `rank`; blocks differ by at most one unit."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_m_stats(buf: torch.Tensor) -> torch.Tensor:
    """Sum a flat m-phase buffer over ranks IN PLACE: ``[s (3K), z2, T]`` (3K+2 floats, one NCCL all-reduce on the stream
    the statistics kernel ran on) or the 1-float residual of the objective."""
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf


def allgather_blocks(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Concatenate per-rank contiguous blocks (dim 0) produced with `shard_range` on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
