"""Slice-sharded execution of a batch of series across the GPUs of one box.

Every stage of the hot path is per-slice except the coronal image (one row per slice plus a
global min/max), so a series is cut into contiguous z-ranges, one per rank, and
the only exchange is

  * all-gather of the per-slice coronal rows   (n_local x W int16 per series per rank)
  * all-gather of the per-shard (min, max)     (2 int32 per series per rank)
  * all-reduce of the selected slice indices   (4 int32 per series; series s is decided on rank s % world)

The plumbing below works on whatever ``torch.distributed`` backend is initialised (NCCL over
NVLink on the box, gloo in the CPU tests) and on a single process without a process group.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_slices: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous z-range [z0, z1) of rank ``rank`` (in InstanceNumber order); sizes differ by at most 1."""
    base, rem = divmod(n_slices, world_size)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def owner_of_series(s: int, world_size: int) -> int:
    return s % world_size


def gather_rows(rows_local: torch.Tensor, minmax_local: torch.Tensor, n_slices: int):
    """rows_local [S, n_local, W] int16 (this rank's z-range, sorted), minmax_local [S, 2] int32 ->
    (rows [S, n_slices, W], minmax [S, 2]) identical on every rank."""
    rank, ws = world()
    if ws == 1:
        return rows_local, minmax_local
    S, nl, W = rows_local.shape
    nmax = max(shard_range(n_slices, ws, r)[1] - shard_range(n_slices, ws, r)[0] for r in range(ws))
    send = rows_local
    if nl < nmax:                                      # ragged shards: pad to the largest
        send = torch.zeros((S, nmax, W), dtype=rows_local.dtype, device=rows_local.device)
        send[:, :nl] = rows_local
    # neither NCCL nor gloo has a 16-bit integer type: ship the rows as bytes
    send8 = send.contiguous().view(torch.uint8)
    parts8 = [torch.empty_like(send8) for _ in range(ws)]
    dist.all_gather(parts8, send8)
    parts = [p.view(rows_local.dtype) for p in parts8]
    mms = [torch.empty_like(minmax_local) for _ in range(ws)]
    dist.all_gather(mms, minmax_local.contiguous())
    rows = torch.cat([parts[r][:, :shard_range(n_slices, ws, r)[1] - shard_range(n_slices, ws, r)[0]] for r in range(ws)], 1)
    mm = torch.stack(mms)                              # [ws, S, 2]
    minmax = torch.stack((mm[:, :, 0].min(0).values, mm[:, :, 1].max(0).values), 1).contiguous()
    return rows.contiguous(), minmax


def share_selected(sel_mine: torch.Tensor) -> torch.Tensor:
    """sel_mine [S, 4] int32 with the rows of series this rank does not own zeroed -> full table."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(sel_mine, op=dist.ReduceOp.SUM)
    return sel_mine
