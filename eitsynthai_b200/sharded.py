"""Slice-sharded execution of a batch of series across the GPUs of one box.

Every stage of the hot path is per-slice except the coronal image (one row per slice plus a
global min/max), so a series is cut into contiguous z-ranges, one per rank, and
the only exchange is

  * one all-gather (``all_gather_into_tensor``) of the per-slice coronal rows (n_local x W int16 per series
    per rank) together with the per-shard (min, max) (2 int32 per series per rank)
  * one all-reduce of the selected slice indices (4 int32 per series; series s is decided on rank s % world)

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


def _flip(rows: torch.Tensor, flip_z) -> torch.Tensor:
    for s in flip_z or ():
        rows[s] = rows[s].flip(0)
    return rows


def gather_rows(rows_local: torch.Tensor, minmax_local: torch.Tensor, n_slices: int, flip_z=None):
    """rows_local [S, n_local, W] int16 (this rank's z-range, ascending z), minmax_local [S, 2] int32 ->
    (rows [S, n_slices, W], minmax [S, 2]) identical on every rank.  ONE collective: rows and min/max travel
    in the same byte buffer (neither NCCL nor gloo has a 16-bit integer type).

    ``flip_z``: series whose coronal image is z-reversed (FFS, or PatientOrientation[1] == 'P' when not HFS;
    utils.py:130-132,155-160).  The reversal is applied to the GATHERED rows: reversing inside each shard and
    concatenating in rank order would give [rev(shard0), rev(shard1), ...] instead."""
    rank, ws = world()
    if ws == 1:
        return _flip(rows_local, flip_z), minmax_local
    S, nl, W = rows_local.shape
    sizes = [shard_range(n_slices, ws, r)[1] - shard_range(n_slices, ws, r)[0] for r in range(ws)]
    nmax = max(sizes)
    row_bytes = S * nmax * W * 2
    send = torch.zeros((row_bytes + S * 8,), dtype=torch.uint8, device=rows_local.device)
    send[:row_bytes].view(torch.int16).view(S, nmax, W)[:, :nl] = rows_local          # ragged shards: padded to the largest
    send[row_bytes:].view(torch.int32).view(S, 2).copy_(minmax_local)
    recv = torch.empty((ws * (row_bytes + S * 8),), dtype=torch.uint8, device=rows_local.device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.view(ws, row_bytes + S * 8)
    parts = recv[:, :row_bytes].view(torch.int16).view(ws, S, nmax, W)
    if all(sz == nmax for sz in sizes):
        rows = parts.permute(1, 0, 2, 3).reshape(S, ws * nmax, W)
    else:
        rows = torch.cat([parts[r, :, :sizes[r]] for r in range(ws)], 1)
    mm = recv[:, row_bytes:].view(torch.int32).view(ws, S, 2)
    minmax = torch.stack((mm[:, :, 0].min(0).values, mm[:, :, 1].max(0).values), 1).contiguous()
    return _flip(rows.contiguous(), flip_z), minmax


def share_selected(sel_mine: torch.Tensor) -> torch.Tensor:
    """sel_mine [S, 4] int32 with the rows of series this rank does not own zeroed -> full table."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(sel_mine, op=dist.ReduceOp.SUM)
    return sel_mine
