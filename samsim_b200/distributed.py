"""Multi-GPU plumbing: one process per GPU, columns sharded, NO collective on the stepping path.

The only exchange SAMSIM-style ensembles need is the optional reduction / gather of diagnostics at output cadence
(SURVEY section 8e): 18 numbers per rank (sum/min/max of ice thickness, bulk salinity, freeboard, snow depth,
surface temperature, N_active) -> one all-reduce each for SUM / MIN / MAX over NCCL (gloo on CPU in the tests).
"""
from __future__ import annotations

import numpy as np

NAMES = ["thickness", "bulk_salin", "freeboard", "thick_snow", "T_top", "N_active"]


def shard(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of columns of `rank`: [col0, col0 + n); the remainder goes to the first ranks."""
    base, rem = divmod(total, world)
    n = base + (1 if rank < rem else 0)
    col0 = rank * base + min(rank, rem)
    return col0, n


def reduce_ensemble(local: dict, ncol_local: int, group=None, device=None) -> dict:
    """local = Engine.reduce_diag() of this rank; returns the ensemble-wide mean/min/max per diagnostic."""
    import torch
    import torch.distributed as dist
    s = torch.tensor([local[n]["sum"] for n in NAMES] + [float(ncol_local)], dtype=torch.float64, device=device)
    mn = torch.tensor([local[n]["min"] for n in NAMES], dtype=torch.float64, device=device)
    mx = torch.tensor([local[n]["max"] for n in NAMES], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    total = float(s[-1].item())
    return {n: {"mean": float(s[j].item()) / total, "min": float(mn[j].item()), "max": float(mx[j].item())} for j, n in enumerate(NAMES)} | {"columns": int(total)}


def gather_columns(values, group=None, dst: int = 0):
    """Gather a per-column diagnostic (1-D tensor on this rank's device) to rank `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return values
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=values.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([values.numel()], dtype=torch.int64, device=values.device), group=group)
    out = [torch.empty(int(sz.item()), dtype=values.dtype, device=values.device) for sz in sizes]
    dist.all_gather(out, values, group=group) if len({int(sz.item()) for sz in sizes}) == 1 else _uneven(out, values, group)
    return torch.cat(out) if dist.get_rank(group) == dst else None


def _uneven(out, values, group):
    import torch.distributed as dist
    for r, buf in enumerate(out):  # r = rank inside `group`; broadcast wants the GLOBAL rank of the source
        if r == dist.get_rank(group):
            buf.copy_(values)
        src = dist.get_global_rank(group, r) if group is not None else r
        dist.broadcast(buf, src=src, group=group)
