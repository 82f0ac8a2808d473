"""Multi-GPU plumbing for the batched paths: registrations are independent units, so a batch is split by
index across the ranks with NO data-path collective; only results (19 scalars per registration) and timings travel.
One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, world: int, rank: int) -> range:
    """Contiguous, balanced split of n independent units: rank r gets [lo, hi); sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return range(lo, hi)


def gather_results(local: np.ndarray, n_total: int, world: int, rank: int, device=None) -> np.ndarray:
    """All ranks obtain the (n_total, width) result table from their shard_range() slices (padded all_gather)."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, dtype=np.float64).reshape(len(local), -1)
    width = local.shape[1] if local.size else 0
    if world == 1:
        return local
    w = torch.tensor([width], dtype=torch.int64, device=device)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    width = int(w.item())
    per = (n_total + world - 1) // world
    buf = torch.zeros((per, width), dtype=torch.float64, device=device)
    if len(local):
        buf[:len(local)] = torch.from_numpy(local.reshape(len(local), width)).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = [out[r][:len(shard_range(n_total, world, r))].cpu().numpy() for r in range(world)]
    return np.vstack(parts) if parts else np.zeros((0, width))


def max_over_ranks(value: float, world: int, device=None) -> float:
    """Device-timed durations are reported as the max over ranks (never wall clock, never the mean)."""
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
