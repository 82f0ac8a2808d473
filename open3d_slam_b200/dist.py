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


# ----------------------------------------------------------------------------------------------------------------------
# shared registration targets: a submap that several ranks register against is built ONCE, by its owner,
# and broadcast -- NCCL over NVLink on GPUs (gloo in the CPU tests).  The payload is the submap's voxel content
# {point, normal per voxel} = 48 bytes per voxel; every receiver builds its own NN index from it (an index is ~2x the
# bytes of the cloud it indexes, and its build is a handful of launches).
# ----------------------------------------------------------------------------------------------------------------------
def owner_of(unit: int, world: int) -> int:
    """Round-robin ownership of shared units (targets): unit u is built by rank u % world."""
    return unit % world


def broadcast_point_sets(local: dict, n_units: int, world: int, rank: int, device=None, needed=None):
    """local: {unit: (xyz (n,3) float64 tensor, nrm (n,3) float64 tensor)} for the units THIS rank owns (owner_of), tensors on
    `device`.  Returns ({unit: (xyz, nrm)} for every unit in `needed` (default: all units), bytes this rank received).
    One size all_gather, then one broadcast per unit straight out of / into the tensors -- no staging copy."""
    import torch
    import torch.distributed as dist
    needed = set(range(n_units)) if needed is None else set(needed)
    if world == 1:
        return {u: local[u] for u in needed}, 0
    sizes = torch.zeros(n_units, dtype=torch.int64, device=device)
    for u, (x, _n) in local.items():
        sizes[u] = x.shape[0]
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM)      # every unit has exactly one owner
    sizes = sizes.cpu().tolist()
    out, received = {}, 0
    for u in range(n_units):
        src = owner_of(u, world)
        n = int(sizes[u])
        if src == rank:
            x, nr = local[u]
            buf = torch.stack([x.reshape(-1), nr.reshape(-1)]) if n else torch.zeros((2, 0), dtype=torch.float64, device=device)
        else:
            buf = torch.empty((2, 3 * n), dtype=torch.float64, device=device)
        dist.broadcast(buf, src=src)
        if src != rank:
            received += buf.numel() * 8
        if u in needed:
            out[u] = (buf[0].reshape(n, 3), buf[1].reshape(n, 3))
    return out, received
