"""Multi-GPU host logic.

Catalogue mode: haloes are independent units; they are partitioned by cost
w_h = N_h * (N_h + N_ext,h) with greedy LPT (largest first onto the least-loaded rank), each
rank runs its own batched plan, and there is no collective during compute -- only a gather
of the per-halo results at the end.

Split mode (one giant halo): every rank holds the full particle set; target groups are
dealt round-robin to the ranks and the potentials are combined by one all-reduce per pass
inside the plan (halma_plan_join, csrc/api.cu).  `split_owner` restates the device-side
ownership rule for tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def halo_costs(offsets, ext_offsets: Sequence = ()) -> np.ndarray:
    n = np.diff(np.asarray(offsets, dtype=np.int64)).astype(np.float64)
    ext = np.zeros_like(n)
    for e in ext_offsets:
        ext += np.diff(np.asarray(e, dtype=np.int64))
    return n * (n + ext)


def lpt_partition(costs, n_ranks: int) -> List[np.ndarray]:
    """Greedy longest-processing-time partition.  Deterministic (ties by halo id), so every
    rank computes the same answer without communication.  Returns ascending halo ids per rank."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.lexsort((np.arange(len(costs)), -costs))
    load = np.zeros(n_ranks)
    out: List[list] = [[] for _ in range(n_ranks)]
    for h in order:
        r = int(np.argmin(load))          # first minimum: deterministic
        out[r].append(int(h))
        load[r] += costs[h]
    return [np.array(sorted(ids), dtype=np.int64) for ids in out]


def partition_imbalance(costs, parts) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    loads = np.array([np.sum(np.asarray(costs)[p]) for p in parts], dtype=np.float64)
    return float(loads.max() / loads.mean()) if loads.mean() > 0 else 1.0


def take_haloes(offsets, arrays: Sequence[np.ndarray], halo_ids):
    """Sub-catalogue (offsets, arrays) holding only `halo_ids`, in that order."""
    offsets = np.asarray(offsets, dtype=np.int64)
    halo_ids = np.asarray(halo_ids, dtype=np.int64)
    sizes = (offsets[1:] - offsets[:-1])[halo_ids]
    new_off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    if len(halo_ids):
        take = np.concatenate([np.arange(offsets[h], offsets[h + 1]) for h in halo_ids])
    else:
        take = np.zeros(0, np.int64)
    return new_off, [np.asarray(a)[take] for a in arrays]


def split_owner(n_members: int, group_size: int, n_ranks: int) -> np.ndarray:
    """Rank that evaluates each current member position in split mode: target groups of
    `group_size` consecutive members are dealt round-robin (potential.cu decode_ticket,
    loop_kernels.cu k_fold_partials)."""
    return (np.arange(n_members) // group_size) % n_ranks


def gather_catalogue(local_halo_ids, local_rows: np.ndarray, n_halo: int, group=None) -> Optional[np.ndarray]:
    """Gather per-halo result rows (float64[n_local, k]) to rank 0 in global halo order.
    Uses torch.distributed (NCCL or gloo); returns None on the other ranks."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    payload = (np.asarray(local_halo_ids, dtype=np.int64), np.asarray(local_rows, dtype=np.float64))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if rank != 0:
        return None
    k = payload[1].shape[1] if payload[1].ndim == 2 else 1
    out = np.full((n_halo, k), np.nan)
    seen = np.zeros(n_halo, bool)
    for ids, rows in gathered:
        rows = rows.reshape(len(ids), k)
        if np.any(seen[ids]):
            raise RuntimeError("a halo was processed by two ranks")
        seen[ids] = True
        out[ids] = rows
    if not seen.all():
        raise RuntimeError("%d haloes were not processed by any rank" % int((~seen).sum()))
    del torch
    return out
