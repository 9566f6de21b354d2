"""Pair-sharded all-pairs ECS over several GPUs: one process per GPU, torch.distributed for plumbing.

Partitions are dealt to ranks in zigzag order (0..w-1, w-1..0, ...), which balances the number of
pairs (i, j>i) per rank; every rank runs the wave kernel over its share of the work items on its own GPU and produces a
partial per-cell sum (and, optionally, its rows of the ECS matrix).  The only exchange step of the whole
path is ONE all-reduce of the N-vector (8 MB at N = 1M) and, when the matrix is wanted, one of the M x M
buffer -- there is no data-path collective inside the kernels because pairs are independent
(SURVEY.md section 8e).  The reference's analogue is foreach %dopar% over pairs (R/ECS.R:953-959).
"""
from __future__ import annotations

import numpy as np

__all__ = ["shard_owner", "owned_partitions", "consistency_sharded"]


def shard_owner(i: int, world: int) -> int:
    """Rank that owns partition i (and with it all pairs (i, j > i)); mirrors owner_of() in ca_api.cu."""
    r = i % (2 * world)
    return r if r < world else 2 * world - 1 - r


def owned_partitions(m: int, rank: int, world: int):
    return [i for i in range(m - 1) if shard_owner(i, world) == rank]


def consistency_sharded(n, m, weights, partial_fn, finish_fn, group=None, want_sim=False, device=None):
    """Run one rank's share and combine.

    partial_fn(rank, world, ecc_tensor, sim_tensor_or_None): accumulates this rank's partial sums into the
        zeroed float64 tensors (GPU: _capi.DeviceLabels.consistency_partial on their data_ptr()).
    finish_fn(ecc_tensor, sim_tensor_or_None): adds the identical-pair constant, mirrors the matrix.
    Returns (ecc, sim) tensors, identical on every rank.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ecc = torch.zeros(n, dtype=torch.float64, device=device)
    sim = torch.zeros(m * m, dtype=torch.float64, device=device) if want_sim else None
    if ecc.is_cuda:
        torch.cuda.current_stream().synchronize()  # the zero fills must land before the library's stream adds
    partial_fn(rank, world, ecc, sim)
    if world > 1:
        dist.all_reduce(ecc, op=dist.ReduceOp.SUM, group=group)
        if sim is not None:
            dist.all_reduce(sim, op=dist.ReduceOp.SUM, group=group)
        if ecc.is_cuda:
            torch.cuda.current_stream().synchronize()
    finish_fn(ecc, sim)
    return ecc, (None if sim is None else sim.view(m, m))
