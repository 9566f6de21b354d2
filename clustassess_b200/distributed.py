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

__all__ = ["shard_owner", "owned_partitions", "consistency_sharded", "upload_sharded"]


def shard_owner(i: int, world: int) -> int:
    """Rank that owns partition i (and with it all pairs (i, j > i)); mirrors owner_of() in ca_api.cu."""
    r = i % (2 * world)
    return r if r < world else 2 * world - 1 - r


def owned_partitions(m: int, rank: int, world: int):
    return [i for i in range(m - 1) if shard_owner(i, world) == rank]


def upload_sharded(host_labels, device_labels, group=None):
    """Host -> device upload of the M x N int32 label matrix, sharded over the ranks: every rank copies only its
    contiguous block of partitions over PCIe (from pinned host memory) and the blocks are exchanged with ONE NCCL
    all-gather over NVLink, so the whole job moves the matrix across PCIe once instead of once per GPU.

    host_labels:   (M, N) int32 torch tensor in (pinned) host memory -- every rank sees the same data.
    device_labels: (M, N) int32 CUDA tensor, filled in place on every rank.
    M is split into world blocks of ceil(M / world) partitions; the tail block is padded inside a scratch buffer.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    m, n = host_labels.shape
    if world == 1:
        device_labels.copy_(host_labels, non_blocking=True)
        return device_labels
    blk = -(-m // world)
    lo, hi = min(rank * blk, m), min((rank + 1) * blk, m)
    if m % world == 0:
        mine = device_labels[lo:hi]
        mine.copy_(host_labels[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(device_labels.view(-1), mine.reshape(-1), group=group)
        return device_labels
    full = torch.empty((world * blk, n), dtype=device_labels.dtype, device=device_labels.device)
    mine = full[rank * blk:(rank + 1) * blk]
    if hi > lo:
        mine[: hi - lo].copy_(host_labels[lo:hi], non_blocking=True)
    dist.all_gather_into_tensor(full.view(-1), mine.reshape(-1).clone(), group=group)
    device_labels.copy_(full[:m])
    return device_labels


def consistency_sharded(n, m, weights, partial_fn, finish_fn, group=None, want_sim=False, device=None, fixed_point=False):
    """Run one rank's share and combine.

    partial_fn(rank, world, ecc_tensor, sim_tensor_or_None): accumulates this rank's partial sums into the
        zeroed float64 tensors (GPU: _capi.DeviceLabels.consistency_partial on their data_ptr()).
    finish_fn(ecc_tensor, sim_tensor_or_None): adds the identical-pair constant, mirrors the matrix.
    fixed_point: the partial sums are the library's 64-bit fixed-point integers (ca_consistency_partial_dev): they
        are summed over the ranks AS INTEGERS, which makes the result independent of the reduction order
        (bit-identical for any world size); finish_fn converts them to doubles.
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
        dist.all_reduce(ecc.view(torch.int64) if fixed_point else ecc, op=dist.ReduceOp.SUM, group=group)
        if sim is not None:
            dist.all_reduce(sim.view(torch.int64) if fixed_point else sim, op=dist.ReduceOp.SUM, group=group)
        if ecc.is_cuda:
            torch.cuda.current_stream().synchronize()
    finish_fn(ecc, sim)
    return ecc, (None if sim is None else sim.view(m, m))
