"""CPU tests of the N > 1 path: world_size-2 `gloo` runs of the pair-sharded wrapper
(clustassess_b200/distributed.py) with an ORACLE-backed stand-in for the per-rank GPU call, plus the shard
assignment itself (Python mirror == the C ABI's ca_shard_owner, balance across ranks).

The stand-in follows the contract of ca_consistency_partial_dev / ca_consistency_finish_dev exactly
(include/clustassess_b200.h): a rank adds, for every partition i it owns and every j > i,
(w_i / norm) * w_j * ECS_ij into ecc and sum_n ECS_ij[n] into sim[i, j]; finish adds the identical-pair
constant and mirrors / normalises the matrix.  The GPU twin of this test is
tests/test_gpu_parity.py::test_rank_sharding_sums_to_the_whole."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from clustassess_b200 import _capi, synth
from clustassess_b200.distributed import consistency_sharded, owned_partitions, shard_owner, upload_sharded


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_partial(labels, weights):
    m, n = labels.shape
    w = np.ones(m) if weights is None else np.asarray(weights, dtype=np.float64)
    norm = w.sum() * (w.sum() - 1) / 2

    def partial(rank, world, ecc, sim):
        e = ecc.numpy()
        s = None if sim is None else sim.numpy().reshape(m, m)
        for i in owned_partitions(m, rank, world):
            for j in range(i + 1, m):
                ecs = oracle.disjoint_ecs(labels[i], labels[j])
                e += ecs * (w[i] / norm * w[j])
                if s is not None:
                    s[i, j] += ecs.sum()

    def finish(ecc, sim):
        ecc += float(np.sum(w / norm * (w - 1) / 2))
        if sim is not None:
            s = sim.view(m, m)
            up = torch.triu(s, diagonal=1) / n
            s.copy_(up + up.T + torch.eye(m, dtype=torch.float64))

    return partial, finish


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        labels, weights = synth.config("C3", m=9, n=3000)
        m, n = labels.shape
        partial, finish = _oracle_partial(labels, weights)
        # the sharded upload: every rank contributes its block of partitions, one all-gather rebuilds the matrix
        for mm in (m, m - 1):  # divisible and non-divisible by the world size
            host = torch.from_numpy(np.ascontiguousarray(labels[:mm]))
            dev = torch.full((mm, n), -1, dtype=torch.int32)
            upload_sharded(host, dev)
            assert torch.equal(dev, host), "upload_sharded must rebuild the label matrix on every rank"
        ecc, sim = consistency_sharded(n, m, weights, partial, finish, want_sim=True, device="cpu")
        ref_ecc, ref_sim = oracle.weighted_element_consistency(labels, weights, calculate_sim_matrix=True)
        q.put((rank, float(np.abs(ecc.numpy() - ref_ecc).max()), float(np.abs(sim.numpy() - ref_sim).max())))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_ecc, e_sim in res:  # every rank holds the full, identical result after the all-reduce
        assert e_ecc < 1e-9 and e_sim < 1e-9, (rank, e_ecc, e_sim)


def test_single_process_wrapper_without_process_group():
    labels, _ = synth.config("C1", m=6)
    m, n = labels.shape
    partial, finish = _oracle_partial(labels, None)
    ecc, sim = consistency_sharded(n, m, None, partial, finish, want_sim=False, device="cpu")
    assert sim is None
    assert np.abs(ecc.numpy() - oracle.weighted_element_consistency(labels)).max() < 1e-12


def test_shard_assignment_mirrors_the_c_abi_and_is_balanced():
    lib = _capi.lib()
    for world in (1, 2, 3, 4, 8):
        for i in range(64):
            assert lib.ca_shard_owner(i, world) == shard_owner(i, world)
    m = 500
    for world in (2, 4, 8):
        pairs = [sum(m - 1 - i for i in owned_partitions(m, r, world)) for r in range(world)]
        assert sum(pairs) == m * (m - 1) // 2
        assert max(pairs) / min(pairs) < 1.02  # zigzag keeps the pair counts within 2 %
        owned = sorted(i for r in range(world) for i in owned_partitions(m, r, world))
        assert owned == list(range(m - 1))
