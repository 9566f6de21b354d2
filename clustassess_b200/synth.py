"""Seeded synthetic labelings for the ECS hot path.

Two families, both int32, 1-based like ``as.integer(factor)`` in the reference's pipeline
(R/stability-based-parameter-assessment.R:44):

* Louvain-like: one base partition with ``k0`` clusters whose sizes follow a normalised
  log-normal(sigma=1); each of the M partitions is the base with random cluster merges / splits so that
  K_m lands in [k_lo, k_hi], a fraction p_m ~ U(0.01, 0.15) of cells reassigned to a uniformly random
  cluster, and a random label permutation.
* Uniform: i.i.d. uniform labels in 1..K_m (dense tables, no locality) -- the reported worst case.

Every partition is generated from its own ``PCG64([seed, m])`` stream, so any subset of partitions can be
regenerated independently (the bench's CPU sample and the multi-GPU ranks rely on that).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

__all__ = ["louvain_like", "uniform_random", "config", "config_weights", "CONFIGS"]


def _base(n, k0, seed):
    rng = np.random.Generator(np.random.PCG64([seed, 1 << 20]))
    sizes = rng.lognormal(mean=0.0, sigma=1.0, size=k0)
    cdf = np.cumsum(sizes / sizes.sum())
    cdf[-1] = 1.0
    return np.searchsorted(cdf, rng.random(n), side="right").astype(np.int32)


def _compact(lab):
    cnt = np.bincount(lab)
    rank = np.cumsum(cnt > 0) - 1
    return rank[lab].astype(np.int32), int(rank[-1]) + 1


def _one_louvain(base, k0, k_lo, k_hi, seed, m):
    rng = np.random.Generator(np.random.PCG64([seed, m]))
    n = base.size
    k_m = int(rng.integers(k_lo, k_hi + 1))
    if k_m <= k0:
        # merges: k_m groups keep one base cluster each, the other k0-k_m base clusters join a random group
        order = rng.permutation(k0)
        grp = np.empty(k0, dtype=np.int64)
        grp[order[:k_m]] = np.arange(k_m)
        grp[order[k_m:]] = rng.integers(0, k_m, size=k0 - k_m)
        lab = grp[base]
    else:
        # splits: k_m-k0 extra pieces handed to base clusters with probability proportional to size,
        # every split cluster cut at sorted uniform break points
        csize = np.bincount(base, minlength=k0).astype(np.float64)
        extra = rng.multinomial(k_m - k0, csize / csize.sum())
        first = np.concatenate(([0], np.cumsum(extra + 1)[:-1]))
        thr_cl = np.repeat(np.arange(k0), extra)
        thr = np.sort(thr_cl + rng.random(thr_cl.size))
        thr_off = np.concatenate(([0], np.cumsum(extra)[:-1]))
        key = base + rng.random(n)
        piece = np.searchsorted(thr, key, side="right") - thr_off[base]
        lab = first[base] + piece
    k_now = int(lab.max()) + 1
    p_m = rng.uniform(0.01, 0.15)
    hit = rng.random(n) < p_m
    lab = lab.copy()
    lab[hit] = rng.integers(0, k_now, size=int(hit.sum()))
    lab, k_fin = _compact(lab)
    perm = rng.permutation(k_fin).astype(np.int32)
    return perm[lab] + np.int32(1)


def louvain_like(m, n, k0, k_lo, k_hi, seed, parts=None, threads=8):
    """M x N int32 label matrix (partition-major).  ``parts`` = iterable of partition indices to
    generate (default all M); rows come back in that order."""
    base = _base(n, k0, seed)
    parts = list(range(m)) if parts is None else list(parts)
    out = np.empty((len(parts), n), dtype=np.int32)

    def work(t):
        out[t] = _one_louvain(base, k0, k_lo, k_hi, seed, parts[t])

    if threads > 1 and len(parts) > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(work, range(len(parts))))
    else:
        for t in range(len(parts)):
            work(t)
    return out


def uniform_random(m, n, k_lo, k_hi, seed, parts=None):
    parts = list(range(m)) if parts is None else list(parts)
    out = np.empty((len(parts), n), dtype=np.int32)
    for t, p in enumerate(parts):
        rng = np.random.Generator(np.random.PCG64([seed, p]))
        k_m = int(rng.integers(k_lo, k_hi + 1))
        out[t] = rng.integers(1, k_m + 1, size=n, dtype=np.int32)
    return out


# The five BASELINE.json configurations.
CONFIGS = {
    "C1": dict(m=30, n=2_700, k0=9, k_lo=7, k_hi=12, seed=1),
    "C2": dict(m=2, n=100_000, k0=20, k_lo=20, k_hi=20, seed=2),
    "C3": dict(m=100, n=100_000, k0=25, k_lo=15, k_hi=40, seed=3),
    "C4": dict(m=1_000, n=50_000, k0=30, k_lo=10, k_hi=60, seed=4),
    "C5": dict(m=500, n=1_000_000, k0=1_500, k_lo=500, k_hi=2_000, seed=5),
}


def config(name, family="louvain", m=None, n=None, parts=None, threads=8):
    """Labels (and weights, for C3) of a named configuration; ``m``/``n`` override the size."""
    c = dict(CONFIGS[name])
    if m is not None:
        c["m"] = m
    if n is not None:
        c["n"] = n
    weights = None
    if name == "C2" and family == "louvain":
        # one pair: 20 clusters vs the same partition with 15 splits + 5 % noise (35 clusters)
        base = _base(c["n"], 20, c["seed"])
        rng = np.random.Generator(np.random.PCG64([c["seed"], 1]))
        a = base + np.int32(1)
        csize = np.bincount(base, minlength=20).astype(np.float64)
        extra = rng.multinomial(15, csize / csize.sum())
        first = np.concatenate(([0], np.cumsum(extra + 1)[:-1]))
        thr_cl = np.repeat(np.arange(20), extra)
        thr = np.sort(thr_cl + rng.random(thr_cl.size))
        thr_off = np.concatenate(([0], np.cumsum(extra)[:-1]))
        piece = np.searchsorted(thr, base + rng.random(c["n"]), side="right") - thr_off[base]
        b = (first[base] + piece).astype(np.int32)
        hit = rng.random(c["n"]) < 0.05
        b[hit] = rng.integers(0, 35, size=int(hit.sum()))
        labels = np.stack([a.astype(np.int32), b + np.int32(1)])
        return labels, None
    if family == "louvain":
        labels = louvain_like(c["m"], c["n"], c["k0"], c["k_lo"], c["k_hi"], c["seed"], parts=parts, threads=threads)
    elif family == "uniform":
        labels = uniform_random(c["m"], c["n"], c["k_lo"], c["k_hi"], c["seed"], parts=parts)
    else:
        raise ValueError("family must be 'louvain' or 'uniform'")
    weights = config_weights(name, c["m"])
    if weights is not None and parts is not None:
        weights = weights[list(parts)]
    return labels, weights


def config_weights(name, m=None):
    """Integer frequency weights 1 + Poisson(1) for C3 (mirrors `freq` of the unique partitions at one
    resolution, R/stability-3-graph-clustering.R:339-341); None for the unweighted configurations."""
    if name != "C3":
        return None
    c = CONFIGS[name]
    rng = np.random.Generator(np.random.PCG64([c["seed"], 1 << 21]))
    return (1 + rng.poisson(1.0, size=m or c["m"])).astype(np.float64)
