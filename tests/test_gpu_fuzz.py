"""Randomised parity sweep of the batched engine against the CPU oracle (GPU, through the C ABI).

Every case draws the number of partitions, cells, the cluster-size law and the weights at random (seeded), so the
work-unit / phase / item machinery of the wave kernel sees shapes nobody wrote a dedicated test for: partition counts
across tile, super-tile and span boundaries, cluster sizes across the resident / two-chunk / long streaming limits,
many-row items, single-cluster partitions, label gaps and negative labels.  Tolerance: 1e-9 absolute (BASELINE.json).
"""
import numpy as np
import pytest

import oracle
from clustassess_b200 import ecs

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _partition(rng, n, kind):
    if kind == 0:      # uniform labels
        k = int(rng.integers(1, 400))
        x = rng.integers(0, k, size=n)
    elif kind == 1:    # log-normal cluster sizes: a few big clusters, many small ones
        k = int(rng.integers(2, 2000))
        p = rng.lognormal(0.0, 1.5, size=k)
        x = rng.choice(k, size=n, p=p / p.sum())
    elif kind == 2:    # one dominant cluster (streaming item) + noise
        x = np.where(rng.random(n) < rng.uniform(0.3, 0.95), 0, 1 + rng.integers(0, int(rng.integers(1, 50)), size=n))
    elif kind == 3:    # many tiny clusters (at most ~1500 of them: the CPU oracle builds dense K x K tables per pair)
        x = rng.permutation(n) // max(int(rng.integers(1, 4)), -(-n // 1500))
    else:              # a single cluster
        x = np.zeros(n, dtype=np.int64)
    # label gaps (only for small label sets: the reference, and the oracle after it, builds dense tables over the label
    # RANGE) and offsets, negative too: ECS is invariant under any bijective relabelling
    gap = int(rng.integers(1, 4)) if x.max() < 300 else 1
    return (x * gap + int(rng.integers(-50, 50))).astype(np.int32)


SHAPES = [(2, 40_000), (3, 20_000), (7, 7000), (8, 3457), (9, 900), (16, 130), (17, 20_000), (31, 7000), (32, 3457), (33, 900),
          (40, 2000), (66, 1500), (5, 1), (12, 5), (24, 40_000), (48, 900)]


# the routes through the wave kernel the host would otherwise pick by itself (csrc/ca_api.cu route_for): forced, so that every
# random shape also goes through the matrix-from-the-tables launches and through the ratio-table launches
ROUTES = {"own choice": {}, "matrix from the tables": {"CA_RT": "0", "CA_SIM_SCAN": "1"}, "ratio tables": {"CA_RT": "1"},
          "per cell": {"CA_RT": "0", "CA_SIM_SCAN": "0"}}


@pytest.mark.parametrize("route", list(ROUTES))
@pytest.mark.parametrize("seed", range(len(SHAPES)))
def test_random_label_sets(seed, route, monkeypatch):
    if route != "own choice" and seed % 2 == (0 if route == "per cell" else 1) and route != "matrix from the tables":
        pytest.skip("every other shape for the forced per-cell and ratio-table routes")
    for k, v in ROUTES[route].items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(1000 + seed)
    m, n = SHAPES[seed]
    base = _partition(rng, n, int(rng.integers(0, 5)))
    L = []
    for p in range(m):
        r = rng.random()
        if r < 0.4:    # a noisy copy of the base partition (the normal case: similar clusterings)
            x = np.where(rng.random(n) < rng.uniform(0.0, 0.3), _partition(rng, n, 0), base)
        elif r < 0.5:  # an exact copy under relabelling
            x = base.max() - base
        else:
            x = _partition(rng, n, int(rng.integers(0, 5)))
        L.append(x.astype(np.int32))
    L = np.stack(L)
    w = None if rng.random() < 0.4 else (1 + rng.poisson(1.0, size=m)).astype(np.float64)
    ecc, sim = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL, (m, n)
    assert np.abs(res["ecs_matrix"] - sim).max() < TOL, (m, n)
    only = ecs.weighted_element_consistency(list(L), weights=w)
    assert np.abs(only - ecc).max() < TOL
    assert np.abs(ecs.element_sim_matrix(list(L)) - sim).max() < TOL
    ref = int(rng.integers(0, m))
    agr = ecs.element_agreement(L[ref], [L[p] for p in range(m)])
    want = np.mean([oracle.disjoint_ecs(L[ref], L[p]) for p in range(m)], axis=0)
    assert np.abs(agr - want).max() < TOL
