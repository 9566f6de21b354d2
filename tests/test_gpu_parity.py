"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle / golden vectors.

Bars: contingency tables and cluster sizes bit-exact (int32); per-pair ECS
bit-exact (one IEEE fp64 division); ECC / weighted ECC / ECS matrix within 1e-9 absolute (fp64; only the
summation order differs from the reference).  The reference's own property tests
(tests/testthat/test-ECS.R) are restated on the GPU library further down.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from clustassess_b200 import _capi, ecs, synth

pytestmark = pytest.mark.gpu
TOL = 1e-9  # fp64 absolute tolerance for set-level results (north_star)


# ---- golden vectors generated from the unmodified reference C++ ------------------------------------------
def test_golden_pairs(golden):
    for k in range(int(golden["n_pairs"])):
        a, b = golden[f"p{k}_a"], golden[f"p{k}_b"]
        tab, sa, sb = ecs.contingency_table(a, b)
        gt = golden[f"p{k}_table"]
        assert np.array_equal(tab, gt), "table of golden pair %d" % k
        assert np.array_equal(sa, gt.sum(1)) and np.array_equal(sb, gt.sum(0))
        e = ecs.element_sim_elscore(a, b)
        assert np.array_equal(e, golden[f"p{k}_ecs"]), "per-element ECS of golden pair %d must be bit-identical" % k
        avg = float(golden[f"p{k}_avg"])
        if np.isnan(avg):  # the reference's 0/0 quirk for labels unused in both ranges (see test_oracle.py)
            avg = float(golden[f"p{k}_ecs"].mean())
        assert abs(ecs.element_sim(a, b) - avg) < 1e-12


def test_golden_set_level(golden):
    L, w = golden["set_labels"], golden["set_weights"]
    res = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - golden["set_ecc_w"]).max() < TOL
    assert np.abs(res["ecs_matrix"] - golden["set_sim"]).max() < TOL
    assert np.abs(ecs.element_sim_matrix(list(L)) - golden["set_sim"]).max() < TOL
    assert np.abs(ecs.weighted_element_consistency(list(L)) - golden["set_ecc_unw"]).max() < TOL
    assert np.abs(ecs.element_consistency(list(L)) - golden["set_ecc_unw"]).max() < TOL


def test_documented_known_answer(golden):
    # merge_partitions(list(c(1,1,2), c(2,2,2), c("B","B","A")), 1)  -- docs/reference/merge_partitions.html:117-142
    res = ecs.merge_partitions([[1, 1, 2], [2, 2, 2], ["B", "B", "A"]], 1)
    assert [p["freq"] for p in res["partitions"]] == [2, 1]
    assert list(res["partitions"][0]["mb"]) == [1, 1, 2] and list(res["partitions"][1]["mb"]) == [2, 2, 2]
    assert np.abs(np.asarray(res["ecc"]) - golden["doc_ecc"]).max() < 5e-8  # the doc prints 7 digits
    assert np.allclose(res["ecc"], [7 / 9, 7 / 9, 5 / 9], atol=1e-15)
    assert abs(res["partitions"][0]["avg_agreement"] - 0.2777778) < 5e-8


# ---- the reference's own tests (tests/testthat/test-ECS.R), restated ------------------------------------------
def _lists(n_lists=100, k=10, n=100, seed=1234):
    rng = np.random.default_rng(seed)
    return [rng.integers(1, k + 1, size=n) for _ in range(n_lists)]


def test_ecs_between_0_and_1():  # test-ECS.R:1-12
    xs = _lists(200)
    for x, y in zip(xs[::2], xs[1::2]):
        e = ecs.element_sim_elscore(x, y)
        assert 0 <= e.min() and e.max() <= 1


def test_type_invariance():  # test-ECS.R:18-32
    import pandas as pd

    xs = _lists(20)
    for x, y in zip(xs[::2], xs[1::2]):
        e1 = ecs.element_sim_elscore(x, y)
        e2 = ecs.element_sim_elscore(pd.Categorical(x), pd.Categorical(y))
        e3 = ecs.element_sim_elscore(x.astype(str), y.astype(str))
        e4 = ecs.element_sim_elscore(x.astype(np.float64), y.astype(np.float64))
        assert np.array_equal(e1, e2) and np.array_equal(e1, e3) and np.array_equal(e1, e4)
        assert np.array_equal(e1, oracle.corrected_l1_mb(x, y)) or np.abs(e1 - oracle.corrected_l1_mb(x, y)).max() < 1e-12


def test_element_sim_is_mean_of_elscore():  # test-ECS.R:34-44
    xs = _lists(40)
    for x, y in zip(xs[::2], xs[1::2]):
        assert abs(np.mean(ecs.element_sim_elscore(x, y)) - ecs.element_sim(x, y)) < 1e-12


def test_ecc_is_average_of_pairwise_ecs():  # test-ECS.R:98-120
    cl = _lists(100)
    expected = ecs.element_consistency(cl)
    acc = np.zeros(100)
    for i in range(0, 100):
        for j in range(i + 1, 100):
            acc += 2 * oracle.disjoint_ecs(cl[i], cl[j])
    assert np.abs(acc / (100 * 99) - expected).max() < TOL


def test_ecs_matrix():  # test-ECS.R:122-140
    cl = _lists(100)
    mat = ecs.element_sim_matrix(cl)
    assert np.array_equal(mat, mat.T) and np.all(np.diag(mat) == 1)
    rng = np.random.default_rng(0)
    for _ in range(60):
        i, j = sorted(rng.choice(100, size=2, replace=False))
        assert abs(mat[i, j] - ecs.element_sim_elscore(cl[i], cl[j]).mean()) < TOL
    df = ecs.element_sim_matrix(cl[:5], output_type="data.frame")  # R/ECS.R:967-978
    assert list(df["i.clustering"][:6]) == [1, 2, 3, 4, 5, 1] and list(df["j.clustering"][:6]) == [1, 1, 1, 1, 1, 2]
    assert abs(df["element_sim"][5 * 1 + 3] - mat[3, 1]) < 1e-15


def test_merge_partitions_properties():  # test-ECS.R:142-188
    rng = np.random.default_rng(5)
    base = [rng.integers(1, 4, size=60) for _ in range(6)]
    cl = [base[i % 6] if i % 3 else (base[i % 6] % 3) * 7 + 2 for i in range(30)]  # many relabelled duplicates
    for thresh in (1, 0.99, 0.5):
        res = ecs.merge_partitions(cl, ecs_thresh=thresh, return_ecs_matrix=True)
        parts = res["partitions"]
        assert sum(p["freq"] for p in parts) == 30
        for p in parts:
            assert any(np.array_equal(p["mb"], c) for c in cl)
        mbs = [p["mb"] for p in parts]
        if len(mbs) > 1:
            w = [p["freq"] for p in parts]
            assert np.abs(res["ecc"] - ecs.weighted_element_consistency(mbs, weights=w)).max() < TOL
            # The reference permutes only the leading tie block of the matrix (R/ECS.R:1298-1306), so with a
            # PARTIAL tie the off-diagonal blocks keep the pre-tie order; that quirk is kept, and the
            # comparison is made where the reference itself is consistent (its own test has all ties).
            f = [p["freq"] for p in parts]
            t = next((k for k in range(1, len(f)) if f[k] != f[0]), len(f))
            d = np.abs(res["ecs_matrix"] - ecs.element_sim_matrix(mbs))
            assert d[:t, :t].max() < TOL and (t == len(f) or d[t:, t:].max() < TOL)
            if t in (1, len(f)):
                assert d.max() < TOL
        f = [p["freq"] for p in parts]
        assert all(f[i] >= f[i + 1] for i in range(len(f) - 1))
        res2 = ecs.merge_partitions(cl, ecs_thresh=thresh, order_logic="avg_agreement")
        ag = [p["avg_agreement"] for p in res2["partitions"]]
        assert all(ag[i] >= ag[i + 1] for i in range(len(ag) - 1))
    keep, freq = oracle.merge_identical_partitions(np.stack(cl))
    got = ecs.merge_identical_partitions([{"mb": c, "freq": 1} for c in cl])
    assert [g["freq"] for g in got] == freq.tolist()
    assert all(np.array_equal(g["mb"], cl[k]) for g, k in zip(got, keep))


def test_merge_matches_recomputed_ecc_and_matrix():  # test-ECS.R:156-170, same shape of input
    cl = _lists(100)
    merged = ecs.merge_partitions(cl, return_ecs_matrix=True)
    plist = [p["mb"] for p in merged["partitions"]]
    assert len(plist) == 100
    assert np.abs(merged["ecc"] - ecs.element_consistency(plist)).max() < TOL
    assert np.abs(merged["ecs_matrix"] - ecs.element_sim_matrix(plist)).max() < TOL


def test_dedup_reference_quirks():
    # label gap inside the range: never identical, even to itself, unless range == n (R/ECS.R:994-1015)
    cl = [[1, 3, 3, 3], [1, 3, 3, 3], [1, 2, 2, 2], [5, 4, 4, 4], [1, 3, 3, 3]]
    keep, freq = oracle.merge_identical_partitions(np.array(cl))
    got = ecs.merge_identical_partitions([{"mb": c, "freq": 1} for c in cl])
    assert [g["freq"] for g in got] == freq.tolist() and len(got) == keep.size
    cl = [[1, 1, 3], [1, 3, 3], [2, 2, 1]]  # range == n: the reference calls these identical
    keep, freq = oracle.merge_identical_partitions(np.array(cl))
    got = ecs.merge_identical_partitions([{"mb": c, "freq": 1} for c in cl])
    assert [g["freq"] for g in got] == freq.tolist() and len(got) == keep.size
    # ecc is invariant to all of this
    cl = [[1, 3, 3, 3, 9], [1, 3, 3, 3, 9], [1, 2, 2, 2, 2], [5, 4, 4, 4, 4]]
    assert np.abs(ecs.element_consistency(cl) - oracle.element_consistency(np.array(cl))).max() < TOL


# ---- BASELINE.json configurations against the oracle ------------------------------------------------------------
def test_config1_full():
    L, _ = synth.config("C1")
    assert np.abs(ecs.element_consistency(list(L)) - oracle.element_consistency(L)).max() < TOL
    res = ecs.weighted_element_consistency(list(L), calculate_sim_matrix=True)
    ecc, sim = oracle.weighted_element_consistency(L, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL


def test_config2_full():
    (a, b), _ = synth.config("C2")
    tab, sa, sb = ecs.contingency_table(a, b)
    assert tab.shape == (20, 35) and np.array_equal(tab, oracle.cont_table(a, b))
    assert np.array_equal(sa, np.bincount(a)[1:]) and np.array_equal(sb, np.bincount(b)[1:])
    e = ecs.element_sim_elscore(a, b, alpha=0.9)
    assert np.array_equal(e, oracle.disjoint_ecs(a, b))
    assert abs(ecs.element_sim(a, b) - oracle.disjoint_ecs_average(a, b)) < 1e-12


def test_config3_reduced_weighted():
    L, w = synth.config("C3", m=12)
    res = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
    ecc, sim = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL


def test_config4_reduced_matrix():
    L, _ = synth.config("C4", m=24)
    assert np.abs(ecs.element_sim_matrix(list(L)) - oracle.element_sim_matrix(L)).max() < TOL


@pytest.mark.parametrize("family", ["louvain", "uniform"])
def test_config5_reduced(family):
    L, _ = synth.config("C5", m=6, n=400_000, family=family)
    res = ecs.weighted_element_consistency(list(L), calculate_sim_matrix=True)
    ecc, sim = oracle.weighted_element_consistency(L, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL
    # the global-table contingency kernel (ka*kb beyond shared memory) is bit-exact too
    tab, sa, sb = ecs.contingency_table(L[0], L[1])
    assert np.array_equal(tab, oracle.cont_table(L[0], L[1]))
    assert np.array_equal(ecs.element_sim_elscore(L[0], L[1]), oracle.disjoint_ecs(L[0], L[1]))


def test_element_agreement():
    L, _ = synth.config("C3", m=7, n=30_000)
    got = ecs.element_agreement(L[0], list(L[1:]))
    want = np.mean([oracle.corrected_l1_mb(L[0], L[p]) for p in range(1, 7)], axis=0)  # R/ECS.R:1639-1647
    assert np.abs(got - want).max() < TOL


# ---- edge cases ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 5, 127, 2049, 4099])
def test_ragged_sizes(n):
    rng = np.random.default_rng(n)
    L = rng.integers(-3, 6, size=(5, n)).astype(np.int32)
    ecc, sim = oracle.weighted_element_consistency(L, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL
    assert np.array_equal(ecs.element_sim_elscore(L[0], L[1]), oracle.disjoint_ecs(L[0], L[1]))


def test_two_partitions_single_cluster_and_singletons():
    n = 3000
    one = np.ones(n, dtype=np.int32)
    single = np.arange(n, dtype=np.int32) + 10
    rng = np.random.default_rng(3)
    mid = rng.integers(0, 40, size=n).astype(np.int32)
    L = np.stack([one, single, mid, single[::-1].copy()])
    ecc, sim = oracle.weighted_element_consistency(L, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL
    two = ecs.weighted_element_consistency([one, mid])
    assert np.abs(two - oracle.weighted_element_consistency(np.stack([one, mid]))).max() < TOL


def test_label_gaps_negative_labels_and_weights():
    rng = np.random.default_rng(11)
    n = 5000
    L = np.stack([rng.choice([-7, -2, 0, 5, 900], size=n), rng.choice([3, 4, 10_000], size=n),
                  rng.integers(100, 140, size=n), rng.choice([1, 65_000], size=n)]).astype(np.int32)
    w = np.array([3.0, 1.0, 2.0, 5.0])
    ecc, sim = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL
    tab, _, _ = ecs.contingency_table(L[0], L[2])
    assert np.array_equal(tab, oracle.cont_table(L[0], L[2]))  # empty rows/cols kept, like the reference


def test_huge_cluster_wide_counters_and_streaming_items():
    rng = np.random.default_rng(13)
    n = 150_000  # two clusters of ~75k cells: beyond 16-bit counters, processed as streaming items
    L = np.stack([rng.integers(0, 2, size=n), rng.integers(0, 3, size=n), rng.integers(0, 1500, size=n),
                  (np.arange(n) < 70_000).astype(np.int64)]).astype(np.int32)
    ecc, sim = oracle.weighted_element_consistency(L, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL


def test_many_clusters_one_block_per_sm_configuration():
    rng = np.random.default_rng(17)
    n = 60_000
    L = np.stack([rng.integers(0, 9000, size=n), rng.integers(0, 20, size=n), rng.integers(0, 6000, size=n)]).astype(np.int32)
    ecc, sim = oracle.weighted_element_consistency(L, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL


def test_cluster_counts_up_to_the_table_row_limit_and_beyond():
    """A partner with 24 000 clusters needs 96 KB per table row: one row per item, one partner per phase.  One with
    30 000 does not fit a table buffer at all, one with 70 000 not even the uint16 labels: the reference has no such limit
    (src/calculate_ecs.cpp:7-19), so the library loops over the pairs with a hashed contingency table instead -- same
    results, same entry points."""
    rng = np.random.default_rng(29)
    n = 120_000
    L = np.stack([rng.integers(0, 24_000, size=n), rng.integers(0, 7, size=n), rng.integers(0, 300, size=n)]).astype(np.int32)
    ecc, sim = oracle.weighted_element_consistency(L, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL
    for k_big in (30_000, 70_000):
        L[0] = rng.integers(0, k_big, size=n)
        L[0][:k_big] = np.arange(k_big)
        w = np.array([2.0, 1.0, 3.0])
        ecc, sim = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
        res = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
        assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL
        assert np.abs(ecs.element_sim_matrix(list(L)) - oracle.element_sim_matrix(L)).max() < TOL
        got = ecs.element_agreement(L[0], [L[1], L[2]])
        want = (oracle.disjoint_ecs(L[0], L[1]) + oracle.disjoint_ecs(L[0], L[2])) / 2
        assert np.abs(got - want).max() < TOL
    # the per-pair entry points never had the limit
    assert np.array_equal(ecs.element_sim_elscore(L[0], L[2]), oracle.disjoint_ecs(L[0], L[2]))


def test_all_singleton_partitions_pass_the_dedup_pre_pass():
    """Partitions of n > 65535 singletons: are_identical_memberships returns TRUE through its nrow == length(mb) shortcut
    (R/ECS.R:999), so element_consistency merges them -- the pre-pass only needs ranks, whatever the cluster count."""
    n = 70_000
    rng = np.random.default_rng(31)
    a, b = np.arange(n, dtype=np.int32), rng.permutation(n).astype(np.int32)
    L = np.stack([a, b, a])
    assert np.array_equal(ecs.element_consistency(list(L)), np.ones(n))
    c = rng.integers(0, 50, size=n).astype(np.int32)
    L = np.stack([a, c, b])
    assert np.abs(ecs.element_consistency(list(L)) - oracle.element_consistency(L)).max() < TOL


def test_identical_partitions_and_m_not_multiple_of_tile():
    rng = np.random.default_rng(19)
    base = rng.integers(0, 12, size=4000).astype(np.int32)
    L = np.stack([base, base, (base + 5) % 12, rng.integers(0, 12, size=4000).astype(np.int32)] * 9 + [base])  # M = 37
    assert np.abs(ecs.element_consistency(list(L)) - oracle.element_consistency(L)).max() < TOL
    assert np.abs(ecs.weighted_element_consistency(list(L)) - oracle.weighted_element_consistency(L)).max() < TOL


def test_errors_through_the_c_abi():
    lib = _capi.lib()
    a = np.array([1, 2, np.iinfo(np.int32).min], dtype=np.int32)
    out = np.zeros(3)
    assert lib.ca_ecs_pair(a, a, 3, out.ctypes.data, None) == -4  # CA_E_NA_LABEL
    L = np.stack([a, a])
    assert lib.ca_consistency(L.reshape(-1), 3, 2, None, out, None) == -4
    assert lib.ca_consistency(L.reshape(-1), 3, 1, None, out, None) == _capi.CA_E_ARG
    b = np.array([1, 2, 3], dtype=np.int32)
    tab = np.zeros(4, dtype=np.int32)
    assert lib.ca_contingency(b, b, 3, 1, 1, 2, 2, tab, None, None) == -5  # CA_E_RANGE: label 3 outside [1, 3)
    res = ecs.weighted_element_consistency([np.arange(70_000), np.arange(70_000)])  # > 65535 clusters: the pair-loop fallback
    assert np.array_equal(res, np.ones(70_000))


def test_rank_sharding_sums_to_the_whole():
    """ca_consistency_partial_dev over the ranks of a world of 3, accumulated into the same buffers, equals
    the single-rank job (the allreduce is a plain sum)."""
    import torch

    L, w = synth.config("C3", m=11, n=20_000)
    m, n = L.shape
    ecc_ref, sim_ref = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    dl = _capi.DeviceLabels(n, m).upload(L)
    ecc = torch.zeros(n, dtype=torch.float64, device="cuda")
    sim = torch.zeros(m * m, dtype=torch.float64, device="cuda")
    for r in range(3):
        dl.consistency_partial(w, r, 3, ecc.data_ptr(), sim.data_ptr())
    _capi.consistency_finish(n, m, w, ecc.data_ptr(), sim.data_ptr())
    torch.cuda.synchronize()
    assert np.abs(ecc.cpu().numpy() - ecc_ref).max() < TOL
    assert np.abs(sim.cpu().numpy().reshape(m, m) - sim_ref).max() < TOL
    dl.close()


# ---- the wave engine's own structure (work units of 32 partners, phases, resident / streaming items) --------------
@pytest.mark.parametrize("m", [2, 15, 16, 17, 31, 33, 49, 64, 65, 97])
def test_super_tile_boundaries_weights_and_matrix(m):
    """Partition counts around the 8-partner tile, 16-partner super-tile and 32-partner work-unit boundaries, non-unit weights."""
    rng = np.random.default_rng(100 + m)
    n = 6000
    base = rng.integers(0, 30, size=n)
    L = np.stack([np.where(rng.random(n) < 0.2, rng.integers(0, 30 + p % 7, size=n), base) for p in range(m)]).astype(np.int32)
    w = (1 + rng.poisson(1.0, size=m)).astype(np.float64)
    ecc, sim = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL
    assert np.abs(ecs.weighted_element_consistency(list(L)) - oracle.weighted_element_consistency(L)).max() < TOL


def test_agreement_against_many_partitions():
    L, _ = synth.config("C3", m=41, n=20_000)
    got = ecs.element_agreement(L[0], list(L[1:]))
    want = np.mean([oracle.disjoint_ecs(L[0], L[p]) for p in range(1, 41)], axis=0)
    assert np.abs(got - want).max() < TOL


def test_mixed_item_kinds_in_one_job():
    """One partition with a cluster beyond a resident item (streaming), one beyond 65535 cells (32-bit counters),
    many tiny clusters (items with many table rows) and a partner with > 65535-cell clusters (size look-up path)."""
    rng = np.random.default_rng(23)
    n = 200_000
    big = np.where(np.arange(n) < 70_000, 0, 1 + rng.integers(0, 5, size=n))          # 70k cells in one cluster
    mid = np.where(np.arange(n) % 3 == 0, 0, 1 + rng.integers(0, 400, size=n))        # a 66k cluster + small ones
    tiny = rng.integers(0, 5000, size=n)                                               # ~40 cells per cluster
    L = np.stack([big, mid, tiny, rng.permutation(big), rng.integers(0, 50, size=n)]).astype(np.int32)
    w = np.array([2.0, 1.0, 3.0, 1.0, 1.0])
    ecc, sim = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL


@pytest.mark.parametrize("big", [3455, 3456, 3457, 3460, 5000, 6911, 6912, 6913, 6916, 20000])
def test_streaming_chunk_boundaries(big):
    """One cluster around the sizes where an item stops being resident (3456 positions = 864 threads x 4 cells), where
    a streaming item has exactly two chunks (kept in registers) and where it has more (re-streamed in the gather);
    partners with K % 4 == 0 (the pad column opens a new 16-byte group), K = 1 and K = 2."""
    rng = np.random.default_rng(big)
    n = big + 9000
    own = np.concatenate([np.zeros(big, dtype=np.int64), 1 + rng.integers(0, 7, size=n - big)])   # K = 8
    perm = rng.permutation(n)
    L = np.stack([own[perm],
                  rng.integers(0, 4, size=n),          # K = 4
                  np.zeros(n, dtype=np.int64),          # K = 1
                  (rng.random(n) < 0.5).astype(np.int64),  # K = 2
                  np.where(rng.random(n) < 0.1, rng.integers(0, 12, size=n), own[perm]),  # a noisy copy: K = 12
                  own[rng.permutation(n)]]).astype(np.int32)
    w = np.array([1.0, 2.0, 1.0, 1.0, 3.0, 1.0])
    ecc, sim = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL
    assert np.abs(ecs.weighted_element_consistency(list(L)) - oracle.weighted_element_consistency(L)).max() < TOL


def test_many_rows_and_row_stride_remainders():
    """Items with many table rows (small clusters) against partners whose cluster counts cover every remainder of the
    4-entry row alignment, so that the pad column sits at every position of its 16-byte group."""
    rng = np.random.default_rng(77)
    n = 40_000
    L = np.stack([rng.integers(0, k, size=n) for k in (900, 901, 902, 903, 37, 5, 1999, 2000)]).astype(np.int32)
    ecc, sim = oracle.weighted_element_consistency(L, None, calculate_sim_matrix=True)
    res = ecs.weighted_element_consistency(list(L), calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc).max() < TOL and np.abs(res["ecs_matrix"] - sim).max() < TOL


def test_resident_partitions_subsets_without_reupload():
    """Pipeline glue: a resident label set answers comparisons of arbitrary subsets -- contiguous
    runs, scattered and repeated indices, with weights and with the matrix -- exactly like a fresh host-side call."""
    L, w = synth.config("C3", m=23, n=25_000)
    res = ecs.ResidentPartitions(list(L))
    assert len(res) == 23
    rng = np.random.default_rng(8)
    for idx in ([0, 1], list(range(4, 15)), [22, 3, 17, 3, 9], sorted(rng.choice(23, 12, replace=False).tolist()), None):
        sel = L if idx is None else L[idx]
        wi = None if idx is None else w[idx]
        want_ecc, want_sim = oracle.weighted_element_consistency(sel, wi, calculate_sim_matrix=True)
        got = res.weighted_element_consistency(idx, weights=wi, calculate_sim_matrix=True)
        assert np.abs(got["ecc"] - want_ecc).max() < TOL and np.abs(got["ecs_matrix"] - want_sim).max() < TOL
        assert np.abs(res.element_sim_matrix(idx) - want_sim).max() < TOL
        assert np.abs(res.element_consistency(idx) - oracle.weighted_element_consistency(sel)).max() < TOL
    with pytest.raises(ValueError):
        res.element_consistency([5])
    with pytest.raises(_capi.ClustAssessError):
        res.element_sim_matrix([0, 23])
    res.close()


def test_full_size_properties_config5_shape():
    """Size-independent properties at BASELINE.json's cell count (1M cells, K up to 2000): permuting the partitions
    and relabelling their clusters leaves ecc unchanged; duplicating the list leaves the weighted result unchanged
    (R/ECS.R:1470-1491: merging identical partitions is an algebraic identity); 0 <= ecc <= 1."""
    L, _ = synth.config("C5", m=12)
    ecc = ecs.weighted_element_consistency(list(L))
    assert ecc.min() >= 0.0 and ecc.max() <= 1.0 + 1e-12
    rng = np.random.default_rng(5)
    order = rng.permutation(12)
    L2 = np.stack([(L[p].max() + 1 - L[p]) for p in order]).astype(np.int32)  # reversed label order, shuffled partitions
    assert np.abs(ecs.weighted_element_consistency(list(L2)) - ecc).max() < TOL
    # the list taken twice with unit weights == the list once with weight 2 each
    twice = ecs.weighted_element_consistency(list(L) + list(L))
    w2 = ecs.weighted_element_consistency(list(L), weights=np.full(12, 2.0))
    assert np.abs(twice - w2).max() < TOL
    # a sample of cells against the oracle's per-pair routine
    idx = rng.choice(L.shape[1], size=2000, replace=False)
    acc = np.zeros(2000)
    for i in range(12):
        for j in range(i + 1, 12):
            acc += oracle.disjoint_ecs(L[i], L[j])[idx]
    assert np.abs(ecc[idx] - acc / 66.0).max() < TOL


def _matrix_entry(L, w=None, want_sim=True):
    """ca_consistency with the partition-major M x N matrix (the entry bench.py's e2e leg times)."""
    m, n = L.shape
    ecc = np.zeros(n)
    sim = np.zeros((m, m)) if want_sim else None
    lab = np.ascontiguousarray(L, dtype=np.int32)
    _capi.check(_capi.lib().ca_consistency(lab.reshape(-1), n, m, _capi._opt(w), ecc, _capi._opt(sim)))
    return ecc, sim


def test_matrix_and_column_entry_points_agree():
    """ca_consistency / ca_sim_matrix (one M x N matrix) and ca_consistency_cols / ca_sim_matrix_cols (M column
    pointers, what the Rcpp glue passes) are the same job; both against the oracle."""
    L, w = synth.config("C3", m=9, n=20_011)
    ecc_m, sim_m = _matrix_entry(L, w)
    cols = [np.array(L[p]) for p in range(9)]  # separately allocated columns
    res = ecs.weighted_element_consistency(cols, weights=w, calculate_sim_matrix=True)
    exp = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    assert np.abs(ecc_m - exp[0]).max() < TOL and np.abs(sim_m - exp[1]).max() < TOL
    assert np.abs(res["ecc"] - exp[0]).max() < TOL and np.abs(res["ecs_matrix"] - exp[1]).max() < TOL
    sim2 = np.zeros((9, 9))
    _capi.check(_capi.lib().ca_sim_matrix(np.ascontiguousarray(L).reshape(-1), L.shape[1], 9, sim2))
    assert np.abs(sim2 - ecs.element_sim_matrix(cols)).max() < TOL and np.abs(sim2 - exp[1]).max() < TOL
    # agreement and identical-partition merging: matrix form against the column form the Python mirror uses
    lab = np.ascontiguousarray(L[1:]).reshape(-1)
    agr = np.zeros(L.shape[1])
    _capi.check(_capi.lib().ca_agreement(np.ascontiguousarray(L[0]), lab, L.shape[1], 8, agr))
    assert np.abs(agr - ecs.element_agreement(cols[0], cols[1:])).max() < TOL
    want = np.mean([oracle.corrected_l1_mb(L[0], L[p]) for p in range(1, 9)], axis=0)  # R/ECS.R:1639-1647
    assert np.abs(agr - want).max() < TOL
    dup = np.ascontiguousarray(np.stack([L[0], L[1], L[0].max() + 1 - L[0], L[2], L[1]]), dtype=np.int32)  # 0 == 2, 1 == 4
    keep, fout, u = np.zeros(5, np.int32), np.zeros(5), C.c_int32()
    _capi.check(_capi.lib().ca_merge_identical(dup.reshape(-1), dup.shape[1], 5, None, keep, fout, C.byref(u)))
    got = ecs.merge_identical_partitions([{"mb": np.array(r), "freq": 1} for r in dup])
    assert u.value == 3 and keep[:3].tolist() == [0, 1, 3] and fout[:3].tolist() == [2.0, 2.0, 1.0]
    assert [r["freq"] for r in got] == [2, 2, 1] and all(np.array_equal(g["mb"], dup[k]) for g, k in zip(got, (0, 1, 3)))


@pytest.mark.parametrize("n", [700_001, 2_300_003])
def test_staged_upload_of_pageable_buffers(n, monkeypatch):
    """Uploads of 8 MB and more from pageable memory (numpy arrays here, R vectors in the package) are staged through
    pinned slots by several host threads; CA_UPLOAD_DIRECT=1 is the plain cudaMemcpyAsync.  Both must give the same
    job: columns smaller than a 4 MB slot (n = 700 001), columns of several slots with a ragged tail (n = 2 300 003),
    the matrix form (one segment of many slots), the agreement entry and the per-pair entry."""
    L, _ = synth.config("C5", m=5, n=n)
    cols = [np.array(L[p]) for p in range(5)]
    staged_cols = ecs.weighted_element_consistency(cols, calculate_sim_matrix=True)
    staged_mat = _matrix_entry(L)
    staged_agr = ecs.element_agreement(cols[0], cols[1:])
    staged_pair = ecs.element_sim_elscore(cols[0], cols[1])
    monkeypatch.setenv("CA_UPLOAD_DIRECT", "1")
    direct_cols = ecs.weighted_element_consistency(cols, calculate_sim_matrix=True)
    direct_mat = _matrix_entry(L)
    direct_agr = ecs.element_agreement(cols[0], cols[1:])
    direct_pair = ecs.element_sim_elscore(cols[0], cols[1])
    monkeypatch.delenv("CA_UPLOAD_DIRECT")
    # the sim matrix is a sum of exact per-cell ratios in a run-dependent order; ecc likewise: 1e-9, not bit-equality
    for a, b in ((staged_cols["ecc"], direct_cols["ecc"]), (staged_cols["ecs_matrix"], direct_cols["ecs_matrix"]),
                 (staged_mat[0], direct_mat[0]), (staged_mat[1], direct_mat[1]), (staged_cols["ecc"], staged_mat[0]),
                 (staged_agr, direct_agr)):
        assert np.abs(np.asarray(a) - np.asarray(b)).max() < TOL
    assert np.array_equal(staged_pair, direct_pair)
    assert np.array_equal(staged_pair, oracle.disjoint_ecs(L[0], L[1]))
    idx = np.random.default_rng(n).choice(n, size=500, replace=False)
    acc = np.zeros(500)
    for i in range(5):
        for j in range(i + 1, 5):
            acc += oracle.disjoint_ecs(L[i], L[j])[idx]
    assert np.abs(np.asarray(staged_cols["ecc"])[idx] - acc / 10.0).max() < TOL


# ---- merge_partitions against the oracle's golden fixtures ---------------------------------------
def test_merge_partitions_thresholds_ordering_and_ties_match_the_golden_fixtures():
    """ecs_thresh in {0.99, 0.9, 0.5, 1, 0.2}, every order_logic, full and partial frequency ties: partitions kept, their
    order, summed frequencies, avg_agreement, ecc and the (tie-permuted) ECS matrix against tests/golden/merge_golden.json
    (oracle/merge_oracle.py on the reference build's per-pair ECS; R/ECS.R:1057-1135, 1224-1311)."""
    import json
    import os

    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "merge_golden.json")))
    assert len(cases) >= 15
    for c in cases:
        inputs = [np.asarray(x, dtype=np.int32) for x in c["labels"]]
        plist = [{"mb": x, "freq": f} for x, f in zip(inputs, c["freq"])]
        res = ecs.merge_partitions(plist, c["ecs_thresh"], c["order_logic"], return_ecs_matrix=True, check_ties=c["check_ties"])
        ids = [next(k for k, x in enumerate(inputs) if x is p["mb"]) for p in res["partitions"]]
        assert ids == c["ids"], (c["ecs_thresh"], c["order_logic"])
        assert [float(p["freq"]) for p in res["partitions"]] == c["out_freq"]
        assert np.abs(np.asarray(res["ecc"]) - np.asarray(c["ecc"])).max() < TOL
        if c["ecs_matrix"] is not None:
            assert np.abs(res["ecs_matrix"] - np.asarray(c["ecs_matrix"])).max() < TOL
        if c["avg_agreement"] is not None:
            assert np.abs(np.asarray([p["avg_agreement"] for p in res["partitions"]]) - np.asarray(c["avg_agreement"])).max() < TOL


# ---- full sizes inside pytest (the driver's view): C3 and C4 complete, a 4-span job at one million cells ------------------
def _host_threads():
    import os

    return max(1, os.cpu_count() or 1)


def test_config3_full_size_weighted_consistency_and_matrix():
    """BASELINE config 3 at its full size (100 partitions x 100k cells, integer frequency weights): every cell and every
    matrix entry against the oracle (R/ECS.R:1446-1500)."""
    L, w = synth.config("C3")
    assert L.shape == (100, 100_000)
    res = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
    pi, pj = np.triu_indices(L.shape[0], k=1)
    _, acc_unused, means, _ = oracle.pairs_detail(L, pi, pj, nthreads=_host_threads())
    ecc_ref, sim_ref = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc_ref).max() < TOL
    assert np.abs(res["ecs_matrix"] - sim_ref).max() < TOL
    assert np.abs(res["ecs_matrix"][pi, pj] - means).max() < TOL  # the threaded per-pair body agrees with the sequential loop


def test_config4_full_size_matrix():
    """BASELINE config 4 at its full size (element_sim_matrix of 1000 partitions x 50k cells): 499 500 entries on the GPU,
    1500 randomly drawn ones against the oracle's mean(disjointECS) (R/ECS.R:942-965), symmetry and unit diagonal for all."""
    L, _ = synth.config("C4")
    assert L.shape == (1000, 50_000)
    sim = ecs.element_sim_matrix(list(L))
    assert sim.shape == (1000, 1000) and np.array_equal(sim, sim.T) and np.all(np.diag(sim) == 1.0)
    rng = np.random.default_rng(44)
    pi = rng.integers(0, 999, size=1500)
    pj = np.array([rng.integers(i + 1, 1000) for i in pi])
    _, _, means, _ = oracle.pairs_detail(L, pi, pj, nthreads=_host_threads())
    assert np.abs(sim[pi, pj] - means).max() < TOL
    assert np.all((sim >= 0.0) & (sim <= 1.0))


def test_one_million_cells_four_spans_every_cell_against_the_oracle():
    """The headline family at N = 1M with 100 partitions = four 32-partition spans, resident and streaming items, up to 2000
    clusters: ALL 4950 pairs on the host (threaded per-pair body of R/ECS.R:1478-1484, ~0.1 core-seconds per pair), every
    cell's consistency and every matrix entry compared."""
    L, _ = synth.config("C5", m=100)
    m, n = L.shape
    assert n == 1_000_000
    pi, pj = np.triu_indices(m, k=1)
    _, acc, means, _ = oracle.pairs_detail(L, pi, pj, nthreads=_host_threads())
    res = ecs.weighted_element_consistency(list(L), calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - acc / float(pi.size)).max() < TOL
    assert np.abs(res["ecs_matrix"][pi, pj] - means).max() < TOL
    again = ecs.weighted_element_consistency(list(L), calculate_sim_matrix=True)
    assert np.array_equal(again["ecc"], res["ecc"]) and np.array_equal(again["ecs_matrix"], res["ecs_matrix"])  # same bits in every run


# ---- the pipelined host path: upload of later partition groups overlaps the computation of earlier ones -----------------------
def _louvainish(m, n, k_lo, k_hi, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, k_lo, size=n)
    out = np.empty((m, n), dtype=np.int32)
    for p in range(m):
        k = int(rng.integers(k_lo, k_hi + 1))
        x = base.copy()
        hit = rng.random(n) < rng.uniform(0.02, 0.2)
        x[hit] = rng.integers(0, k, size=int(hit.sum()))
        out[p] = rng.permutation(k_hi + 1)[x] + 1
    return out


def test_pipelined_host_path_equals_plain_path_and_oracle(monkeypatch):
    """256 partitions x 140k cells from host buffers (143 MB, 8 spans): the host entry cuts the partitions into groups from the
    top and computes each group's pairs while the next one is still being uploaded.  Same bits as the plain path
    (CA_NO_PIPELINE=1), every cell and every matrix entry against the oracle."""
    monkeypatch.setenv("CA_PIPE_MIN_MB", "128")  # the library pipelines from 512 MB of labels on; this set has 143 MB
    L = _louvainish(256, 140_000, 20, 60, seed=77)
    m, n = L.shape
    w = (1 + np.random.default_rng(78).poisson(1.0, size=m)).astype(np.float64)
    cols = [np.ascontiguousarray(r) for r in L]
    res = ecs.weighted_element_consistency(cols, weights=w, calculate_sim_matrix=True)
    only_ecc = ecs.weighted_element_consistency(cols, weights=w)
    sim_only = ecs.element_sim_matrix(cols)
    packed = np.empty(n, dtype=np.float64)
    _capi.check(_capi.lib().ca_consistency(L.reshape(-1), n, m, w.ctypes.data, packed, None))  # the matrix form of the same entry
    monkeypatch.setenv("CA_NO_PIPELINE", "1")
    plain = ecs.weighted_element_consistency(cols, weights=w, calculate_sim_matrix=True)
    monkeypatch.delenv("CA_NO_PIPELINE")
    assert np.array_equal(res["ecc"], plain["ecc"]) and np.array_equal(res["ecs_matrix"], plain["ecs_matrix"])
    assert np.array_equal(packed, only_ecc)  # matrix form == column form of the same entry
    assert np.abs(only_ecc - plain["ecc"]).max() < 1e-13  # another kernel instantiation (no matrix): same numbers, not necessarily the same bits
    unit = ecs.element_sim_matrix(cols)
    assert np.array_equal(sim_only, unit) and np.abs(sim_only - plain["ecs_matrix"]).max() < 1e-12  # possibly another route to the matrix
    ecc_ref, sim_ref = oracle.weighted_element_consistency(L[:64], w[:64], calculate_sim_matrix=True)
    sub = ecs.weighted_element_consistency(cols[:64], weights=w[:64], calculate_sim_matrix=True)  # 64 partitions: the plain path
    assert np.abs(sub["ecc"] - ecc_ref).max() < TOL and np.abs(sub["ecs_matrix"] - sim_ref).max() < TOL
    pi, pj = np.triu_indices(m, k=1)
    _, _, means, _ = oracle.pairs_detail(L, pi, pj, nthreads=_host_threads())
    assert np.abs(res["ecs_matrix"][pi, pj] - means).max() < TOL
    # weighted ecc of the whole set from the per-pair ECS sums needs all pairs' vectors; check it through the identity
    # ecc = sum_i w_i (w_i - 1) / 2 / norm + sum_{i<j} w_i w_j ECS_ij / norm on a sample of cells instead
    cells = np.random.default_rng(79).choice(n, size=300, replace=False)
    tot = w.sum()
    norm = tot * (tot - 1) / 2
    want = np.full(cells.size, float((w * (w - 1) / 2).sum() / norm))
    _, _, _, cv = oracle.pairs_detail(L, pi, pj, cells=cells, nthreads=_host_threads())
    want += (cv * (w[pi] * w[pj] / norm)[:, None]).sum(axis=0)
    assert np.abs(res["ecc"][cells] - want).max() < TOL


def test_pipelined_host_path_bails_out_when_later_groups_have_many_more_clusters(monkeypatch):
    """The row stride of the partner arrays is fixed from the first uploaded group (the LAST partitions) with a 2x margin; when an
    earlier partition turns out to have far more clusters the pipelined path gives up and the plain path answers -- same result."""
    monkeypatch.setenv("CA_PIPE_MIN_MB", "128")
    rng = np.random.default_rng(81)
    m, n = 256, 131_072
    L = np.stack([rng.integers(0, 40, size=n) for _ in range(m)]).astype(np.int32)
    L[5] = rng.integers(0, 3000, size=n)
    L[5][:3000] = np.arange(3000)
    cols = [np.ascontiguousarray(r) for r in L]
    got = ecs.element_sim_matrix(cols)
    pi = np.array([5] * 6 + [0, 100, 254])
    pj = np.array([6, 40, 100, 200, 254, 255, 5, 255, 255])
    _, _, means, _ = oracle.pairs_detail(L, pi, pj, nthreads=4)
    assert np.abs(got[pi, pj] - means).max() < TOL and np.array_equal(got, got.T)


def ecs_plain_ecc(cols, w, monkeypatch):
    monkeypatch.setenv("CA_RT", "0")
    out = ecs.weighted_element_consistency(cols, weights=w)
    monkeypatch.delenv("CA_RT")
    return out


@pytest.mark.parametrize("shape", ["few clusters", "many clusters", "giant clusters"])
def test_matrix_from_table_entries_equals_matrix_from_cells(monkeypatch, shape):
    """The ECS matrix has two routes through the wave kernel: summed over the non-empty TABLE entries, T^2 / max(S_a, S_b) -- the
    reference's own formula for a pair's average (src/calculate_ecs.cpp:58-67), taken when the partitions have few clusters --
    or per (cell, partner) in the gather.  Both against the oracle, against each other, with and without ecc, weighted and not;
    the table route twice for identical bits.  'giant clusters': a partner cluster of >= 65535 cells (size looked up in global
    memory) and an owner cluster of > 65535 cells (streaming item with 32-bit counters)."""
    rng = np.random.default_rng(91)
    if shape == "few clusters":
        L = _louvainish(24, 6000, 5, 15, seed=92)
    elif shape == "many clusters":
        L = _louvainish(12, 20_000, 250, 400, seed=93)
    else:
        n = 240_000  # few clusters of more than 8192 cells on average: the library picks the ratio tables by itself here
        L = _louvainish(10, n, 8, 20, seed=94)
        L[0] = np.where(rng.random(n) < 0.75, 1, L[0] + 1)      # one cluster of ~150k cells in the first owner
        L[7] = np.where(rng.random(n) < 0.40, 3, L[7] + 5)      # ~80k cells in a partner
    m, n = L.shape
    w = (1 + rng.poisson(1.0, size=m)).astype(np.float64)
    cols = [np.ascontiguousarray(r) for r in L]
    pi, pj = np.triu_indices(m, k=1)
    _, _, means, _ = oracle.pairs_detail(L, pi, pj, nthreads=_host_threads())
    ecc_ref, sim_ref = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    got = {}
    for scan in ("0", "1"):
        monkeypatch.setenv("CA_SIM_SCAN", scan)
        got[scan] = (ecs.element_sim_matrix(cols), ecs.weighted_element_consistency(cols, weights=w, calculate_sim_matrix=True),
                     ecs.weighted_element_consistency(cols, calculate_sim_matrix=True))
        sim, wres, ures = got[scan]
        assert np.abs(sim[pi, pj] - means).max() < TOL and np.array_equal(sim, sim.T) and np.all(np.diag(sim) == 1.0)
        assert np.abs(wres["ecs_matrix"] - sim_ref).max() < TOL and np.abs(wres["ecc"] - ecc_ref).max() < TOL
        assert np.abs(ures["ecs_matrix"] - sim).max() < 1e-12
    assert np.abs(got["0"][0] - got["1"][0]).max() < 1e-12
    assert np.abs(got["0"][1]["ecc"] - got["1"][1]["ecc"]).max() < 1e-13
    again = ecs.element_sim_matrix(cols)  # CA_SIM_SCAN is still 1
    assert np.array_equal(again, got["1"][0])
    # ratio tables (one IEEE division per table entry, the gather only adds; chosen for few, huge clusters): same numbers again
    monkeypatch.setenv("CA_RT", "1")
    rt = ecs.weighted_element_consistency(cols, weights=w, calculate_sim_matrix=True)
    rt_ecc_only = ecs.weighted_element_consistency(cols, weights=w)
    assert np.abs(rt["ecc"] - ecc_ref).max() < TOL and np.abs(rt["ecs_matrix"] - sim_ref).max() < TOL
    assert np.abs(rt["ecc"] - got["1"][1]["ecc"]).max() < 1e-12 and np.abs(rt["ecs_matrix"] - got["1"][1]["ecs_matrix"]).max() < 1e-11
    assert np.abs(rt_ecc_only - rt["ecc"]).max() < 1e-13
    assert np.array_equal(ecs.weighted_element_consistency(cols, weights=w, calculate_sim_matrix=True)["ecc"], rt["ecc"])  # same bits twice
    monkeypatch.delenv("CA_RT")
    monkeypatch.delenv("CA_SIM_SCAN")
    chosen = ecs.element_sim_matrix(cols)  # the library's own choice is one of the two
    assert np.array_equal(chosen, got["0"][0]) or np.array_equal(chosen, got["1"][0])
    own = ecs.weighted_element_consistency(cols, weights=w)
    assert np.array_equal(own, rt_ecc_only) or np.array_equal(own, ecs_plain_ecc(cols, w, monkeypatch))
    if shape == "giant clusters":
        assert np.array_equal(own, rt_ecc_only)
