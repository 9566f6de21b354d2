"""CPU tests that PIN the oracle (oracle/ecs_oracle.c) before anything is checked against it:
against the committed golden vectors (generated from the unmodified reference C++ by
oracle/gen_golden.py), against the reference itself when oracle/_ref is present, against the one
documented known answer, and through the properties the reference's own tests assert
(tests/testthat/test-ECS.R)."""
import numpy as np
import pytest

import oracle
from clustassess_b200 import synth


def test_pairs_match_golden(golden):
    for k in range(int(golden["n_pairs"])):
        a, b = golden[f"p{k}_a"], golden[f"p{k}_b"]
        assert np.array_equal(oracle.cont_table(a, b), golden[f"p{k}_table"])
        assert np.array_equal(oracle.disjoint_ecs(a, b), golden[f"p{k}_ecs"])  # one IEEE division: bit-exact
        avg = float(golden[f"p{k}_avg"])
        if np.isnan(avg):
            # reference quirk (src/calculate_ecs.cpp:65): a label unused in BOTH ranges gives an empty
            # row and an empty column whose cell adds 0*0/(n*0) = NaN.  The oracle (and the product)
            # skip such cells and return mean(disjointECS), which is what R/ECS.R:59-100 documents.
            t = golden[f"p{k}_table"]
            assert (t.sum(1) == 0).any() and (t.sum(0) == 0).any()
            avg = float(golden[f"p{k}_ecs"].mean())
        assert abs(oracle.disjoint_ecs_average(a, b) - avg) < 1e-12
        # second, independent formula (R/ECS.R:468-507; test-ECS.R:18-32 type invariance)
        assert np.abs(oracle.corrected_l1_mb(a, b) - golden[f"p{k}_ecs"]).max() < 1e-12


def test_set_level_matches_golden(golden):
    L, w = golden["set_labels"], golden["set_weights"]
    ecc, sim = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    assert np.abs(ecc - golden["set_ecc_w"]).max() < 1e-12
    assert np.abs(sim - golden["set_sim"]).max() < 1e-12
    assert np.abs(oracle.element_sim_matrix(L) - golden["set_sim"]).max() < 1e-12
    assert np.abs(oracle.weighted_element_consistency(L) - golden["set_ecc_unw"]).max() < 1e-12
    # dedup + weights == plain unweighted ECC over all M (algebraic invariance)
    assert np.abs(oracle.element_consistency(L) - golden["set_ecc_unw"]).max() < 1e-12


def test_documented_known_answer(golden):
    # docs/reference/merge_partitions.html:117-142
    L = golden["doc_labels"]
    keep, freq = oracle.merge_identical_partitions(L)
    assert keep.tolist() == [0, 1] and freq.tolist() == golden["doc_freq"].tolist()
    assert np.abs(oracle.element_consistency(L) - golden["doc_ecc"]).max() < 5e-8  # printed to 7 digits
    assert np.allclose(oracle.element_consistency(L), [7 / 9, 7 / 9, 5 / 9], atol=1e-15)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
def test_against_unmodified_reference():
    rng = np.random.default_rng(7)
    for _ in range(50):
        n = int(rng.integers(1, 400))
        a = rng.integers(-5, rng.integers(-4, 30), size=n).astype(np.int32)
        b = rng.integers(3, rng.integers(4, 60), size=n).astype(np.int32)
        assert np.array_equal(oracle.cont_table(a, b), oracle.ref_cont_table(a, b))
        assert np.array_equal(oracle.disjoint_ecs(a, b), oracle.ref_disjoint_ecs(a, b))
        r = oracle.ref_disjoint_ecs_average(a, b)
        if not np.isnan(r):  # NaN: the empty-row-and-column quirk, see test_pairs_match_golden
            assert abs(oracle.disjoint_ecs_average(a, b) - r) < 1e-12
    L, _ = synth.config("C1")
    assert np.array_equal(oracle.disjoint_ecs(L[0], L[1]), oracle.ref_disjoint_ecs(L[0], L[1]))


def test_reference_properties():
    # tests/testthat/test-ECS.R:1-12, 34-44, 98-140 restated on the oracle
    rng = np.random.default_rng(1234)
    L = rng.integers(1, 11, size=(20, 100)).astype(np.int32)
    for i in range(0, 20, 2):
        e = oracle.disjoint_ecs(L[i], L[i + 1])
        assert 0 <= e.min() and e.max() <= 1
        assert abs(e.mean() - oracle.disjoint_ecs_average(L[i], L[i + 1])) < 1e-12
    ecc = oracle.element_consistency(L)
    acc = np.zeros(100)
    for i in range(20):
        for j in range(20):
            if i != j:
                acc += oracle.disjoint_ecs(L[i], L[j])
    assert np.abs(ecc - acc / (20 * 19)).max() < 1e-12
    sim = oracle.element_sim_matrix(L)
    assert np.array_equal(sim, sim.T) and np.all(np.diag(sim) == 1)
    assert abs(sim[3, 11] - oracle.disjoint_ecs(L[3], L[11]).mean()) < 1e-12


def test_dedup_quirks():
    # label gap inside the range -> all-zero row -> never identical, even to itself (R/ECS.R:1003-1015)
    assert not oracle.are_identical_memberships([1, 3, 3, 3], [1, 3, 3, 3])
    assert oracle.are_identical_memberships([1, 3, 3], [1, 3, 3])  # nrow == n shortcut fires first (R/ECS.R:999)
    assert oracle.are_identical_memberships([1, 2, 2], [5, 4, 4])
    assert not oracle.are_identical_memberships([1, 2, 2], [1, 1, 2])
    assert oracle.are_identical_memberships([1, 2, 3], [3, 1, 2])  # all singletons (R/ECS.R:999)
    with pytest.raises(ValueError):
        oracle.element_consistency(np.ones((1, 5), dtype=np.int32))


def test_generators_are_seeded_and_distinct():
    L1, _ = synth.config("C1")
    L2, _ = synth.config("C1")
    assert np.array_equal(L1, L2) and L1.dtype == np.int32 and L1.min() == 1
    keep, _ = oracle.merge_identical_partitions(L1)
    assert keep.size == L1.shape[0]  # U == M: dedup removes nothing
    sub, _ = synth.config("C1", parts=[3, 17])
    assert np.array_equal(sub, L1[[3, 17]])
    L3, w3 = synth.config("C3", m=6, n=5000)
    assert w3.shape == (6,) and np.all(w3 >= 1) and np.all(w3 == np.round(w3))
    a, b = synth.config("C2")[0]
    assert len(np.unique(a)) == 20 and len(np.unique(b)) == 35


# ---- merge_partitions below the identity threshold, ordering and ties ---------------------------
def _merge_golden():
    import json
    import os

    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "merge_golden.json")))


def test_merge_oracle_reproduces_the_golden_fixtures_with_the_c_port():
    """tests/golden/merge_golden.json was generated with the reference build's per-pair ECS (oracle/gen_golden.py); the same
    restatement of R/ECS.R:1057-1135, 1166-1311 on top of the plain-C port must give the same partitions, order and numbers."""
    from oracle import merge_oracle

    for c in _merge_golden():
        plist = [{"mb": np.asarray(x, dtype=np.int32), "freq": f} for x, f in zip(c["labels"], c["freq"])]
        res = merge_oracle.merge_partitions(plist, c["ecs_thresh"], c["order_logic"], return_ecs_matrix=True, check_ties=c["check_ties"])
        assert [p["id"] for p in res["partitions"]] == c["ids"]
        assert [float(p["freq"]) for p in res["partitions"]] == c["out_freq"]
        assert np.abs(np.asarray(res["ecc"]) - np.asarray(c["ecc"])).max() < 1e-12
        if c["ecs_matrix"] is not None:
            assert np.abs(res["ecs_matrix"] - np.asarray(c["ecs_matrix"])).max() < 1e-12
        if c["avg_agreement"] is not None:
            assert np.abs(np.asarray([p["avg_agreement"] for p in res["partitions"]]) - np.asarray(c["avg_agreement"])).max() < 1e-12


def test_merge_oracle_documented_known_answer_and_invariants():
    """docs/reference/merge_partitions.html:117-142: partitions (1,1,2) freq 2 and (2,2,2) freq 1, ecc 7/9 7/9 5/9,
    avg_agreement 0.2777778 for the second; merging conserves the total frequency at every threshold."""
    from oracle import merge_oracle

    res = merge_oracle.merge_partitions([np.array([1, 1, 2]), np.array([2, 2, 2]), np.array([2, 2, 1])], 1, "freq")
    assert [p["freq"] for p in res["partitions"]] == [2.0, 1.0]
    assert np.allclose(res["ecc"], [7 / 9, 7 / 9, 5 / 9], atol=1e-12)
    assert abs(res["partitions"][1]["avg_agreement"] - 0.2777778) < 1e-7
    for c in _merge_golden():
        assert sum(c["out_freq"]) == sum(c["freq"])
        kept = [len(x["ids"]) for x in _merge_golden() if x["labels"] == c["labels"] and x["freq"] == c["freq"] and x["order_logic"] == c["order_logic"]]
        assert kept == sorted(kept, reverse=True) or len(kept) == 1  # a lower threshold never keeps more partitions
