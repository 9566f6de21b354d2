"""Generate tests/golden/ecs_golden.npz from the UNMODIFIED reference C++ (oracle/_ref/libref_ecs.so).

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):

    python oracle/gen_golden.py

Per-pair vectors (table, per-element ECS, average) come straight from the reference's
myContTable / disjointECS / disjointECSaverage (src/calculate_ecs.cpp:7-70).  Set-level vectors
(weighted ECC, ECS matrix) are composed HERE in numpy from the reference's per-pair outputs, following
the R loop order of R/ECS.R:1466-1499 and R/ECS.R:942-965 -- an independent statement from the C
restatement in ecs_oracle.c, so the two pin each other.  The one deterministic known answer the
reference documents (docs/reference/merge_partitions.html:117-142) is stored verbatim.
tests/golden/merge_golden.json holds merge_partitions results for ecs_thresh in {0.99, 0.9, 0.5, 1, 0.2} with every
order_logic, full and partial frequency ties: oracle/merge_oracle.py (R/ECS.R:1057-1135, 1166-1311 restated in R's own
terms) on top of the reference build's per-pair ECS.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402


def pair_cases(rng):
    cases = []
    for _ in range(6):  # the shape the reference's own tests use (tests/testthat/test-ECS.R:4-5)
        cases.append((rng.integers(1, 11, size=100), rng.integers(1, 11, size=100)))
    cases.append((np.array([1, 1, 2]), np.array([2, 2, 2])))                    # R/ECS.R:1164 example
    cases.append((np.array([1, 3, 3, 7, 7, 7]), np.array([5, 5, 9, 9, 9, 5])))  # gaps inside the range
    cases.append((np.array([-3, -3, 0, 2, 2]), np.array([10, 11, 12, 13, 14])))  # negatives, singletons
    cases.append((np.ones(17, dtype=int), np.ones(17, dtype=int) * 4))          # one cluster each
    cases.append((np.arange(1, 41), np.arange(40, 0, -1)))                       # all singletons
    cases.append((rng.integers(1, 4, size=1000), rng.integers(1, 300, size=1000)))
    cases.append((rng.integers(1, 700, size=5000), rng.integers(1, 900, size=5000)))
    return [(np.asarray(a, dtype=np.int32), np.asarray(b, dtype=np.int32)) for a, b in cases]


def wec_from_ref(labels, weights):
    """R/ECS.R:1466-1499 with the reference's own disjointECS per pair."""
    m, n = labels.shape
    w = np.ones(m) if weights is None else np.asarray(weights, dtype=np.float64)
    total = w.sum()
    norm = total * (total - 1) / 2
    cons = np.zeros(n)
    sim = np.zeros((m, m))
    for i in range(m - 1):
        cons = cons + w[i] / norm * (w[i] - 1) / 2
        for j in range(i + 1, m):
            e = oracle.ref_disjoint_ecs(labels[i], labels[j])
            sim[i, j] = sim[j, i] = e.mean()
            cons = cons + e * w[i] / norm * w[j]
        sim[i, i] = 1
    cons = cons + w[m - 1] / norm * (w[m - 1] - 1) / 2
    sim[m - 1, m - 1] = 1
    return cons, sim


def merge_cases(rng):
    """Inputs for merge_partitions with ecs_thresh < 1 and for the ordering / tie logic (R/ECS.R:1057-1135, 1224-1311):
    a base partition and noisy relatives at graded distances, so that the pairwise ECS straddles the thresholds; frequencies
    with full and partial ties."""
    n = 240
    base = rng.integers(1, 7, size=n)

    def noisy(frac, k=6):
        x = base.copy()
        hit = rng.random(n) < frac
        x[hit] = rng.integers(1, k + 1, size=int(hit.sum()))
        return x

    fam = [base, noisy(0.004), noisy(0.02), noisy(0.05), noisy(0.12), noisy(0.3), noisy(0.6), rng.integers(1, 5, size=n),
           (base % 6) * 2 + 3, noisy(0.02, 9)]
    cases = []
    for thresh in (0.99, 0.9, 0.5):
        for order_logic in ("freq", "avg_agreement", "none"):
            cases.append(dict(labels=fam, freq=[1] * len(fam), ecs_thresh=thresh, order_logic=order_logic, check_ties=True))
    cases.append(dict(labels=fam, freq=[3, 1, 3, 2, 1, 3, 2, 1, 1, 2], ecs_thresh=0.9, order_logic="freq", check_ties=True))
    cases.append(dict(labels=fam, freq=[3, 1, 3, 2, 1, 3, 2, 1, 1, 2], ecs_thresh=0.9, order_logic="freq", check_ties=False))
    cases.append(dict(labels=fam[:6], freq=[2, 2, 5, 5, 5, 1], ecs_thresh=1, order_logic="freq", check_ties=True))
    cases.append(dict(labels=fam[:6], freq=[2, 2, 5, 5, 5, 1], ecs_thresh=1, order_logic="avg_agreement", check_ties=True))
    cases.append(dict(labels=[fam[0], fam[8], fam[3]], freq=[1, 4, 2], ecs_thresh=1, order_logic="freq", check_ties=True))  # identical pair merges
    cases.append(dict(labels=fam[:4], freq=[1, 1, 1, 1], ecs_thresh=0.2, order_logic="freq", check_ties=True))  # everything merges
    return cases


def write_merge_golden(rng, dst):
    import json

    from oracle import merge_oracle

    out = []
    for c in merge_cases(rng):
        plist = [{"mb": np.asarray(x, dtype=np.int32), "freq": f} for x, f in zip(c["labels"], c["freq"])]
        res = merge_oracle.merge_partitions(plist, c["ecs_thresh"], c["order_logic"], return_ecs_matrix=True, check_ties=c["check_ties"],
                                            use_ref=True)  # per-pair ECS from the unmodified reference build
        rec = dict(labels=[np.asarray(x).tolist() for x in c["labels"]], freq=list(c["freq"]), ecs_thresh=c["ecs_thresh"],
                   order_logic=c["order_logic"], check_ties=c["check_ties"],
                   ids=[int(p["id"]) for p in res["partitions"]], out_freq=[float(p["freq"]) for p in res["partitions"]],
                   avg_agreement=[float(p["avg_agreement"]) for p in res["partitions"]] if "avg_agreement" in res["partitions"][0] else None,
                   ecc=np.asarray(res["ecc"]).tolist(), ecs_matrix=np.asarray(res["ecs_matrix"]).tolist() if "ecs_matrix" in res else None)
        out.append(rec)
    json.dump(out, open(dst, "w"))
    print("wrote", dst, os.path.getsize(dst), "bytes,", len(out), "cases; partitions kept per case:", [len(r["ids"]) for r in out])


def main():
    if not oracle.have_ref():
        raise SystemExit("oracle/_ref/libref_ecs.so missing: run `make -C oracle` where /root/reference exists")
    rng = np.random.Generator(np.random.PCG64(20240607))
    out = {}
    cases = pair_cases(rng)
    out["n_pairs"] = np.int64(len(cases))
    for k, (a, b) in enumerate(cases):
        out[f"p{k}_a"], out[f"p{k}_b"] = a, b
        out[f"p{k}_table"] = oracle.ref_cont_table(a, b)
        out[f"p{k}_ecs"] = oracle.ref_disjoint_ecs(a, b)
        out[f"p{k}_avg"] = np.float64(oracle.ref_disjoint_ecs_average(a, b))
    # set-level: 12 partitions x 300 cells, K in 2..9, two identical copies and integer weights
    m, n = 12, 300
    labels = np.stack([rng.integers(1, rng.integers(3, 10), size=n) for _ in range(m)]).astype(np.int32)
    labels[7] = (labels[2] % 9) * 3 + 5       # a relabelled copy of partition 2 (same partition)
    labels[10] = labels[4]                    # an exact copy
    weights = (1 + rng.poisson(1.0, size=m)).astype(np.float64)
    out["set_labels"], out["set_weights"] = labels, weights
    out["set_ecc_w"], out["set_sim"] = wec_from_ref(labels, weights)
    out["set_ecc_unw"], _ = wec_from_ref(labels, None)
    # documented known answer: merge_partitions(list(c(1,1,2), c(2,2,2), c("B","B","A")), 1)
    out["doc_labels"] = np.array([[1, 1, 2], [2, 2, 2], [2, 2, 1]], dtype=np.int32)  # "B","B","A" as factor codes
    out["doc_ecc"] = np.array([0.7777778, 0.7777778, 0.5555556])
    out["doc_freq"] = np.array([2.0, 1.0])
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ecs_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")
    # merge_partitions below the identity threshold, ordering and ties: its own random stream, so the file above does not change
    write_merge_golden(np.random.Generator(np.random.PCG64(20261020)), os.path.join(os.path.dirname(dst), "merge_golden.json"))


if __name__ == "__main__":
    main()
