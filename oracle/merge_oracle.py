"""CPU restatement of merge_partitions and merge_partitions_ecs (R/ECS.R:1057-1135, 1166-1311).  TEST INFRASTRUCTURE ONLY.

The threshold merge and the ordering / tie logic are O(M^2) scalar R code on top of the per-pair ECS; this file follows that R
code operation by operation in R's own terms -- 1-based indices, `which.max` over the column-major vector, `min(..., na.rm = T)`,
`order(..., decreasing = TRUE)` (radix: stable) -- on top of the per-pair routines of the oracle (liboracle.so, or the
unmodified reference build when `use_ref` is set).  It is written independently of the product's host logic
(clustassess_b200/ecs.py) so that the two pin each other; the fixtures of tests/golden/merge_golden.json come from here
(oracle/gen_golden.py).

A partition record is a dict {"mb": int array, "freq": number, "id": index into the input list}.
"""
from __future__ import annotations

import math

import numpy as np

from . import disjoint_ecs, have_ref, merge_identical_partitions as _merge_identical, ref_disjoint_ecs


def _pair(a, b, use_ref):
    return (ref_disjoint_ecs if (use_ref and have_ref()) else disjoint_ecs)(a, b)


def _r_mean(x):
    """R's mean(): long-double sum / n, then one correction pass (src/main/summary.c)."""
    x = np.asarray(x, dtype=np.longdouble)
    s = x.sum() / x.size
    return float(s + (x - s).sum() / x.size)


def sim_matrix(mbs, use_ref=False):
    """element_sim_matrix_flat_disjoint (R/ECS.R:931-965): mean ECS per pair, mirrored, unit diagonal."""
    m = len(mbs)
    sim = np.ones((m, m))
    for i in range(m - 1):
        for j in range(i + 1, m):
            sim[i, j] = sim[j, i] = _r_mean(_pair(mbs[i], mbs[j], use_ref))
    return sim


def weighted_consistency(mbs, weights, use_ref=False):
    """weighted_element_consistency(calculate_sim_matrix = TRUE) (R/ECS.R:1466-1499), in the R loop's order."""
    m, n = len(mbs), len(mbs[0])
    w = np.asarray(weights, dtype=np.float64)
    total = w.sum()
    norm = total * (total - 1) / 2
    cons = np.zeros(n)
    sim = np.zeros((m, m))
    for i in range(m - 1):
        cons = cons + w[i] / norm * (w[i] - 1) / 2
        for j in range(i + 1, m):
            e = _pair(mbs[i], mbs[j], use_ref)
            sim[i, j] = sim[j, i] = _r_mean(e)
            cons = cons + e * w[i] / norm * w[j]
        sim[i, i] = 1
    cons = cons + w[m - 1] / norm * (w[m - 1] - 1) / 2
    sim[m - 1, m - 1] = 1
    return cons, sim


def merge_partitions_ecs(partition_list, ecs_thresh=0.99, use_ref=False):
    """R/ECS.R:1057-1135."""
    nparts = len(partition_list)
    if nparts == 1:
        return partition_list
    sim = sim_matrix([p["mb"] for p in partition_list], use_ref)  # R matrix sim[i, j], used 1-based below
    S = lambda i, j: sim[i - 1, j - 1]
    groups = {}  # names(partition_groups) in insertion order
    for i in range(1, nparts + 1):
        sim[i - 1, i - 1] = math.nan
        groups[i] = [i]
    while len(groups) > 1:
        if np.nanmax(sim) < ecs_thresh:  # max(sim_matrix, na.rm = TRUE) < ecs_thresh
            break
        # which.max(sim_matrix)[1]: first maximum of the column-major vector, NAs skipped, 1-based
        colmajor = sim.flatten(order="F")
        index = int(np.nanargmax(colmajor)) + 1
        first = index % nparts
        second = index // nparts + 1
        if first == 0:
            first = nparts
            second = second - 1
        groups[first] = groups[first] + groups[second]
        del groups[second]
        for i in list(groups.keys()):
            if first < i:
                if not math.isnan(S(first, i)):
                    vals = [S(first, i), S(second, i), S(i, second)]
                    sim[first - 1, i - 1] = min(v for v in vals if not math.isnan(v))
            else:
                if not math.isnan(S(i, first)):
                    vals = [S(i, first), S(second, i), S(i, second)]
                    sim[i - 1, first - 1] = min(v for v in vals if not math.isnan(v))
        sim[second - 1, :] = math.nan
        sim[:, second - 1] = math.nan
    merged = []
    for kept, members in groups.items():
        rec = dict(partition_list[kept - 1])
        rec["freq"] = sum(partition_list[x - 1]["freq"] for x in members)
        merged.append(rec)
    return merged


def _order_decreasing(values):
    """order(x, decreasing = TRUE): radix sort, stable for ties."""
    return sorted(range(len(values)), key=lambda t: -values[t])


def merge_partitions(partition_list, ecs_thresh=1, order_logic="freq", return_ecs_matrix=False, check_ties=True, use_ref=False):
    """R/ECS.R:1166-1311 for a list of membership vectors or of {mb, freq} records."""
    if not isinstance(partition_list[0], dict):
        partition_list = [{"mb": np.asarray(x), "freq": 1} for x in partition_list]
    partition_list = [dict(p, id=k) for k, p in enumerate(partition_list)]
    if ecs_thresh == 1:
        labels = np.stack([np.asarray(p["mb"], dtype=np.int32) for p in partition_list])
        keep, freq = _merge_identical(labels, [float(p["freq"]) for p in partition_list])
        part_list = [dict(partition_list[int(k)], freq=float(f)) for k, f in zip(keep, freq)]
    else:
        part_list = merge_partitions_ecs(partition_list, ecs_thresh, use_ref)
    n = len(part_list[0]["mb"])
    if len(part_list) == 1:
        return {"partitions": part_list, "ecc": np.ones(n)}
    ecc, mat = weighted_consistency([p["mb"] for p in part_list], [p["freq"] for p in part_list], use_ref)
    if order_logic not in ("freq", "avg_agreement"):
        out = {"partitions": part_list, "ecc": ecc}
        if return_ecs_matrix:
            out["ecs_matrix"] = mat
        return out
    n_unique = mat.shape[0]
    n_total = sum(p["freq"] for p in part_list)
    for i in range(1, n_unique + 1):
        total = (partition_list[i - 1]["freq"] - 1) / (n_total - 1)  # the PRE-merge list, as the reference reads it (R/ECS.R:1240)
        for j in range(1, n_unique + 1):
            if i == j:
                continue
            total = total + mat[i - 1, j - 1] * partition_list[j - 1]["freq"] / (n_total - 1)
        part_list[i - 1]["avg_agreement"] = total
    ordered = _order_decreasing([p[order_logic] for p in part_list])
    part_list = [part_list[t] for t in ordered]
    if return_ecs_matrix:
        mat = mat[np.ix_(ordered, ordered)]
    check_ties = check_ties and len(part_list) > 1
    if check_ties:
        second_key = "avg_agreement" if order_logic == "freq" else "freq"
        ref_value = part_list[0][order_logic]
        i = len(part_list)  # R leaves the loop variable at its last value when it never breaks
        for t in range(2, len(part_list) + 1):
            if part_list[t - 1][order_logic] != ref_value:
                i = t - 1
                break
        if i > 1:
            ties = _order_decreasing([p[second_key] for p in part_list[:i]])
            part_list[:i] = [part_list[t] for t in ties]
            if return_ecs_matrix:
                reordered = mat[np.ix_(ties, ties)]
                mat = mat.copy()
                mat[:i, :i] = reordered
    out = {"partitions": part_list, "ecc": ecc}
    if return_ecs_matrix:
        out["ecs_matrix"] = mat
    return out
