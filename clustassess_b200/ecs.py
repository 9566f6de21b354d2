"""Host-side mirror of ClustAssess's flat-disjoint ECS API (reference: R/ECS.R).

Same function names, argument meaning and error behaviour as the R package, with the native work done
by the B200 library through the C ABI (include/clustassess_b200.h) instead of the Rcpp routines of
src/calculate_ecs.cpp.  The R package itself would keep its R/ECS.R wrappers and call the same C ABI
through the shim in r/clustassess_shim.c; this module exists because the build
container has no R, and it is what the parity tests and bench.py drive.

Scope: flat disjoint memberships only (numeric / integer / factor / character vectors).  Overlapping or
hierarchical clusterings (R/ECS.R:197-322, 510-807) are out of scope (component 5).
`alpha` is accepted everywhere and ignored, exactly as on the reference's flat path (R/ECS.R:185-190).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi

__all__ = [
    "element_sim", "element_sim_elscore", "element_sim_matrix", "element_consistency",
    "weighted_element_consistency", "element_agreement", "merge_partitions", "merge_identical_partitions",
    "contingency_table", "ResidentPartitions",
]


# ---- label coercion (R: Rcpp input_parameter<IntegerVector>, src/RcppExports.cpp:20-21) -----------------
def _names(x):
    if not (hasattr(x, "to_numpy") and hasattr(x, "index")):  # pandas Series carry R-like names
        return None
    return np.asarray(x.index)


def _is_character(x):
    a = x.to_numpy() if hasattr(x, "to_numpy") else np.asarray(x)
    return a.dtype.kind in "USO" and not _is_factor(x)


def _is_factor(x):
    return hasattr(x, "cat") or type(x).__name__ == "Categorical"


def _as_labels(x):
    """One membership vector -> contiguous int32, the way the reference's .Call boundary coerces it:
    integers as they are, doubles truncated, factors to their codes (1-based), characters to factor
    codes (R/ECS.R:184-185 routes characters to an R loop; SURVEY.md 8f rank 3 folds them into the
    integer path -- any bijective recoding gives the same ECS)."""
    if _is_factor(x):
        codes = np.asarray(x.cat.codes if hasattr(x, "cat") else x.codes)
        if (codes < 0).any():
            raise ValueError("NA labels are not supported")
        return np.ascontiguousarray(codes.astype(np.int32) + 1)
    a = x.to_numpy() if hasattr(x, "to_numpy") else np.asarray(x)
    if a.ndim != 1:
        raise ValueError("a clustering must be a one-dimensional membership vector")
    if a.dtype.kind in "USO":
        if any(v is None for v in a.tolist()):
            raise ValueError("NA labels are not supported")
        _, inv = np.unique(a.astype(str), return_inverse=True)
        return np.ascontiguousarray(inv.astype(np.int32) + 1)
    if a.dtype.kind == "f":
        if np.isnan(a).any():
            raise ValueError("NA labels are not supported")
        return np.ascontiguousarray(np.trunc(a).astype(np.int32))
    if a.dtype.kind in "iub":
        return np.ascontiguousarray(a.astype(np.int32, copy=False))  # int32 input is passed on without a copy
    raise ValueError("unsupported label type %s" % a.dtype)


def _check_pair(c1, c2):
    # R/ECS.R:74-79 and R/ECS.R:177-182
    if len(c1) != len(c2):
        raise ValueError("clustering1 and clustering2 do not have the same length.")
    n1, n2 = _names(c1), _names(c2)
    if n1 is not None and n2 is not None and np.any(n1 != n2):
        raise ValueError("Not all elements of clustering1 and clustering2 are the same.")


def _label_columns(clustering_list):
    """The membership vectors as int32 columns plus the array of their addresses for the *_cols entry points --
    what column_pointers() of r/calculate_ecs_b200.cpp hands over: R's list of vectors as it lies in memory, no
    M x N copy on the host."""
    cols = [_as_labels(x) for x in clustering_list]
    n = cols[0].size if cols else 0
    for c in cols:
        if c.size != n:
            raise ValueError("The partitions do not have the same length")
    ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    return cols, ptrs, len(cols), n


def _wrap(values, like):
    names = _names(like)
    if names is None:
        return values
    import pandas as pd

    return pd.Series(values, index=names)


# ---- per-pair ------------------------------------------------------------------------------------------
def contingency_table(a, b, min_a=None, min_b=None):
    """myContTable (src/calculate_ecs.cpp:7-19): int32 table with rows = labels of a, cols = labels of b,
    plus the cluster sizes (row / column sums)."""
    a, b = _as_labels(a), _as_labels(b)
    if a.size != b.size:
        raise ValueError("clustering1 and clustering2 do not have the same length.")
    L = _capi.lib()
    lo, hi = C.c_int32(), C.c_int32()
    _capi.check(L.ca_label_range(a, a.size, C.byref(lo), C.byref(hi)))
    min_a = lo.value if min_a is None else int(min_a)
    ka = hi.value - min_a + 1
    _capi.check(L.ca_label_range(b, b.size, C.byref(lo), C.byref(hi)))
    min_b = lo.value if min_b is None else int(min_b)
    kb = hi.value - min_b + 1
    tab = np.zeros((kb, ka), dtype=np.int32)  # column-major ka x kb
    sa = np.zeros(ka, dtype=np.int32)
    sb = np.zeros(kb, dtype=np.int32)
    _capi.check(L.ca_contingency(a, b, a.size, min_a, min_b, ka, kb, tab.reshape(-1), sa.ctypes.data, sb.ctypes.data))
    return tab.T, sa, sb


def element_sim_elscore(clustering1, clustering2, alpha=0.9, **_ignored):
    """Per-element ECS of two flat memberships (R/ECS.R:160-195 -> disjointECS)."""
    _check_pair(clustering1, clustering2)
    a, b = _as_labels(clustering1), _as_labels(clustering2)
    out = np.empty(a.size, dtype=np.float64)
    if a.size:
        _capi.check(_capi.lib().ca_ecs_pair(a, b, a.size, out.ctypes.data, None))
    return _wrap(out, clustering1)


def element_sim(clustering1, clustering2, alpha=0.9, **_ignored):
    """Average ECS of two flat memberships (R/ECS.R:59-100): numeric/factor inputs go to
    disjointECSaverage (R/ECS.R:72-82), anything else to mean(element_sim_elscore) (R/ECS.R:84-99)."""
    if not (_is_character(clustering1) or _is_character(clustering2)):
        _check_pair(clustering1, clustering2)
        a, b = _as_labels(clustering1), _as_labels(clustering2)
        out = C.c_double()
        _capi.check(_capi.lib().ca_ecs_pair_mean(a, b, a.size, C.byref(out)))
        return out.value
    return float(np.mean(np.asarray(element_sim_elscore(clustering1, clustering2, alpha))))


# ---- set level -------------------------------------------------------------------------------------------
def weighted_element_consistency(clustering_list, weights=None, calculate_sim_matrix=False):
    """Weighted element-wise consistency (R/ECS.R:1446-1500)."""
    if len(clustering_list) <= 1:
        raise ValueError("At least two clusterings are required to calculate the consistency.")
    cols, ptrs, m, n = _label_columns(clustering_list)
    w = None
    if weights is not None:
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.shape != (m,):
            raise ValueError("weights must have one entry per clustering")
    ecc = np.zeros(n, dtype=np.float64)
    sim = np.zeros((m, m), dtype=np.float64) if calculate_sim_matrix else None
    _capi.check(_capi.lib().ca_consistency_cols(ptrs, n, m, _capi._opt(w), ecc, _capi._opt(sim)))
    ecc = _wrap(ecc, clustering_list[0])
    return {"ecc": ecc, "ecs_matrix": sim} if calculate_sim_matrix else ecc


def element_sim_matrix(clustering_list, output_type="matrix", alpha=0.9, **_ignored):
    """Pairwise mean-ECS matrix (R/ECS.R:849-882 -> element_sim_matrix_flat_disjoint R/ECS.R:931-981)."""
    if output_type not in ("data.frame", "matrix"):
        raise ValueError("output_type must be data.frame or matrix.")
    cols, ptrs, m, n = _label_columns(clustering_list)
    sim = np.zeros((m, m), dtype=np.float64)
    _capi.check(_capi.lib().ca_sim_matrix_cols(ptrs, n, m, sim))
    if output_type == "matrix":
        return sim
    # data.frame(i.clustering = rep(1:m, m), j.clustering = rep(1:m, each = m), element_sim = c(sim))
    cols = {
        "i.clustering": np.tile(np.arange(1, m + 1), m),
        "j.clustering": np.repeat(np.arange(1, m + 1), m),
        "element_sim": sim.T.reshape(-1),  # c() of an R matrix is column-major
    }
    try:
        import pandas as pd

        return pd.DataFrame(cols)
    except ImportError:  # pragma: no cover
        return cols


def merge_identical_partitions(clustering_list):
    """merge_identical_partitions (R/ECS.R:1021-1054) on a list of {mb, freq} records."""
    recs = [dict(r) for r in clustering_list]
    if len(recs) == 1:
        return recs
    cols, ptrs, m, n = _label_columns([r["mb"] for r in recs])
    freq = np.ascontiguousarray([float(r["freq"]) for r in recs], dtype=np.float64)
    keep = np.zeros(m, dtype=np.int32)
    fout = np.zeros(m, dtype=np.float64)
    u = C.c_int32()
    _capi.check(_capi.lib().ca_merge_identical_cols(ptrs, n, m, freq.ctypes.data, keep, fout, C.byref(u)))
    out = []
    for t in range(u.value):
        r = dict(recs[int(keep[t])])
        f = float(fout[t])
        r["freq"] = int(f) if f == int(f) else f
        out.append(r)
    return out


def _merge_partitions_ecs(partition_list, ecs_thresh=0.99):
    """merge_partitions_ecs (R/ECS.R:1057-1135): single-linkage-style merge on the ECS matrix.  Pure
    O(M^2) scalar logic on top of the GPU matrix; restated operation by operation, including the
    column-major which.max and the one-triangle updates."""
    nparts = len(partition_list)
    if nparts == 1:
        return partition_list
    sim = np.array(element_sim_matrix([p["mb"] for p in partition_list]), dtype=np.float64)
    np.fill_diagonal(sim, np.nan)
    groups = {i: [i] for i in range(1, nparts + 1)}  # insertion-ordered like an R named list
    while len(groups) > 1:
        if np.nanmax(sim) < ecs_thresh:
            break
        index = int(np.nanargmax(sim.T.reshape(-1))) + 1  # which.max: first maximum, column-major, 1-based
        first = index % nparts
        second = index // nparts + 1
        if first == 0:
            first = nparts
            second -= 1
        groups[first] = groups[first] + groups[second]
        del groups[second]
        f, s = first - 1, second - 1
        for i1 in list(groups.keys()):
            i = i1 - 1
            if first < i1:
                if not np.isnan(sim[f, i]):
                    sim[f, i] = np.nanmin([sim[f, i], sim[s, i], sim[i, s]])
            else:
                if not np.isnan(sim[i, f]):
                    sim[i, f] = np.nanmin([sim[i, f], sim[s, i], sim[i, s]])
        sim[s, :] = np.nan
        sim[:, s] = np.nan
    merged = []
    for kept, members in groups.items():
        rec = dict(partition_list[kept - 1])
        rec["freq"] = sum(partition_list[x - 1]["freq"] for x in members)
        merged.append(rec)
    return merged


def merge_partitions(partition_list, ecs_thresh=1, order_logic=("freq", "avg_agreement", "none"),
                     return_ecs_matrix=False, check_ties=True):
    """Merge flat partitions whose pairwise ECS exceeds a threshold (R/ECS.R:1166-1311)."""
    if not isinstance(ecs_thresh, (int, float)) or isinstance(ecs_thresh, bool):
        raise ValueError("ecs_thresh parameter should be numeric")
    if not isinstance(order_logic, str):
        order_logic = order_logic[0]
    if order_logic not in ("freq", "avg_agreement", "none"):
        raise ValueError("order parameter should be either `freq`, `avg_agreement` or `none`")

    if isinstance(partition_list, dict):
        if "partitions" in partition_list:  # R/ECS.R:1191-1193
            return merge_partitions(partition_list["partitions"], ecs_thresh, order_logic)
        return {k: merge_partitions(v, ecs_thresh, order_logic) for k, v in partition_list.items()}
    if not isinstance(partition_list[0], dict):  # plain membership vectors (R/ECS.R:1182-1188)
        partition_list = [{"mb": x, "freq": 1} for x in partition_list]
    elif not all(k in partition_list[0] for k in ("mb", "freq")):
        return [merge_partitions(sub, ecs_thresh, order_logic) for sub in partition_list]

    need_matrix = return_ecs_matrix or order_logic in ("freq", "avg_agreement")
    if ecs_thresh == 1 and len(partition_list) > 1:
        # the partitions cross PCIe ONCE: the dedup pre-pass (R/ECS.R:1204) and the consistency of the kept partitions
        # (R/ECS.R:1214-1222) work on the same device-resident label set
        cols, _, m, n = _label_columns([p["mb"] for p in partition_list])
        dev = _capi.DeviceLabels(n, m, primary=True).upload_columns(cols)
        try:
            keep, freq = dev.merge_identical([float(p["freq"]) for p in partition_list])
            part_list = []
            for t, f in zip(keep, freq):
                r = dict(partition_list[int(t)])
                r["freq"] = int(f) if f == int(f) else float(f)
                part_list.append(r)
            if len(part_list) == 1:
                return {"partitions": part_list, "ecc": np.ones(n)}
            # nothing merged and one device: the resident set is used as it is; otherwise the kept partitions are copied
            # device to device (onto every selected GPU when there are several)
            idx = None if (len(keep) == m and _capi.lib().ca_device_count() <= 1) else keep
            ecc, mat = dev.consistency(idx, freq, need_matrix)
        finally:
            dev.close()
        ecc = _wrap(ecc, part_list[0]["mb"])
    else:
        if ecs_thresh == 1:
            part_list = merge_identical_partitions(partition_list)
        else:
            part_list = _merge_partitions_ecs(partition_list, ecs_thresh)
        n = len(part_list[0]["mb"])
        if len(part_list) == 1:
            return {"partitions": part_list, "ecc": np.ones(n)}
        cons = weighted_element_consistency([p["mb"] for p in part_list], weights=[p["freq"] for p in part_list],
                                            calculate_sim_matrix=True)
        ecc, mat = cons["ecc"], cons["ecs_matrix"]
    if order_logic not in ("freq", "avg_agreement"):
        out = {"partitions": part_list, "ecc": ecc}
        if return_ecs_matrix:
            out["ecs_matrix"] = mat
        return out

    u = mat.shape[0]
    n_total = sum(p["freq"] for p in part_list)
    for i in range(u):
        # the reference reads freq from the PRE-merge list here (R/ECS.R:1240,1246); kept as is
        total = (partition_list[i]["freq"] - 1) / (n_total - 1)
        for j in range(u):
            if i != j:
                total += mat[i, j] * partition_list[j]["freq"] / (n_total - 1)
        part_list[i]["avg_agreement"] = total

    order = sorted(range(u), key=lambda i: -part_list[i][order_logic])  # stable, decreasing
    part_list = [part_list[i] for i in order]
    if return_ecs_matrix:
        mat = mat[np.ix_(order, order)]
    check_ties = check_ties and len(part_list) > 1
    if check_ties:
        second_key = "avg_agreement" if order_logic == "freq" else "freq"
        ref_val = part_list[0][order_logic]
        i = len(part_list)
        for t in range(1, len(part_list)):
            if part_list[t][order_logic] != ref_val:
                i = t
                break
        if i > 1:
            ties = sorted(range(i), key=lambda t: -part_list[t][second_key])
            part_list[:i] = [part_list[t] for t in ties]
            if return_ecs_matrix:
                mat = mat.copy()
                mat[:i, :i] = mat[np.ix_(ties, ties)]
    out = {"partitions": part_list, "ecc": ecc}
    if return_ecs_matrix:
        out["ecs_matrix"] = mat
    return out


def element_consistency(clustering_list, alpha=0.9, **_ignored):
    """Element-wise consistency of a set of flat clusterings (R/ECS.R:1350-1386):
    merge_partitions(ecs_thresh = 1, order_logic = "none") and its $ecc."""
    if len(clustering_list) == 1:
        raise ValueError("At least two clusterings are required to calculate the consistency.")
    res = merge_partitions(clustering_list, ecs_thresh=1, order_logic="none", return_ecs_matrix=False)
    return _wrap(np.asarray(res["ecc"]), clustering_list[0])


def element_agreement(reference_clustering, clustering_list, alpha=0.9, **_ignored):
    """Mean ECS of every clustering of a list against a reference (R/ECS.R:1544-1648, flat route
    element_agreement_flat_disjoint R/ECS.R:1625-1648)."""
    ref = _as_labels(reference_clustering)
    cols, ptrs, m, n = _label_columns(clustering_list)
    if n != ref.size:
        raise ValueError("The partitions do not have the same length")
    out = np.zeros(n, dtype=np.float64)
    _capi.check(_capi.lib().ca_agreement_cols(ref, ptrs, n, m, out))
    return out


# ---- pipeline glue: partitions kept on the device across calls ----------------
class ResidentPartitions:
    """The runs of one configuration, uploaded once and compared many times.

    The stability pipeline hands the SAME membership vectors to the ECS routines up to three times: per number
    of clusters k (R/stability-3-graph-clustering.R:271-300), per resolution (:25-69) and per k across
    resolutions (:303-343).  This holds them as one device-resident label set (ca_labels_t); every comparison
    names its partitions by index, nothing is uploaded again.  Results are the ones the list-taking functions of
    this module return for `[clustering_list[i] for i in idx]`.
    """

    def __init__(self, clustering_list):
        if len(clustering_list) < 1:
            raise ValueError("At least one clustering is required.")
        cols, _, m, n = _label_columns(clustering_list)
        self._like = clustering_list[0]
        self._dev = _capi.DeviceLabels(n, m).upload_columns(cols)

    def __len__(self):
        return self._dev.m

    def weighted_element_consistency(self, idx=None, weights=None, calculate_sim_matrix=False):
        """weighted_element_consistency (R/ECS.R:1446-1500) of the partitions idx (default: all)."""
        k = len(self) if idx is None else len(idx)
        if k <= 1:
            raise ValueError("At least two clusterings are required to calculate the consistency.")
        ecc, sim = self._dev.consistency(idx, weights, calculate_sim_matrix)
        ecc = _wrap(ecc, self._like)
        return {"ecc": ecc, "ecs_matrix": sim} if calculate_sim_matrix else ecc

    def element_consistency(self, idx=None):
        """element_consistency (R/ECS.R:1350-1386) of distinct partitions: unit weights."""
        return self.weighted_element_consistency(idx)

    def element_sim_matrix(self, idx=None):
        """element_sim_matrix, output_type = "matrix" (R/ECS.R:931-965)."""
        return self._dev.sim_matrix(idx)

    def close(self):
        self._dev.close()
