"""CPU tests of everything that does not need a GPU: the C-ABI library loads and exports every symbol
include/clustassess_b200.h declares, the host-side planner keeps its invariants, argument errors are
raised with the reference's messages, and compute calls FAIL LOUDLY without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from clustassess_b200 import _capi, ecs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "clustassess_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ca_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _capi.lib()
    for name in declared:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    assert declared == set(_capi.SYMBOLS), (declared ^ set(_capi.SYMBOLS))
    assert lib.ca_version() >= 100


def _plan(sizes, align=4, cap=2048, max_rows=7):
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    k = sizes.size
    cs = np.zeros(k, np.int32)
    cr = np.zeros(k, np.int32)
    items = np.zeros(4 * (k + 1), np.int32)
    npos = C.c_int64()
    ni = _capi.lib().ca_plan_blocks(sizes, k, align, cap, max_rows, cs, cr, items, k + 1, C.byref(npos))
    assert ni >= 0, _capi.lib().ca_last_error()
    return cs, cr, items[: 4 * ni].reshape(-1, 4), npos.value


@pytest.mark.parametrize("seed", list(range(12)))
def test_planner_invariants(seed):
    rng = np.random.default_rng(seed)
    k = int(rng.integers(1, 400))
    sizes = np.maximum(1, rng.lognormal(4.0, 1.6, size=k).astype(np.int64))
    if seed == 3:
        sizes[0] = 70000  # a cluster that needs 32-bit counters
    align, cap, max_rows = 4, 2048, int(rng.integers(1, 12))
    cs, cr, items, npos = _plan(sizes, align, cap, max_rows)
    pad = (sizes + align - 1) // align * align
    # segments are aligned, disjoint and cover [0, npos)
    order = np.argsort(cs)
    assert (cs % align == 0).all()
    assert (cs[order][1:] == (cs + pad)[order][:-1]).all() and cs[order][0] == 0
    assert npos == pad.sum()
    # every cluster lies in exactly one item; rows are 0..nrows-1 inside it; limits hold
    covered = 0
    for pos0, np_, nrows, wide in items:
        inside = np.where((cs >= pos0) & (cs < pos0 + np_))[0]
        assert inside.size == nrows and pad[inside].sum() == np_
        assert sorted(cr[inside].tolist()) == list(range(nrows))
        assert nrows <= max_rows
        if np_ > cap:
            assert nrows == 1
        assert bool(wide) == bool(sizes[inside].max() > 65535 and np_ > cap)
        covered += nrows
    assert covered == k
    # best-fit-decreasing keeps resident blocks reasonably full
    res = items[items[:, 1] <= cap]
    small = pad[pad <= cap]
    lower = max(-(-int(small.sum()) // cap), -(-small.size // max_rows))
    assert res.shape[0] <= 2 * lower + 1


def test_planner_rejects_bad_input():
    cs = np.zeros(2, np.int32)
    items = np.zeros(12, np.int32)
    npos = C.c_int64()
    bad = np.array([5, 0], dtype=np.int32)
    assert _capi.lib().ca_plan_blocks(bad, 2, 4, 2048, 4, cs, cs.copy(), items, 3, C.byref(npos)) == _capi.CA_E_ARG
    assert b"size" in _capi.lib().ca_last_error()


def test_argument_errors_match_the_reference():
    # R/ECS.R:177-178, 1374-1376, 1451-1452, 861-863
    with pytest.raises(ValueError, match="do not have the same length"):
        ecs.element_sim_elscore([1, 2, 3], [1, 2])
    with pytest.raises(ValueError, match="do not have the same length"):
        ecs.element_sim([1, 2, 3], [1, 2])
    with pytest.raises(ValueError, match="At least two clusterings"):
        ecs.element_consistency([[1, 2, 3]])
    with pytest.raises(ValueError, match="At least two clusterings"):
        ecs.weighted_element_consistency([[1, 2, 3]])
    with pytest.raises(ValueError, match="output_type must be"):
        ecs.element_sim_matrix([[1, 2], [2, 1]], output_type="list")
    with pytest.raises(ValueError, match="ecs_thresh"):
        ecs.merge_partitions([[1, 2], [2, 1]], ecs_thresh="a")
    with pytest.raises(ValueError, match="order parameter"):
        ecs.merge_partitions([[1, 2], [2, 1]], order_logic="size")
    import pandas as pd

    with pytest.raises(ValueError, match="Not all elements"):
        ecs.element_sim_elscore(pd.Series([1, 2], index=["a", "b"]), pd.Series([1, 2], index=["a", "c"]))


def test_label_coercion():
    import pandas as pd

    assert ecs._as_labels([1.9, -2.9, 3.0]).tolist() == [1, -2, 3]  # doubles truncate like Rcpp's int cast
    assert ecs._as_labels(pd.Categorical(["b", "a", "b"])).tolist() == [2, 1, 2]
    assert ecs._as_labels(["x", "y", "x"]).tolist() == [1, 2, 1]
    with pytest.raises(ValueError, match="NA"):
        ecs._as_labels([1.0, float("nan")])


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_a_device():
    with pytest.raises(_capi.ClustAssessError, match="no CPU fallback"):
        ecs.element_sim_elscore([1, 1, 2], [2, 2, 2])
    with pytest.raises(_capi.ClustAssessError):
        ecs.element_consistency([[1, 1, 2], [2, 2, 2]])


def test_span_sharding_of_the_in_library_multi_gpu_path():
    """One process, several GPUs: partitions are dealt to devices in 32-partition spans, longest processing
    time first on the pair counts sum_{i in span} (m - 1 - i).  Every span has exactly one owner, every device gets work when there
    are enough spans, and the pair counts are balanced -- at config 5 (500 partitions, 8 devices) within 1 %."""
    lib = _capi.lib()

    def owners(m, g):
        out = np.zeros((m + 31) // 32, dtype=np.int32)
        ns = lib.ca_plan_spans(m, g, out, out.size)
        assert ns == out.size
        return out

    def loads(m, g):
        o = owners(m, g)
        load = np.zeros(g, dtype=np.int64)
        for i in range(m - 1):
            load[o[i // 32]] += m - 1 - i
        return o, load

    o, load = loads(500, 8)
    assert sorted(set(o.tolist())) == list(range(8)) and load.sum() == 500 * 499 // 2
    assert load.max() / load.mean() < 1.01
    assert all(sorted(np.flatnonzero(o == d).tolist()) == [d, 15 - d] for d in range(8))  # the zigzag of ca_shard_owner, at span granularity
    for m, g, bound in [(1000, 8, 1.01), (1000, 4, 1.01), (500, 2, 1.01), (257, 3, 1.1), (100, 2, 1.1), (64, 2, 1.6)]:
        o, load = loads(m, g)  # two spans on two devices (64, 2) cannot be balanced at span granularity: 3 : 1 pairs
        assert o.min() >= 0 and o.max() < g and load.sum() == m * (m - 1) // 2
        assert load.max() / load.mean() < bound, (m, g, load)
    assert owners(33, 8).tolist()[0] != owners(33, 8).tolist()[1]  # two spans, two devices; the other six stay idle
    assert lib.ca_plan_spans(0, 2, np.zeros(1, dtype=np.int32), 1) < 0 and lib.ca_plan_spans(100, 99, np.zeros(4, dtype=np.int32), 4) < 0


def test_route_through_the_engine_from_cluster_count_and_cells(monkeypatch):
    """Host logic behind the batched entries (csrc/ca_api.cu, route_for): with few clusters the ECS matrix is summed over the
    contingency-table entries -- the reference's own formula for a pair's average, src/calculate_ecs.cpp:58-67 -- instead of per
    (cell, partner); few HUGE clusters switch the consistency to ratio tables.  The decisions measured as the faster ones in
    profiles/r2_table_routes.txt, the environment overrides, and the rule that ratio tables need ecc."""
    import ctypes as C

    lib = _capi.lib()
    monkeypatch.delenv("CA_RT", raising=False)
    monkeypatch.delenv("CA_SIM_SCAN", raising=False)

    def route(kmax, n, ecc, sim):
        rt, scan = C.c_int(-1), C.c_int(-1)
        _capi.check(lib.ca_plan_route(kmax, n, int(ecc), int(sim), C.byref(rt), C.byref(scan)))
        return bool(rt.value), bool(scan.value)

    # (kmax, n): matrix only / ecc + matrix / ecc only
    assert route(60, 50_000, False, True) == (False, True)        # config 4, element_sim_matrix: no gather at all
    assert route(40, 100_000, True, True) == (False, True)        # config 3 with the matrix
    assert route(2000, 1_000_000, True, True) == (False, False)   # config 5: long sparse rows, per cell
    assert route(2000, 1_000_000, False, True) == (False, False)
    assert route(2000, 1_000_000, True, False) == (False, False)  # the headline job: neither
    assert route(12, 2700, False, True) == (False, True) and route(12, 2700, True, True) == (False, False)  # config 1
    assert route(64, 1_000_000, True, False) == (True, False)     # few huge clusters: ratio tables (which also sum the matrix)
    assert route(64, 1_000_000, True, True) == (True, False)
    assert route(64, 1_000_000, False, True) == (False, True)     # ... but never without ecc
    assert route(128, 1_000_000, True, False) == (False, False) and route(60, 50_000, True, False) == (False, False)
    assert route(0, 0, True, True)[0] is False
    monkeypatch.setenv("CA_RT", "1")
    assert route(2000, 1_000_000, True, True) == (True, False) and route(2000, 1_000_000, False, True)[0] is False
    monkeypatch.setenv("CA_RT", "0")
    monkeypatch.setenv("CA_SIM_SCAN", "1")
    assert route(64, 1_000_000, True, True) == (False, True) and route(2000, 1_000_000, True, False) == (False, False)
    monkeypatch.setenv("CA_SIM_SCAN", "0")
    assert route(60, 50_000, False, True) == (False, False)
    rt, scan = C.c_int(), C.c_int()
    assert lib.ca_plan_route(-1, 10, 1, 1, C.byref(rt), C.byref(scan)) < 0
