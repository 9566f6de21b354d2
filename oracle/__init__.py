"""CPU oracle for the flat-disjoint ECS hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker / reported baseline.  The product
(``clustassess_b200``) never imports it.

Two libraries sit behind it (built by ``oracle/Makefile``):

* ``liboracle.so``        plain-C restatement (``ecs_oracle.c``), each function citing the reference
                          file:line it follows.  Always available.
* ``_ref/libref_ecs.so``  the UNMODIFIED reference ``src/calculate_ecs.cpp`` compiled against a stand-in
                          ``Rcpp.h``.  Built in the container where ``/root/reference`` exists; the
                          prebuilt file travels to the GPU box.  ``have_ref()`` tells whether it loaded.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_I32P = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_F64P = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(ref_root: str = "/root/reference") -> None:
    """Compile liboracle.so and (when the reference sources are present) _ref/libref_ecs.so."""
    subprocess.run(["make", "-s", "-C", _HERE, f"REF={ref_root}"], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


_lib = None
_ref = None


def _ensure():
    global _lib, _ref
    if _lib is None:
        p = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(p):
            build()
        _lib = C.CDLL(p)
        _lib.orc_cont_table_into.argtypes = [_I32P, _I32P, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        _lib.orc_disjoint_ecs.argtypes = [_I32P, _I32P, C.c_int64, _F64P]
        _lib.orc_disjoint_ecs_average.argtypes = [_I32P, _I32P, C.c_int64]
        _lib.orc_disjoint_ecs_average.restype = C.c_double
        _lib.orc_corrected_l1_mb.argtypes = [_I32P, _I32P, C.c_int64, _F64P]
        _lib.orc_weighted_element_consistency.argtypes = [_I32P, C.c_int64, C.c_int32, C.c_void_p, _F64P, C.c_void_p]
        _lib.orc_element_sim_matrix.argtypes = [_I32P, C.c_int64, C.c_int32, _F64P]
        _lib.orc_are_identical_memberships.argtypes = [_I32P, _I32P, C.c_int64]
        _lib.orc_merge_identical_partitions.argtypes = [_I32P, C.c_int64, C.c_int32, C.c_void_p, _I32P, _F64P]
        _lib.orc_merge_identical_partitions.restype = C.c_int32
        _lib.orc_element_consistency.argtypes = [_I32P, C.c_int64, C.c_int32, _F64P]
        _lib.orc_time_pairs.argtypes = [_I32P, C.c_int64, _I32P, _I32P, C.c_int64, C.c_int32, C.c_void_p]
        _lib.orc_time_pairs.restype = C.c_double
        _lib.orc_time_pairs_ex.argtypes = [_I32P, C.c_int64, _I32P, _I32P, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int64, C.c_void_p]
        _lib.orc_time_pairs_ex.restype = C.c_double
        _ref = _load(os.path.join(_HERE, "_ref", "libref_ecs.so"))
        if _ref is not None:
            _ref.ref_myContTable.argtypes = [_I32P, _I32P, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
            _ref.ref_disjointECS.argtypes = [_I32P, _I32P, C.c_int64, _F64P]
            _ref.ref_disjointECSaverage.argtypes = [_I32P, _I32P, C.c_int64]
            _ref.ref_disjointECSaverage.restype = C.c_double
            _ref.ref_time_pairs.argtypes = [_I32P, C.c_int64, _I32P, _I32P, C.c_int64, C.c_int32, C.c_void_p]
            _ref.ref_time_pairs.restype = C.c_double
            _ref.ref_time_pairs_ex.argtypes = [_I32P, C.c_int64, _I32P, _I32P, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_int64, C.c_void_p]
            _ref.ref_time_pairs_ex.restype = C.c_double
    return _lib


def have_ref() -> bool:
    _ensure()
    return _ref is not None


def _i32(x):
    return np.ascontiguousarray(np.asarray(x), dtype=np.int32)


def _mat(labels):
    """list of M vectors / M x N array -> contiguous int32 M x N (partition-major)."""
    a = np.ascontiguousarray(np.asarray(labels), dtype=np.int32)
    if a.ndim != 2:
        raise ValueError("labels must be M x N")
    return a


def _table(fn, a, b, min_a, min_b):
    a, b = _i32(a), _i32(b)
    if min_a is None:
        min_a = int(a.min())
    if min_b is None:
        min_b = int(b.min())
    nr, nc = C.c_int32(), C.c_int32()
    fn(a, b, a.size, min_a, min_b, None, C.byref(nr), C.byref(nc))
    out = np.zeros((nc.value, nr.value), dtype=np.int32)  # column-major storage
    fn(a, b, a.size, min_a, min_b, out.ctypes.data, C.byref(nr), C.byref(nc))
    return out.T  # logical [row=a-label, col=b-label]


# ---- restatement (liboracle.so) ---------------------------------------------------------------
def cont_table(a, b, min_a=None, min_b=None):
    """myContTable (src/calculate_ecs.cpp:7-19): int32 table, rows = a-labels, cols = b-labels."""
    return _table(_ensure().orc_cont_table_into, a, b, min_a, min_b)


def disjoint_ecs(a, b):
    """disjointECS (src/calculate_ecs.cpp:22-46)."""
    a, b = _i32(a), _i32(b)
    out = np.empty(a.size, dtype=np.float64)
    _ensure().orc_disjoint_ecs(a, b, a.size, out)
    return out


def disjoint_ecs_average(a, b):
    """disjointECSaverage (src/calculate_ecs.cpp:49-70), n*max formed without int overflow."""
    a, b = _i32(a), _i32(b)
    return float(_ensure().orc_disjoint_ecs_average(a, b, a.size))


def corrected_l1_mb(a, b):
    """corrected_l1_mb (R/ECS.R:468-507)."""
    a, b = _i32(a), _i32(b)
    out = np.empty(a.size, dtype=np.float64)
    _ensure().orc_corrected_l1_mb(a, b, a.size, out)
    return out


def weighted_element_consistency(labels, weights=None, calculate_sim_matrix=False):
    """weighted_element_consistency (R/ECS.R:1446-1500)."""
    L = _mat(labels)
    m, n = L.shape
    ecc = np.empty(n, dtype=np.float64)
    sim = np.zeros((m, m), dtype=np.float64) if calculate_sim_matrix else None
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    rc = _ensure().orc_weighted_element_consistency(L, n, m, None if w is None else w.ctypes.data, ecc,
                                                    None if sim is None else sim.ctypes.data)
    if rc:
        raise ValueError("At least two clusterings are required to calculate the consistency.")
    return (ecc, sim) if calculate_sim_matrix else ecc


def element_sim_matrix(labels):
    """element_sim_matrix_flat_disjoint (R/ECS.R:931-981), matrix output."""
    L = _mat(labels)
    m, n = L.shape
    sim = np.zeros((m, m), dtype=np.float64)
    _ensure().orc_element_sim_matrix(L, n, m, sim)
    return sim


def are_identical_memberships(a, b):
    """are_identical_memberships (R/ECS.R:984-1016)."""
    a, b = _i32(a), _i32(b)
    return bool(_ensure().orc_are_identical_memberships(a, b, a.size))


def merge_identical_partitions(labels, freq=None):
    """merge_identical_partitions (R/ECS.R:1021-1054) -> (kept indices, merged frequencies)."""
    L = _mat(labels)
    m, n = L.shape
    keep = np.zeros(m, dtype=np.int32)
    fout = np.zeros(m, dtype=np.float64)
    f = None if freq is None else np.ascontiguousarray(freq, dtype=np.float64)
    u = _ensure().orc_merge_identical_partitions(L, n, m, None if f is None else f.ctypes.data, keep, fout)
    return keep[:u].copy(), fout[:u].copy()


def element_consistency(labels):
    """element_consistency (R/ECS.R:1350-1386) through merge_partitions(thresh 1, "none")."""
    L = _mat(labels)
    m, n = L.shape
    ecc = np.empty(n, dtype=np.float64)
    if _ensure().orc_element_consistency(L, n, m, ecc):
        raise ValueError("At least two clusterings are required to calculate the consistency.")
    return ecc


def time_pairs(labels, pi, pj, nthreads=1, impl="port"):
    """Wall seconds for the per-pair body of the ECC loop over the given pairs (bench helper).
    impl="reference" uses the unmodified reference disjointECS from _ref; "port" the restatement."""
    L = _mat(labels)
    m, n = L.shape
    pi, pj = _i32(pi), _i32(pj)
    _ensure()
    if impl == "reference":
        if _ref is None:
            raise RuntimeError("oracle/_ref/libref_ecs.so not built")
        return float(_ref.ref_time_pairs(L, n, pi, pj, pi.size, int(nthreads), None))
    return float(_lib.orc_time_pairs(L, n, pi, pj, pi.size, int(nthreads), None))


def pairs_detail(labels, pi, pj, cells=None, nthreads=1, impl="port"):
    """The per-pair body of the ECC loop (R/ECS.R:1478-1484) over the given pairs, keeping what a parity check needs:
    returns (seconds, sum over the pairs of ECS -- an N-vector, mean(ECS) per pair, ECS per pair at `cells` or None)."""
    L = _mat(labels)
    m, n = L.shape
    pi, pj = _i32(pi), _i32(pj)
    _ensure()
    acc = np.zeros(n, dtype=np.float64)
    means = np.zeros(pi.size, dtype=np.float64)
    ci = None if cells is None else np.ascontiguousarray(cells, dtype=np.int64)
    cv = None if ci is None else np.zeros((pi.size, ci.size), dtype=np.float64)
    fn = _lib.orc_time_pairs_ex
    if impl == "reference":
        if _ref is None:
            raise RuntimeError("oracle/_ref/libref_ecs.so not built")
        fn = _ref.ref_time_pairs_ex
    secs = float(fn(L, n, pi, pj, pi.size, int(nthreads), acc.ctypes.data, means.ctypes.data, None if ci is None else ci.ctypes.data,
                    0 if ci is None else ci.size, None if cv is None else cv.ctypes.data))
    return secs, acc, means, cv


# ---- the unmodified reference (oracle/_ref/libref_ecs.so) ----------------------------------------
def ref_cont_table(a, b, min_a=None, min_b=None):
    _ensure()
    return _table(_ref.ref_myContTable, a, b, min_a, min_b)


def ref_disjoint_ecs(a, b):
    a, b = _i32(a), _i32(b)
    _ensure()
    out = np.empty(a.size, dtype=np.float64)
    _ref.ref_disjointECS(a, b, a.size, out)
    return out


def ref_disjoint_ecs_average(a, b):
    a, b = _i32(a), _i32(b)
    _ensure()
    return float(_ref.ref_disjointECSaverage(a, b, a.size))
