"""The Rcpp glue (route A of INTEGRATION.md): r/calculate_ecs_b200.cpp, the file that replaces the reference's
src/calculate_ecs.cpp, compiled UNMODIFIED against a stand-in Rcpp (tests/rcpp_standin/) and called through a small
C harness that plays BEGIN_RCPP / END_RCPP (a C++ exception of the glue = an R condition).

CPU part (-m "not gpu"): the glue compiles and links against the library's header; argument errors are raised with the
reference's messages before the device is touched; without a device every routine ends in an R condition carrying
the library's message (no CPU fallback).  GPU part (-m gpu): every exported routine against the oracle.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from clustassess_b200 import _capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STANDIN = os.path.join(ROOT, "tests", "rcpp_standin")
OUT = os.path.join(STANDIN, "_build", "libglue_test.so")
TOL = 1e-9


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


class Glue:
    def __init__(self, path):
        self.L = C.CDLL(path)
        P, I, X = C.c_void_p, C.c_int, C.c_ssize_t  # pointers must be declared: ctypes would pass bare ints as 32-bit
        for name, res, args in [
            ("gh_error", C.c_char_p, []), ("gh_release", None, []),
            ("gh_cont_table", I, [P, P, X, I, I, P, I, P, P]),
            ("gh_disjoint_ecs", I, [P, X, P, X, P]), ("gh_disjoint_ecs_average", I, [P, X, P, X, P]),
            ("gh_consistency", I, [P, P, I, I, P, I, I, P, P]), ("gh_sim_matrix", I, [P, P, I, I, P]),
            ("gh_agreement", I, [P, X, P, P, I, P]), ("gh_merge_identical", I, [P, P, I, P, I, P, P, P]),
            ("gh_resident", I, [P, P, I, P, I, P, I, I, P, P, P]),
        ]:
            f = getattr(self.L, name)
            f.restype, f.argtypes = res, args

    @staticmethod
    def _cols(cols):
        arrs = [np.ascontiguousarray(c, dtype=np.int32) for c in cols]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        lens = (C.c_ssize_t * len(arrs))(*[a.size for a in arrs])
        return arrs, ptrs, lens

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.gh_error().decode())

    def cont_table(self, a, b, min_a, min_b):
        a, b = np.ascontiguousarray(a, dtype=np.int32), np.ascontiguousarray(b, dtype=np.int32)
        out = np.zeros(1 << 16, dtype=np.int32)
        ka, kb = C.c_int(), C.c_int()
        self._check(self.L.gh_cont_table(a.ctypes.data, b.ctypes.data, a.size, int(min_a), int(min_b), out.ctypes.data,
                                         out.size, C.byref(ka), C.byref(kb)))
        return out[: ka.value * kb.value].reshape((ka.value, kb.value), order="F")  # IntegerMatrix is column-major

    def disjoint_ecs(self, a, b):
        a, b = np.ascontiguousarray(a, dtype=np.int32), np.ascontiguousarray(b, dtype=np.int32)
        out = np.zeros(max(a.size, 1))
        self._check(self.L.gh_disjoint_ecs(a.ctypes.data, a.size, b.ctypes.data, b.size, out.ctypes.data))
        return out[: a.size]

    def disjoint_ecs_average(self, a, b):
        a, b = np.ascontiguousarray(a, dtype=np.int32), np.ascontiguousarray(b, dtype=np.int32)
        out = C.c_double()
        self._check(self.L.gh_disjoint_ecs_average(a.ctypes.data, a.size, b.ctypes.data, b.size, C.byref(out)))
        return out.value

    def consistency(self, cols, weights=None, want_matrix=False, real_every=0):
        arrs, ptrs, lens = self._cols(cols)
        m, n = len(arrs), arrs[0].size
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        ecc, sim = np.zeros(max(n, 1)), np.zeros((m, m))
        self._check(self.L.gh_consistency(ptrs, lens, m, real_every, None if w is None else w.ctypes.data,
                                          0 if w is None else w.size, int(want_matrix), ecc.ctypes.data, sim.ctypes.data))
        return ecc[:n], (sim if want_matrix else None)

    def sim_matrix(self, cols, real_every=0):
        arrs, ptrs, lens = self._cols(cols)
        sim = np.zeros((len(arrs), len(arrs)))
        self._check(self.L.gh_sim_matrix(ptrs, lens, len(arrs), real_every, sim.ctypes.data))
        return sim

    def agreement(self, ref, cols):
        arrs, ptrs, lens = self._cols(cols)
        ref = np.ascontiguousarray(ref, dtype=np.int32)
        out = np.zeros(max(ref.size, 1))
        self._check(self.L.gh_agreement(ref.ctypes.data, ref.size, ptrs, lens, len(arrs), out.ctypes.data))
        return out[: ref.size]

    def merge_identical(self, cols, freq):
        arrs, ptrs, lens = self._cols(cols)
        f = np.ascontiguousarray(freq, dtype=np.float64)
        keep, fout, u = np.zeros(len(arrs), np.int32), np.zeros(len(arrs)), C.c_int()
        self._check(self.L.gh_merge_identical(ptrs, lens, len(arrs), f.ctypes.data, f.size, keep.ctypes.data, fout.ctypes.data, C.byref(u)))
        return keep[: u.value], fout[: u.value]

    def resident(self, cols, idx, weights=None):
        arrs, ptrs, lens = self._cols(cols)
        ix = np.ascontiguousarray(idx, dtype=np.int32)
        k, n = ix.size, arrs[0].size
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        ecc, sim, sim_only = np.zeros(n), np.zeros((k, k)), np.zeros((k, k))
        self._check(self.L.gh_resident(ptrs, lens, len(arrs), ix.ctypes.data, k, None if w is None else w.ctypes.data, 0 if w is None else w.size, 1,
                                       ecc.ctypes.data, sim.ctypes.data, sim_only.ctypes.data))
        return ecc, sim, sim_only

    def release(self):
        self.L.gh_release()


@pytest.fixture(scope="module")
def glue():
    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    _capi.lib()  # the library must have been built (python __graft_entry__.py)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    libdir, libname = os.path.dirname(_capi.LIB_PATH), os.path.basename(_capi.LIB_PATH)
    cmd = [cxx, "-std=c++14", "-O1", "-g", "-Wall", "-fPIC", "-shared", "-I" + STANDIN, "-I" + os.path.join(ROOT, "include"), "-o", OUT,
           os.path.join(STANDIN, "glue_harness.cpp"), "-L" + libdir, "-l" + libname[3:-3], "-Wl,-rpath," + libdir]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, "r/calculate_ecs_b200.cpp does not compile against the Rcpp stand-in:\n" + p.stderr
    g = Glue(OUT)
    yield g
    g.release()


# ---- CPU ----------------------------------------------------------------------------------------------------------
def test_glue_exports_the_reference_routines():
    """The three routines of src/calculate_ecs.cpp keep name and signature in the glue, so that
    Rcpp::compileAttributes() regenerates R/RcppExports.R:4-14 unchanged; the batched ones are additional exports."""
    src = open(os.path.join(ROOT, "r", "calculate_ecs_b200.cpp")).read()
    exported = [ln for ln in src.split("// [[Rcpp::export]]\n")[1:]]
    sigs = [e.split("{")[0].strip() for e in exported]
    assert "IntegerMatrix myContTable(IntegerVector &a, IntegerVector &b, int minim_mb1, int minim_mb2)" in sigs
    assert "NumericVector disjointECS(IntegerVector mb1, IntegerVector mb2)" in sigs
    assert "double disjointECSaverage(IntegerVector mb1, IntegerVector mb2)" in sigs
    names = {s.split("(")[0].split()[-1] for s in sigs}
    assert {"ecsConsistency", "ecsSimMatrix", "ecsAgreement", "ecsMergeIdentical", "ecsResident", "ecsResidentConsistency",
            "ecsResidentSimMatrix"} <= names


def test_glue_argument_errors(glue):
    with pytest.raises(RuntimeError, match="clustering1 and clustering2 do not have the same length"):  # R/ECS.R:177-178
        glue.disjoint_ecs([1, 2, 3], [1, 2])
    with pytest.raises(RuntimeError, match="do not have the same length"):
        glue.disjoint_ecs_average([1, 2, 3], [1, 2])
    with pytest.raises(RuntimeError, match="At least two clusterings are required"):  # R/ECS.R:1451-1452
        glue.consistency([[1, 2, 3]])
    with pytest.raises(RuntimeError, match="The partitions do not have the same length"):
        glue.consistency([[1, 2, 3], [1, 2]])
    with pytest.raises(RuntimeError, match="weights must have one entry per clustering"):
        glue.consistency([[1, 2, 3], [1, 2, 2]], weights=[1.0])
    with pytest.raises(RuntimeError, match="The partitions do not have the same length"):
        glue.agreement([1, 2], [[1, 2, 3], [1, 2, 2]])
    with pytest.raises(RuntimeError, match="freq must have one entry per clustering"):  # a short freq would be read out of bounds
        glue.merge_identical([[1, 2, 3], [1, 2, 2]], [1.0])
    glue.release()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_glue_has_no_cpu_fallback(glue):
    two = [[1, 1, 2], [2, 2, 2]]
    for call in (lambda: glue.disjoint_ecs(*two), lambda: glue.disjoint_ecs_average(*two), lambda: glue.cont_table(two[0], two[1], 1, 2),
                 lambda: glue.consistency(two, want_matrix=True), lambda: glue.sim_matrix(two), lambda: glue.agreement(two[0], two),
                 lambda: glue.merge_identical(two, [1.0, 1.0]), lambda: glue.resident(two, [1, 2])):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    glue.release()


# ---- GPU: every exported routine against the oracle -------------------------------------------------------------------
@pytest.mark.gpu
def test_glue_per_pair_routines(glue):
    (a, b), _ = synth.config("C2", n=20_000)
    assert np.array_equal(glue.disjoint_ecs(a, b), oracle.disjoint_ecs(a, b))
    assert abs(glue.disjoint_ecs_average(a, b) - oracle.disjoint_ecs_average(a, b)) < 1e-12
    assert np.array_equal(glue.cont_table(a, b, a.min(), b.min()), oracle.cont_table(a, b))
    glue.release()


@pytest.mark.gpu
def test_glue_batched_routines(glue):
    L, w = synth.config("C3", m=7, n=15_013)
    cols = [L[p] for p in range(7)]
    ecc_ref, sim_ref = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    ecc, sim = glue.consistency(cols, weights=w, want_matrix=True, real_every=3)  # every third vector arrives as a REALSXP
    assert np.abs(ecc - ecc_ref).max() < TOL and np.abs(sim - sim_ref).max() < TOL
    ecc_u, none = glue.consistency(cols)
    assert none is None and np.abs(ecc_u - oracle.weighted_element_consistency(L)).max() < TOL
    assert np.abs(glue.sim_matrix(cols, real_every=2) - oracle.element_sim_matrix(L)).max() < TOL
    want = np.mean([oracle.corrected_l1_mb(L[0], L[p]) for p in range(1, 7)], axis=0)  # R/ECS.R:1639-1647
    assert np.abs(glue.agreement(L[0], cols[1:]) - want).max() < TOL
    dup = [L[0], L[1], L[0].max() + 1 - L[0], L[2], L[1]]  # 0 == 2, 1 == 4
    keep, freq = glue.merge_identical(dup, [1, 1, 1, 1, 1])
    assert keep.tolist() == [1, 2, 4] and freq.tolist() == [2.0, 2.0, 1.0]  # 1-based indices for R
    # resident set: uploaded once, compared on a subset by (1-based) index
    idx = [2, 4, 5, 7]
    sub = L[[1, 3, 4, 6]]
    ecc_r, sim_r, sim_only = glue.resident(cols, idx, weights=w[[1, 3, 4, 6]])
    ecc_s, sim_s = oracle.weighted_element_consistency(sub, w[[1, 3, 4, 6]], calculate_sim_matrix=True)
    assert np.abs(ecc_r - ecc_s).max() < TOL and np.abs(sim_r - sim_s).max() < TOL and np.abs(sim_only - sim_s).max() < TOL
    assert ecc_r.size == L.shape[1]  # sized from the label set (ca_labels_shape), not from a caller-supplied length
    with pytest.raises(RuntimeError, match="outside the resident label set"):
        glue.resident(cols, [2, 9])
    with pytest.raises(RuntimeError, match="weights must have one entry per clustering"):
        glue.resident(cols, [1, 2, 3], weights=[1.0, 2.0])
    glue.release()
