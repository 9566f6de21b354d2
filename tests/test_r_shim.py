"""The R-side binding itself (route B of INTEGRATION.md): r/clustassess_shim.c, the .Call entry points a maintainer
adds to the package, compiled against a miniature of R's C runtime (tests/r_standin/: SEXPs, PROTECT stack, Rf_error
as a longjmp) and driven the way R's .Call would drive them.

CPU part (-m "not gpu"): the shim compiles and links against the library, registers the reference's three symbol
names (src/RcppExports.cpp:130-132) plus the batched ones with the right arities, raises the reference's argument
errors (R/ECS.R:177-178, 1451-1452) as R conditions before touching the device, and turns the library's "no device"
code into an R error instead of computing anything on the CPU.
GPU part (-m gpu): results through the shim against the oracle -- INTSXP / REALSXP / factor inputs, list inputs,
weights, matrix shapes and names -- with the PROTECT stack balanced after every call and the inputs untouched.
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
STANDIN = os.path.join(ROOT, "tests", "r_standin")
OUT = os.path.join(STANDIN, "_build", "libshim_test.so")
TOL = 1e-9
NILSXP, LGLSXP, INTSXP, REALSXP, VECSXP = 0, 10, 13, 14, 19


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


class MiniR:
    """ctypes view of the shim + mini R harness."""

    def __init__(self, path):
        L = self.L = C.CDLL(path)
        P = C.c_void_p
        for name, res, args in [
            ("mr_reset", None, []), ("mr_null", P, []), ("mr_int", P, [P, C.c_ssize_t]), ("mr_factor", P, [P, C.c_ssize_t]),
            ("mr_real", P, [P, C.c_ssize_t]), ("mr_lgl", P, [C.c_int]), ("mr_list", P, [C.c_ssize_t]),
            ("mr_list_set", None, [P, C.c_ssize_t, P]), ("mr_call", P, [P, C.c_int, P, P, P, P]), ("mr_error", C.c_char_p, []),
            ("mr_protect_depth", C.c_int, []), ("mr_protect_underflow", C.c_int, []), ("mr_type", C.c_int, [P]),
            ("mr_len", C.c_ssize_t, [P]), ("mr_data", P, [P]), ("mr_elt", P, [P, C.c_ssize_t]), ("mr_name", C.c_char_p, [P, C.c_ssize_t]),
            ("mr_dim", C.c_int, [P, C.c_int]), ("mr_finalize", None, [P]), ("mr_finalizers_run", C.c_int, []),
        ]:
            f = getattr(L, name)
            f.restype, f.argtypes = res, args

    # ---- R values ----
    def int(self, v):
        a = np.ascontiguousarray(v, dtype=np.int32)
        return self.L.mr_int(a.ctypes.data, a.size)

    def factor(self, codes):
        a = np.ascontiguousarray(codes, dtype=np.int32)
        return self.L.mr_factor(a.ctypes.data, a.size)

    def real(self, v):
        a = np.ascontiguousarray(v, dtype=np.float64)
        return self.L.mr_real(a.ctypes.data, a.size)

    def list(self, items):
        l = self.L.mr_list(len(items))
        for i, x in enumerate(items):
            self.L.mr_list_set(l, i, x)
        return l

    def null(self):
        return self.L.mr_null()

    def lgl(self, v):
        return self.L.mr_lgl(int(v))

    # ---- .Call ----
    def call(self, symbol, *args):
        """Returns the result SEXP; raises RuntimeError(message) when the entry raised an R error."""
        fn = C.cast(getattr(self.L, symbol), C.c_void_p)
        a = list(args) + [None] * (4 - len(args))
        res = self.L.mr_call(fn, len(args), *a)
        assert not self.L.mr_protect_underflow(), "UNPROTECT of more items than were protected"
        if not res:
            raise RuntimeError(self.L.mr_error().decode())
        assert self.L.mr_protect_depth() == 0, "PROTECT stack not balanced on return: %d" % self.L.mr_protect_depth()
        return res

    def values(self, x):
        t, n = self.L.mr_type(x), self.L.mr_len(x)
        if t == NILSXP:
            return None
        if t == VECSXP:
            return {self.L.mr_name(x, i).decode(): self.values(self.L.mr_elt(x, i)) for i in range(n)}
        ct = {INTSXP: C.c_int32, LGLSXP: C.c_int32, REALSXP: C.c_double}[t]
        a = np.ctypeslib.as_array(C.cast(self.L.mr_data(x), C.POINTER(ct)), shape=(max(n, 1),))[:n].copy()
        if self.L.mr_dim(x, 0) >= 0:
            a = a.reshape((self.L.mr_dim(x, 0), self.L.mr_dim(x, 1)), order="F")  # R matrices are column-major
        return a

    def reset(self):
        self.L.mr_reset()


@pytest.fixture(scope="module")
def R():
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    _capi.lib()  # the library must have been built (python __graft_entry__.py)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    libdir = os.path.dirname(_capi.LIB_PATH)
    libname = os.path.basename(_capi.LIB_PATH)
    assert libname.startswith("lib") and libname.endswith(".so")
    cmd = [cc, "-O1", "-g", "-Wall", "-fPIC", "-shared", "-I" + STANDIN, "-I" + os.path.join(ROOT, "include"), "-o", OUT,
           os.path.join(ROOT, "r", "clustassess_shim.c"), os.path.join(STANDIN, "mini_r.c"),
           "-L" + libdir, "-l" + libname[3:-3], "-Wl,-rpath," + libdir, "-lm"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, "the shim does not compile against the R C API stand-in:\n" + p.stderr
    r = MiniR(OUT)
    yield r
    r.reset()


# ---- CPU ----------------------------------------------------------------------------------------------------------
def test_shim_registers_the_reference_symbols(R):
    class Def(C.Structure):
        _fields_ = [("name", C.c_char_p), ("fun", C.c_void_p), ("numArgs", C.c_int)]

    table = (Def * 11).in_dll(R.L, "ca_shim_entries")
    got = {}
    for d in table:
        if d.name is None:
            break
        got[d.name.decode()] = d.numArgs
        assert d.fun == C.cast(getattr(R.L, d.name.decode()), C.c_void_p).value
    # src/RcppExports.cpp:130-132: the three names and arities the package registers today
    assert got["_ClustAssess_myContTable"] == 4 and got["_ClustAssess_disjointECS"] == 2 and got["_ClustAssess_disjointECSaverage"] == 2
    assert got["_ClustAssess_ecs_consistency"] == 3 and got["_ClustAssess_ecs_sim_matrix"] == 1
    # the same set of batched entries route A exports (r/calculate_ecs_b200.cpp): agreement, dedup, resident label sets
    assert got["_ClustAssess_ecs_agreement"] == 2 and got["_ClustAssess_ecs_merge_identical"] == 2
    assert got["_ClustAssess_ecs_resident"] == 1 and got["_ClustAssess_ecs_resident_consistency"] == 4
    assert got["_ClustAssess_ecs_resident_sim_matrix"] == 2 and len(got) == 10
    assert table[len(got)].name is None  # NULL-terminated like R_CallMethodDef tables must be


def test_shim_raises_argument_errors_as_r_conditions(R):
    R.reset()
    with pytest.raises(RuntimeError, match="clustering1 and clustering2 do not have the same length"):  # R/ECS.R:177-178
        R.call("_ClustAssess_disjointECS", R.int([1, 2, 3]), R.int([1, 2]))
    with pytest.raises(RuntimeError, match="do not have the same length"):
        R.call("_ClustAssess_disjointECSaverage", R.real([1, 2, 3]), R.int([1, 2]))
    with pytest.raises(RuntimeError, match="At least two clusterings are required"):  # R/ECS.R:1451-1452
        R.call("_ClustAssess_ecs_consistency", R.list([R.int([1, 2, 3])]), R.null(), R.lgl(False))
    with pytest.raises(RuntimeError, match="The partitions do not have the same length"):
        R.call("_ClustAssess_ecs_consistency", R.list([R.int([1, 2, 3]), R.int([1, 2])]), R.null(), R.lgl(False))
    with pytest.raises(RuntimeError, match="The partitions do not have the same length"):
        R.call("_ClustAssess_ecs_sim_matrix", R.list([R.int([1, 2, 3]), R.real([1, 2])]))
    two = [R.int([1, 2, 3]), R.int([1, 2, 2])]
    with pytest.raises(RuntimeError, match="weights must have one entry per clustering"):  # a short vector would be read out of bounds
        R.call("_ClustAssess_ecs_consistency", R.list(two), R.real([1.0]), R.lgl(False))
    with pytest.raises(RuntimeError, match="freq must have one entry per clustering"):
        R.call("_ClustAssess_ecs_merge_identical", R.list(two), R.real([1.0, 1.0, 1.0]))
    with pytest.raises(RuntimeError, match="The partitions do not have the same length"):
        R.call("_ClustAssess_ecs_agreement", R.int([1, 2]), R.list(two))
    with pytest.raises(RuntimeError, match="not a resident label set"):
        R.call("_ClustAssess_ecs_resident_sim_matrix", R.int([1]), R.int([1, 2]))
    R.reset()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_shim_has_no_cpu_fallback(R):
    R.reset()
    for sym, args in [("_ClustAssess_disjointECS", (R.int([1, 1, 2]), R.int([2, 2, 2]))),
                      ("_ClustAssess_disjointECSaverage", (R.int([1, 1, 2]), R.int([2, 2, 2]))),
                      ("_ClustAssess_myContTable", (R.int([1, 1, 2]), R.int([2, 2, 2]), R.int([1]), R.int([2]))),
                      ("_ClustAssess_ecs_consistency", (R.list([R.int([1, 1, 2]), R.int([2, 2, 2])]), R.null(), R.lgl(True))),
                      ("_ClustAssess_ecs_sim_matrix", (R.list([R.int([1, 1, 2]), R.int([2, 2, 2])]),)),
                      ("_ClustAssess_ecs_agreement", (R.int([1, 1, 2]), R.list([R.int([1, 1, 2]), R.int([2, 2, 2])]))),
                      ("_ClustAssess_ecs_merge_identical", (R.list([R.int([1, 1, 2]), R.int([2, 2, 2])]), R.real([1.0, 1.0]))),
                      ("_ClustAssess_ecs_resident", (R.list([R.int([1, 1, 2]), R.int([2, 2, 2])]),))]:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            R.call(sym, *args)
    R.reset()


# ---- GPU: parity through the .Call entry points ------------------------------------------------------------------------
@pytest.mark.gpu
def test_shim_per_pair_routines_match_the_reference(R):
    R.reset()
    (a, b), _ = synth.config("C2", n=20_000)
    sa, sb = R.int(a), R.int(b)
    ecs = R.values(R.call("_ClustAssess_disjointECS", sa, sb))
    assert np.array_equal(ecs, oracle.disjoint_ecs(a, b))  # one IEEE division per element: bit-identical
    assert np.array_equal(R.values(sa), a) and np.array_equal(R.values(sb), b)  # inputs are never written
    # REALSXP inputs are truncated like Rcpp's input_parameter<IntegerVector> (src/RcppExports.cpp:34-35); factors are their codes
    ecs_real = R.values(R.call("_ClustAssess_disjointECS", R.real(a + 0.9), R.factor(b)))
    assert np.array_equal(ecs_real, ecs)
    avg = R.values(R.call("_ClustAssess_disjointECSaverage", sa, sb))
    assert avg.shape == (1,) and abs(avg[0] - oracle.disjoint_ecs_average(a, b)) < 1e-12
    tab = R.values(R.call("_ClustAssess_myContTable", sa, sb, R.int([a.min()]), R.int([b.min()])))
    want = oracle.cont_table(a, b)
    assert tab.dtype == np.int32 and np.array_equal(tab, want[0] if isinstance(want, tuple) else want)
    R.reset()


@pytest.mark.gpu
def test_shim_batched_routines_match_the_reference(R):
    R.reset()
    L, w = synth.config("C3", m=7, n=15_013)
    ecc_ref, sim_ref = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    cols = [R.int(L[p]) if p % 2 == 0 else R.real(L[p].astype(np.float64)) for p in range(7)]  # mixed INTSXP / REALSXP list
    res = R.values(R.call("_ClustAssess_ecs_consistency", R.list(cols), R.real(w), R.lgl(True)))
    assert list(res) == ["ecc", "ecs_matrix"]
    assert res["ecc"].shape == (15_013,) and np.abs(res["ecc"] - ecc_ref).max() < TOL
    assert res["ecs_matrix"].shape == (7, 7) and np.abs(res["ecs_matrix"] - sim_ref).max() < TOL
    # integer weights are coerced, NULL weights mean 1, no matrix -> NULL in its slot (R/ECS.R:1466-1468)
    res_i = R.values(R.call("_ClustAssess_ecs_consistency", R.list(cols), R.int(w.astype(np.int32)), R.lgl(False)))
    assert res_i["ecs_matrix"] is None and np.abs(res_i["ecc"] - ecc_ref).max() < TOL
    res_u = R.values(R.call("_ClustAssess_ecs_consistency", R.list(cols), R.null(), R.lgl(False)))
    assert np.abs(res_u["ecc"] - oracle.weighted_element_consistency(L)).max() < TOL
    sim = R.values(R.call("_ClustAssess_ecs_sim_matrix", R.list(cols)))
    assert sim.shape == (7, 7) and np.abs(sim - oracle.element_sim_matrix(L)).max() < TOL
    assert np.array_equal(sim, sim.T) and np.all(np.diag(sim) == 1.0)
    # an NA in the input is an R error carrying the library's message, not undefined behaviour
    bad = L[0].copy()
    bad[5] = np.iinfo(np.int32).min
    with pytest.raises(RuntimeError, match="NA"):
        R.call("_ClustAssess_ecs_sim_matrix", R.list([R.int(bad), R.int(L[1])]))
    R.reset()


@pytest.mark.gpu
def test_shim_agreement_dedup_and_resident_sets(R):
    """The entries that bring route B level with route A: element_agreement's flat branch (R/ECS.R:1639-1647),
    merge_identical_partitions (R/ECS.R:1021-1054) and the resident label sets of the pipeline glue."""
    R.reset()
    L, w = synth.config("C3", m=7, n=15_013)
    cols = [R.int(L[p]) for p in range(7)]
    want = np.mean([oracle.corrected_l1_mb(L[0], L[p]) for p in range(1, 7)], axis=0)
    agr = R.values(R.call("_ClustAssess_ecs_agreement", R.real(L[0].astype(np.float64)), R.list(cols[1:])))
    assert agr.shape == (15_013,) and np.abs(agr - want).max() < TOL
    dup = [L[0], L[1], L[0].max() + 1 - L[0], L[2], L[1]]  # 0 == 2, 1 == 4
    res = R.values(R.call("_ClustAssess_ecs_merge_identical", R.list([R.int(x) for x in dup]), R.int([1, 1, 1, 1, 1])))
    assert list(res) == ["keep", "freq"] and res["keep"].tolist() == [1, 2, 4] and res["freq"].tolist() == [2.0, 2.0, 1.0]
    # resident: one upload, comparisons by 1-based index, result length taken from the set itself
    ptr = R.call("_ClustAssess_ecs_resident", R.list(cols))
    idx, sel = [2, 4, 5, 7], [1, 3, 4, 6]
    ecc_s, sim_s = oracle.weighted_element_consistency(L[sel], w[sel], calculate_sim_matrix=True)
    out = R.values(R.call("_ClustAssess_ecs_resident_consistency", ptr, R.int(idx), R.real(w[sel]), R.lgl(True)))
    assert out["ecc"].shape == (15_013,) and np.abs(out["ecc"] - ecc_s).max() < TOL and np.abs(out["ecs_matrix"] - sim_s).max() < TOL
    sim = R.values(R.call("_ClustAssess_ecs_resident_sim_matrix", ptr, R.real([2.0, 4.0, 5.0, 7.0])))
    assert sim.shape == (4, 4) and np.abs(sim - sim_s).max() < TOL
    with pytest.raises(RuntimeError, match="outside the resident label set"):
        R.call("_ClustAssess_ecs_resident_sim_matrix", ptr, R.int([1, 8]))
    with pytest.raises(RuntimeError, match="weights must have one entry per clustering"):
        R.call("_ClustAssess_ecs_resident_consistency", ptr, R.int([1, 2, 3]), R.real([1.0, 2.0]), R.lgl(False))
    # the finalizer releases the device memory exactly once; a released set is an R error, not a crash
    before = R.L.mr_finalizers_run()
    R.L.mr_finalize(ptr)
    R.L.mr_finalize(ptr)
    assert R.L.mr_finalizers_run() == before + 1
    with pytest.raises(RuntimeError, match="has been released"):
        R.call("_ClustAssess_ecs_resident_sim_matrix", ptr, R.int([1, 2]))
    R.reset()
