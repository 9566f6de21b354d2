"""ctypes binding of the C ABI declared in include/clustassess_b200.h.

This is the Python twin of the R shim (r/clustassess_shim.c): it only marshals numpy buffers into the
`extern "C"` entry points.  There is no fallback: if the shared library is missing, or no B200 is
visible, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLUSTASSESS_B200_LIB") or os.path.join(_HERE, "libclustassess_b200.so")

CA_OK = 0
CA_E_ARG = -3

_I32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_F64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

# every symbol the header declares: name -> (restype, argtypes)
SYMBOLS = {
    "ca_init": (C.c_int, [C.c_void_p, C.c_int]),
    "ca_shutdown": (None, []),
    "ca_last_error": (C.c_char_p, []),
    "ca_version": (C.c_int, []),
    "ca_device_count": (C.c_int, []),
    "ca_set_stream": (C.c_int, [C.c_void_p]),
    "ca_use_stream": (C.c_int, [C.c_void_p, C.c_int]),
    "ca_label_range": (C.c_int, [_I32, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ca_contingency": (C.c_int, [_I32, _I32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I32, C.c_void_p, C.c_void_p]),
    "ca_ecs_pair": (C.c_int, [_I32, _I32, C.c_int64, C.c_void_p, C.c_void_p]),
    "ca_ecs_pair_mean": (C.c_int, [_I32, _I32, C.c_int64, C.POINTER(C.c_double)]),
    "ca_consistency": (C.c_int, [_I32, C.c_int64, C.c_int32, C.c_void_p, _F64, C.c_void_p]),
    "ca_consistency_cols": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, _F64, C.c_void_p]),
    "ca_sim_matrix": (C.c_int, [_I32, C.c_int64, C.c_int32, _F64]),
    "ca_sim_matrix_cols": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _F64]),
    "ca_agreement": (C.c_int, [_I32, _I32, C.c_int64, C.c_int32, _F64]),
    "ca_agreement_cols": (C.c_int, [_I32, C.c_void_p, C.c_int64, C.c_int32, _F64]),
    "ca_merge_identical": (C.c_int, [_I32, C.c_int64, C.c_int32, C.c_void_p, _I32, _F64, C.POINTER(C.c_int32)]),
    "ca_merge_identical_cols": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, _I32, _F64, C.POINTER(C.c_int32)]),
    "ca_labels_create": (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]),
    "ca_labels_create_on_primary": (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]),
    "ca_labels_shape": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "ca_labels_is_sharded": (C.c_int, [C.c_void_p]),
    "ca_labels_wrap": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ca_labels_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ca_labels_upload_col": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "ca_labels_upload_cols": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ca_labels_device_ptr": (C.c_void_p, [C.c_void_p]),
    "ca_labels_invalidate": (C.c_int, [C.c_void_p]),
    "ca_labels_destroy": (C.c_int, [C.c_void_p]),
    "ca_labels_select": (C.c_int, [C.c_void_p, _I32, C.c_int32, C.POINTER(C.c_void_p)]),
    "ca_consistency_resident": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, _F64, C.c_void_p]),
    "ca_sim_matrix_resident": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _F64]),
    "ca_merge_identical_resident": (C.c_int, [C.c_void_p, C.c_void_p, _I32, _F64, C.POINTER(C.c_int32)]),
    "ca_shard_owner": (C.c_int, [C.c_int32, C.c_int32]),
    "ca_consistency_partial_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "ca_consistency_finish_dev": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ca_get_profile": (C.c_int, [_F64, C.c_int]),
    "ca_get_profile_dev": (C.c_int, [C.c_int, _F64, C.c_int]),
    "ca_plan_spans": (C.c_int, [C.c_int32, C.c_int32, _I32, C.c_int32]),
    "ca_plan_route": (C.c_int, [C.c_int32, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ca_plan_blocks": (C.c_int, [_I32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _I32, _I32, _I32, C.c_int32,
                                 C.POINTER(C.c_int64)]),
}

_lib = None


class ClustAssessError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("clustassess_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    """The loaded shared library (raises if it has not been built: there is no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "%s is missing: run `python __graft_entry__.py` (nvcc, sm_100a) first. "
                "clustassess_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != CA_OK:
        raise ClustAssessError(rc, lib().ca_last_error().decode("utf-8", "replace"))


def _opt(a):
    return None if a is None else a.ctypes.data


_PROFILE_KEYS = ["pack_ms", "transpose_ms", "plan_ms", "sort_ms", "engine_ms", "finish_ms", "call_ms", "launches", "streaming_ms", "resident_ms"]


def profile(device_index=None):
    """Stage timings of the last batched call: of the primary device, or of the index-th selected device of a multi-GPU job."""
    out = np.zeros(10, dtype=np.float64)
    if device_index is None:
        lib().ca_get_profile(out, 10)
    else:
        lib().ca_get_profile_dev(int(device_index), out, 10)
    return dict(zip(_PROFILE_KEYS, out.tolist()))


def init(devices=None):
    """ca_init: select the device(s) this process computes on (None = device 0)."""
    if devices is None:
        check(lib().ca_init(None, 0))
    else:
        arr = (C.c_int * len(devices))(*[int(d) for d in devices])
        check(lib().ca_init(arr, len(devices)))
    return lib().ca_device_count()


class DeviceLabels:
    """A device-resident M x N int32 label matrix (ca_labels_t)."""

    def __init__(self, n, m, device_ptr=None, primary=False):
        """device_ptr: wrap caller-owned device memory (M x N int32) instead of allocating.  primary: a single-device set
        even when several devices are selected (what the dedup pre-pass and the rank-sharded entries need)."""
        self.n, self.m = int(n), int(m)
        h = C.c_void_p()
        if device_ptr is None:
            fn = lib().ca_labels_create_on_primary if primary else lib().ca_labels_create
            check(fn(self.n, self.m, C.byref(h)))
        else:
            check(lib().ca_labels_wrap(self.n, self.m, C.c_void_p(device_ptr), C.byref(h)))
        self._h = h

    def upload(self, labels):
        a = np.ascontiguousarray(labels, dtype=np.int32)
        assert a.shape == (self.m, self.n)
        check(lib().ca_labels_upload(self._h, a.ctypes.data))
        return self

    def upload_columns(self, cols):
        """M separately allocated int32 vectors (R's list of vectors) in one staged upload, no host-side M x N copy."""
        assert len(cols) == self.m and all(c.dtype == np.int32 and c.size == self.n and c.flags.c_contiguous for c in cols)
        ptrs = (C.c_void_p * self.m)(*[c.ctypes.data for c in cols])
        check(lib().ca_labels_upload_cols(self._h, ptrs))
        return self

    def upload_pinned(self, ptr):
        check(lib().ca_labels_upload(self._h, ptr))
        return self

    @property
    def device_ptr(self):
        return lib().ca_labels_device_ptr(self._h)

    def invalidate(self):
        check(lib().ca_labels_invalidate(self._h))

    @property
    def sharded(self):
        return bool(lib().ca_labels_is_sharded(self._h))

    def merge_identical(self, freq=None):
        """merge_identical_partitions (R/ECS.R:1021-1054) of this resident (single-device) set -> (kept indices, summed freq)."""
        f = None if freq is None else np.ascontiguousarray(freq, dtype=np.float64)
        keep = np.zeros(self.m, dtype=np.int32)
        fout = np.zeros(self.m, dtype=np.float64)
        u = C.c_int32()
        check(lib().ca_merge_identical_resident(self._h, _opt(f), keep, fout, C.byref(u)))
        return keep[:u.value].copy(), fout[:u.value].copy()

    def select(self, idx):
        """A new resident set holding the partitions idx (device-to-device copies)."""
        ix = np.ascontiguousarray(idx, dtype=np.int32)
        h = C.c_void_p()
        check(lib().ca_labels_select(self._h, ix, ix.size, C.byref(h)))
        out = DeviceLabels.__new__(DeviceLabels)
        out.n, out.m, out._h = self.n, int(ix.size), h
        return out

    def consistency(self, idx=None, weights=None, want_matrix=False):
        """weighted_element_consistency of the resident partitions (all, or those named by idx) -> (ecc, matrix or None)."""
        ix = None if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
        k = self.m if ix is None else int(ix.size)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        if w is not None and w.size != k:
            raise ValueError("weights must have one entry per selected partition")
        ecc = np.empty(self.n, dtype=np.float64)
        sim = np.empty((k, k), dtype=np.float64) if want_matrix else None
        check(lib().ca_consistency_resident(self._h, _opt(ix), k, _opt(w), ecc, _opt(sim)))
        return ecc, sim

    def sim_matrix(self, idx=None):
        ix = None if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
        k = self.m if ix is None else int(ix.size)
        sim = np.empty((k, k), dtype=np.float64)
        check(lib().ca_sim_matrix_resident(self._h, _opt(ix), k, sim))
        return sim

    def consistency_partial(self, weights, rank, world, ecc_ptr, sim_ptr):
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        check(lib().ca_consistency_partial_dev(self._h, _opt(w), rank, world, ecc_ptr, sim_ptr))

    def close(self):
        if self._h is not None and self._h.value:
            lib().ca_labels_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def consistency_finish(n, m, weights, ecc_ptr, sim_ptr):
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    check(lib().ca_consistency_finish_dev(int(n), int(m), _opt(w), ecc_ptr, sim_ptr))
