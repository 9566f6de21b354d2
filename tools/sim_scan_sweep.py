"""tools/sim_scan_sweep.py -- on the GPU box: the ECS matrix summed over the TABLE entries (scan_sim, ca_tile.cu) against the per-cell
way, on resident label sets: job time of (ecc + matrix) and (matrix only) with CA_SIM_SCAN=0 / 1, and what the library's own choice is.
Shapes: the configs plus a sweep of the cluster count at fixed n (where the cost model in run_allpairs puts the switch)."""
import os
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from clustassess_b200 import _capi, synth


def med_ms(f, reps=5, warm=2):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t)
    return 1e3 * statistics.median(ts)


def run(name, L):
    m, n = L.shape
    dev = _capi.DeviceLabels(n, m).upload(np.ascontiguousarray(L))
    kmax = int(max(len(np.unique(r)) for r in L[: min(m, 8)]))
    out = {}
    os.environ["CA_RT"] = "0"  # first the two routes to the matrix without ratio tables
    for mode, f in (("ecc+matrix", lambda: dev.consistency(None, None, True)), ("matrix only", lambda: dev.sim_matrix()), ("ecc only", lambda: dev.consistency())):
        for scan in ("0", "1", None):
            if scan is None:
                os.environ.pop("CA_SIM_SCAN", None)
            else:
                os.environ["CA_SIM_SCAN"] = scan
            if mode == "ecc only" and scan is not None:
                continue
            out[(mode, scan)] = med_ms(f)
    rt = {}
    for mode, f in (("ecc+matrix", lambda: dev.consistency(None, None, True)), ("ecc only", lambda: dev.consistency())):
        for v in ("0", "1", None):
            if v is None:
                os.environ.pop("CA_RT", None)
            else:
                os.environ["CA_RT"] = v
            rt[(mode, v)] = med_ms(f)
    os.environ["CA_RT"] = "0"
    e0, s0 = dev.consistency(None, None, True)
    os.environ["CA_RT"] = "1"
    e1, s1 = dev.consistency(None, None, True)
    os.environ.pop("CA_RT", None)
    print("%-34s ratio tables: ecc only off %8.3f on %8.3f chosen %8.3f | ecc+matrix off %8.3f on %8.3f chosen %8.3f ms | max |on - off| ecc %.2g matrix %.2g"
          % (name, rt[("ecc only", "0")], rt[("ecc only", "1")], rt[("ecc only", None)], rt[("ecc+matrix", "0")], rt[("ecc+matrix", "1")],
             rt[("ecc+matrix", None)], float(np.abs(e0 - e1).max()), float(np.abs(s0 - s1).max())), flush=True)
    os.environ["CA_SIM_SCAN"] = "0"
    a = dev.sim_matrix()
    os.environ["CA_SIM_SCAN"] = "1"
    b = dev.sim_matrix()
    os.environ.pop("CA_SIM_SCAN", None)
    units = m * (m - 1) / 2 * n
    print("%-34s K~%-5d ecc only %8.3f | ecc+matrix: cells %8.3f tables %8.3f chosen %8.3f | matrix only: cells %8.3f tables %8.3f chosen %8.3f ms"
          " | %.3g units/s chosen (matrix only) | max |tables - cells| %.2g"
          % (name, kmax, out[("ecc only", None)], out[("ecc+matrix", "0")], out[("ecc+matrix", "1")], out[("ecc+matrix", None)],
             out[("matrix only", "0")], out[("matrix only", "1")], out[("matrix only", None)], units / out[("matrix only", None)] * 1e3,
             float(np.abs(np.asarray(a) - np.asarray(b)).max())), flush=True)
    dev.close()


which = sys.argv[1:] or ["configs", "ksweep"]
if "configs" in which:
    for cfg, kw in (("C1", {}), ("C3", {}), ("C4", {}), ("C5", {"m": 96})):
        L, _ = synth.config(cfg, **kw)
        L = np.asarray(L) if not isinstance(L, tuple) else np.stack(L)
        run("%s %s" % (cfg, "x".join(map(str, L.shape))), L)
if "ksweep" in which:
    for n in (50000, 1000000):
        for k in (8, 32, 64, 128, 256, 512, 1024):
            if k * 8 > n:
                continue
            m = 128 if n <= 100000 else 64
            L = synth.louvain_like(m, n, k, max(2, k * 3 // 4), k * 5 // 4, seed=7)
            run("louvain %dx%d K0=%d" % (m, n, k), np.asarray(L))
