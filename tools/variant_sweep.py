"""tools/variant_sweep.py NAME... -- on the GPU box: load every tuning build tools/libca_NAME.so (tools/build_variant.sh) into ONE
process, run the C5 job (labels generated once) a few times through each and print the stage timings, the split of the wave
kernel's two launches and a checksum of ecc (all variants must agree bit for bit: the sums are fixed-point integers).
usage: python tools/variant_sweep.py [--config C5] [--steps 3] NAME [NAME ...]      ('default' = the in-tree library)"""
import ctypes as C
import hashlib
import os
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch

from clustassess_b200 import _capi, synth

args = sys.argv[1:]
config, steps = "C5", 3
while args and args[0].startswith("--"):
    if args[0] == "--config":
        config = args[1]
    elif args[0] == "--steps":
        steps = int(args[1])
    args = args[2:]
names = args or ["default"]

t0 = time.time()
labels, weights = synth.config(config)
m, n = labels.shape
host = torch.empty((m, n), dtype=torch.int32, pin_memory=True)
host.numpy()[...] = labels
del labels
print("workload %s: %d x %d generated in %.1f s" % (config, m, n, time.time() - t0), flush=True)
w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
units = m * (m - 1) / 2 * n
ecc = torch.zeros(n, dtype=torch.float64, device="cuda")
ref_sum = None
for name in names:
    path = os.path.join("clustassess_b200", "libclustassess_b200.so") if name == "default" else os.path.join("tools", "libca_%s.so" % name)
    if not os.path.exists(path):
        print("%-16s missing (%s)" % (name, path))
        continue
    L = C.CDLL(os.path.abspath(path))
    for sym, (res, argt) in _capi.SYMBOLS.items():
        if hasattr(L, sym):
            fn = getattr(L, sym)
            fn.restype, fn.argtypes = res, argt

    def check(rc):
        if rc != 0:
            raise RuntimeError("%s: %s" % (name, L.ca_last_error().decode()))

    check(L.ca_init(None, 0))
    h = C.c_void_p()
    check(L.ca_labels_create(n, m, C.byref(h)))
    check(L.ca_labels_upload(h, host.data_ptr()))
    prof = np.zeros(10)
    rows = []
    for it in range(steps + 1):
        check(L.ca_labels_invalidate(h))
        ecc.zero_()
        torch.cuda.synchronize()
        t = time.perf_counter()
        check(L.ca_consistency_partial_dev(h, None if w is None else w.ctypes.data, 0, 1, ecc.data_ptr(), None))
        check(L.ca_consistency_finish_dev(n, m, None if w is None else w.ctypes.data, ecc.data_ptr(), None))
        torch.cuda.synchronize()
        wall = 1e3 * (time.perf_counter() - t)
        k = L.ca_get_profile(prof, 10)
        if it > 0:
            rows.append((wall, prof[4], prof[8] if k > 8 else 0.0, prof[9] if k > 8 else 0.0, prof[0] + prof[1], prof[3], prof[2]))
    med = [statistics.median(r[i] for r in rows) for i in range(7)]
    out = ecc.cpu().numpy()
    digest = hashlib.sha1(out.tobytes()).hexdigest()[:12]
    print("%-16s step %7.2f ms  wave %7.2f (streaming %6.2f + resident %7.2f)  pack+transpose %5.2f  sort %5.2f  plan %5.2f  | %.3e units/s  ecc mean %.17g sha1 %s"
          % (name, med[0], med[1], med[2], med[3], med[4], med[5], med[6], units / (med[0] * 1e-3), float(out.mean()), digest), flush=True)
    check(L.ca_labels_destroy(h))
    L.ca_shutdown()
