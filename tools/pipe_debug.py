"""tools/pipe_debug.py [--pageable] -- on the GPU box: ca_consistency on config 5 from host buffers, three calls; CA_PIPE_DEBUG=1 prints the
host-side timeline of the pipelined path (upload groups, packing, wave launches)."""
import sys, time, numpy as np, os
sys.path.insert(0, ".")
import torch
from clustassess_b200 import _capi, synth
lib = _capi.lib(); _capi.init([0])
labels,_ = synth.config("C5"); m,n = labels.shape
host = torch.empty((m,n), dtype=torch.int32, pin_memory=True); host.numpy()[...] = labels; del labels
out = np.empty(n)
import ctypes as C
ptr = np.ctypeslib.as_array(C.cast(host.data_ptr(), C.POINTER(C.c_int32)), shape=(1,))
if "--pageable" in sys.argv:  # ordinary malloc'ed memory, what R's vectors are
    pg = np.array(host.numpy())
    ptr = pg.reshape(-1)
for it in range(int(os.environ.get("CALLS", "3"))):
    t=time.perf_counter(); _capi.check(lib.ca_consistency(ptr, n, m, None, out, None)); print("call %d: %.2f ms" % (it, 1e3*(time.perf_counter()-t)), file=sys.stderr)
