"""tools/multi_debug.py NGPU -- on the GPU box: host-side timeline of the multi-GPU job (CA_PIPE_DEBUG=1 prints the laps of the first and the
last device's thread), C5 resident, three steps."""
import sys, time, numpy as np
sys.path.insert(0, ".")
import torch
from clustassess_b200 import _capi, synth
ng = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
lib = _capi.lib(); _capi.init(list(range(ng)))
labels, _ = synth.config("C5"); m, n = labels.shape
host = torch.empty((m, n), dtype=torch.int32, pin_memory=True); host.numpy()[...] = labels; del labels
dl = _capi.DeviceLabels(n, m).upload_pinned(host.data_ptr())
ecc = np.empty(n)
for it in range(4):
    dl.invalidate()
    t = time.perf_counter()
    _capi.check(lib.ca_consistency_resident(dl._h, None, 0, None, ecc, None))
    print("step %d: %.2f ms" % (it, 1e3 * (time.perf_counter() - t)), [round(_capi.profile(d)["engine_ms"], 2) for d in range(ng)], file=sys.stderr)
