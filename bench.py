#!/usr/bin/env python
"""bench.py -- headline benchmark of the ECS hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path, host cores

One "step" = one complete element_consistency job over the workload (all M(M-1)/2 partition pairs, all N
cells): label packing, cell-major transposition, per-partition sort, the wave kernel (tables + per-element ECS), all-reduce
(N > 1) and finish.  The default workload is BASELINE.json's config 5 (500 partitions x 1M cells, up to
2000 clusters, Louvain-like family), which fits one GPU; at N > 1 the SAME job is
pair-sharded over the ranks (scaling = strong).

Printed JSON line (rank 0): metric = partition-pairs x cells per second.
  value     inputs (the raw int32 M x N label matrix) already resident in HBM; CUDA-event timed.
  e2e       the same job through the C-ABI entry the R shim binds (ca_consistency) with HOST buffers:
            pinned host -> device copy of the labels and device -> host copy of ecc inside the timed region.
  roofline  the wave kernel (ca_tile.cu): algorithmic bytes (16 B per pair-cell) / its
            CUDA-event duration, against the measured HBM peak of MEASURED_PEAKS.json.
  cpu_baseline  the reference's own disjointECS (oracle/_ref, else the C port) on 1 host core over a
            bounded sample of the same workload's pairs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "partition_pairs_x_cells_per_sec"
UNIT = "pair*cells/s"
BYTES_PER_UNIT = 16.0  # two int32 labels x two passes


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C5", help="workload: C3, C4 or C5 (default: the metric's config)")
    ap.add_argument("--family", default="louvain", choices=["louvain", "uniform"])
    ap.add_argument("--m", type=int, default=None, help="override the number of partitions")
    ap.add_argument("--n", type=int, default=None, help="override the number of cells")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=25.0, help="target CPU work of the cpu_baseline sample")
    return ap.parse_args()


def workload_name(args, m, n):
    return "%s %s: element_consistency over %d partitions x %d cells" % (args.config, args.family, m, n)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
def cpu_baseline(args, m, n, cores, seconds):
    """Time the reference's per-pair body (disjointECS + mean + accumulate, R/ECS.R:1478-1484) on `cores`
    host threads over a bounded random sample of the workload's pairs."""
    import oracle
    from clustassess_b200 import synth

    impl = "reference" if oracle.have_ref() else "port"
    rng = np.random.default_rng(12345)
    # probe one pair to size the sample
    parts = sorted(rng.choice(m, size=min(m, 48), replace=False).tolist())
    labels, _ = synth.config(args.config, family=args.family, m=m, n=n, parts=parts)
    t_probe = oracle.time_pairs(labels, [0], [1], nthreads=1, impl=impl)
    npairs = int(max(cores, min(len(parts) * (len(parts) - 1) // 2, seconds * cores / max(t_probe, 1e-6))))
    allp = [(i, j) for i in range(len(parts)) for j in range(i + 1, len(parts))]
    pick = rng.choice(len(allp), size=min(npairs, len(allp)), replace=False)
    pi = np.array([allp[k][0] for k in pick], dtype=np.int32)
    pj = np.array([allp[k][1] for k in pick], dtype=np.int32)
    secs = oracle.time_pairs(labels, pi, pj, nthreads=cores, impl=impl)
    return {"value": float(pi.size) * n / secs, "unit": UNIT, "cores": int(cores), "kind": impl,
            "sample": "%d random pairs among %d of the workload's %d partitions, %d cells each, %.1f s" %
                      (pi.size, len(parts), m, n, secs), "seconds": secs, "pairs": int(pi.size)}, labels, pi, pj, impl


def run_reference(args):
    """--impl reference: the reference's CPU implementation on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    from clustassess_b200 import synth

    oracle.build()
    c = dict(synth.CONFIGS[args.config])
    m, n = args.m or c["m"], args.n or c["n"]
    cores = os.cpu_count() or 1
    base, labels, pi, pj, impl = cpu_baseline(args, m, n, cores, seconds=6.0)
    times = []
    for it in range(args.warmup + args.steps):
        secs = oracle.time_pairs(labels, pi, pj, nthreads=cores, impl=impl)
        if it >= args.warmup:
            times.append(secs)
    tot = sum(times)
    value = float(pi.size) * n * len(times) / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, m, n), "step": "bounded sample: %d pairs x %d cells per step" % (pi.size, n)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": impl,
                         "sample": "%d random pairs x %d cells per step, %d host threads" % (pi.size, n, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
def load_workload(args, rank, world, m, n):
    """Rank 0 generates the label matrix (seeded numpy), the other ranks read it from /dev/shm."""
    from clustassess_b200 import synth

    path = "/dev/shm/ca_bench_%s_%s_%d_%d.npy" % (args.config, args.family, m, n)
    if world == 1:
        return synth.config(args.config, family=args.family, m=m, n=n)
    import torch.distributed as dist

    weights = synth.config_weights(args.config, m)
    if rank == 0:
        labels, _ = synth.config(args.config, family=args.family, m=m, n=n)
        np.save(path, labels)
    dist.barrier()
    if rank != 0:
        labels = np.load(path, mmap_mode="r")
    dist.barrier()
    if rank == 0:
        try:
            os.remove(path)
        except OSError:
            pass
    return labels, weights


def run_b200(args):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: clustassess_b200 has no CPU fallback")
    import torch.distributed as dist

    from clustassess_b200 import _capi, synth
    from clustassess_b200.distributed import consistency_sharded, upload_sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _capi.lib()
    dev = (C_int * 1)(local)
    _capi.check(lib.ca_init(dev, 1))

    c = dict(synth.CONFIGS[args.config])
    m, n = args.m or c["m"], args.n or c["n"]
    pairs = m * (m - 1) // 2
    units = float(pairs) * n

    t0 = time.time()
    labels, weights = load_workload(args, rank, world, m, n)
    gen_s = time.time() - t0
    # pinned host copy of the inputs (what the R shim would hand over), device-resident copy for `value`
    host = torch.empty((m, n), dtype=torch.int32, pin_memory=True)
    host.numpy()[...] = labels
    del labels
    if world == 1:
        dl = _capi.DeviceLabels(n, m).upload_pinned(host.data_ptr())
    else:  # the label matrix lives in a torch tensor so that NCCL can all-gather the ranks' upload shards into it
        dev_labels = torch.empty((m, n), dtype=torch.int32, device="cuda")
        upload_sharded(host, dev_labels)
        torch.cuda.synchronize()
        dl = _capi.DeviceLabels(n, m, device_ptr=dev_labels.data_ptr())
    stream = torch.cuda.current_stream()
    _capi.check(lib.ca_set_stream(stream.cuda_stream))

    prof_acc = {"engine_ms": 0.0, "pack_ms": 0.0, "transpose_ms": 0.0, "sort_ms": 0.0, "plan_ms": 0.0, "finish_ms": 0.0}
    launches = [0]

    def partial(r, w, ecc, sim):
        dl.invalidate()  # every step redoes packing / transposition / sort: nothing is cached across steps
        dl.consistency_partial(weights, r, w, ecc.data_ptr(), None if sim is None else sim.data_ptr())
        p = _capi.profile()
        for k in prof_acc:
            if k != "finish_ms":
                prof_acc[k] += p[k]
        launches[0] += int(p["launches"])

    def finish(ecc, sim):
        _capi.consistency_finish(n, m, weights, ecc.data_ptr(), None if sim is None else sim.data_ptr())
        launches[0] += 1

    def step_resident():
        ecc, _ = consistency_sharded(n, m, weights, partial, finish, device="cuda", fixed_point=True)
        return ecc

    def step_e2e():
        if world == 1:  # exactly the entry point the R shim binds
            out = np.empty(n, dtype=np.float64)
            w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
            _capi.check(lib.ca_consistency(C_cast(host.data_ptr()), n, m, _capi._opt(w), out, None))
            return out
        upload_sharded(host, dev_labels)   # 1/world of the matrix over PCIe per rank + one NCCL all-gather
        torch.cuda.current_stream().synchronize()
        ecc, _ = consistency_sharded(n, m, weights, partial, finish, device="cuda", fixed_point=True)
        return ecc.cpu().numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ----------------------------------------------------------------------
    for _ in range(args.warmup):
        ecc = step_resident()
    for k in prof_acc:
        prof_acc[k] = 0.0
    launches[0] = 0
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        ecc = step_resident()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    rb = torch.tensor([prof_acc["engine_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rb, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = units * args.steps / (ms_total * 1e-3)
    ecc_mean = float(ecc.mean().item())
    n_launch = launches[0]
    stage_ms = {k: v / args.steps for k, v in prof_acc.items()}  # of the timed resident steps only

    # ---- e2e: host buffers in, host result out ----------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        _capi.check(lib.ca_set_stream(None))
        for _ in range(max(1, min(args.warmup, 2))):
            out = step_e2e()
        ke = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            out = step_e2e()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": units * ke / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(m) * int(n) * 4,
               "d2h_bytes_per_step": int(n) * 8, "ms_per_step": 1e3 * float(dt.item()) / ke, "steps": ke,
               "entry": "ca_consistency (host buffers)" if world == 1 else "sharded upload + all_gather + ca_consistency_partial_dev + all_reduce + finish + D2H",
               "ecc_mean": float(np.mean(out))}

    # ---- the same entry with PAGEABLE host buffers (what R's own vectors are): the library stages the upload through
    # pinned slots on several host threads; CA_UPLOAD_DIRECT=1 is the plain cudaMemcpyAsync, timed beside it ---------
    if e2e is not None and world == 1:
        try:
            pageable = np.array(host.numpy())  # ordinary malloc'ed memory
            w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
            out = np.empty(n, dtype=np.float64)

            def timed(reps):
                _capi.check(lib.ca_consistency(pageable.reshape(-1), n, m, _capi._opt(w), out, None))  # warm-up
                t0 = time.perf_counter()
                for _ in range(reps):
                    _capi.check(lib.ca_consistency(pageable.reshape(-1), n, m, _capi._opt(w), out, None))
                return 1e3 * (time.perf_counter() - t0) / reps

            staged_ms = timed(2)
            os.environ["CA_UPLOAD_DIRECT"] = "1"
            direct_ms = timed(2)
            del os.environ["CA_UPLOAD_DIRECT"]
            e2e["pageable"] = {"value": units / (staged_ms * 1e-3), "ms_per_step": staged_ms, "direct_memcpy_ms_per_step": direct_ms,
                               "upload": "staged: 8 host threads x 2 pinned 4 MB slots (ca_api.cu upload_segments)",
                               "ecc_mean": float(np.mean(out))}
            del pageable
        except Exception as ex:  # the headline numbers above do not depend on this leg
            os.environ.pop("CA_UPLOAD_DIRECT", None)
            e2e["pageable"] = {"error": str(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (wave_kernel: tables + per-element ECS), measured live ------------------------------------------
    peak, peak_src = measured_peak()
    rb_ms_per_step = float(rb.item()) / args.steps           # max over ranks; each rank runs 1/world of the pairs
    units_per_launch_set = units / world
    achieved = units_per_launch_set * BYTES_PER_UNIT / (rb_ms_per_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("workload") == workload_name(args, m, n) and tj.get("n_gpus", 1) == world:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "wave_kernel (ca_tile.cu)", "kernel_ms_per_step": rb_ms_per_step, "kernel_share_of_step": rb_ms_per_step / (ms_total / args.steps),
                "algorithmic_bytes_per_unit": BYTES_PER_UNIT, "peak_source": peak_src}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        import oracle

        oracle.build()
        cpu, _, _, _, _ = cpu_baseline(args, m, n, cores=1, seconds=args.cpu_seconds)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, m, n), "partitions": m, "cells": n, "pairs": pairs,
                   "family": args.family, "l2": "inputs (%.2f GB of labels) exceed the 126 MB L2; no flush needed" % (m * n * 4 / 1e9),
                   "sharding": "pairs over %d rank(s), zigzag by first partition" % world, "generate_s": gen_s},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": n_launch, "clocks": clocks,
        "stage_ms_per_step": stage_ms, "ecc_mean": ecc_mean,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


import ctypes as _C  # noqa: E402

C_int = _C.c_int


def C_cast(ptr):
    """int address -> something the ndpointer-typed prototype accepts (a zero-copy int32 view)."""
    return np.ctypeslib.as_array(_C.cast(ptr, _C.POINTER(_C.c_int32)), shape=(1,))


if __name__ == "__main__":
    a = parse_args()
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
