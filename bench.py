#!/usr/bin/env python
"""bench.py -- headline benchmark of the ECS hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path, host cores

One "step" = one complete element_consistency job over the workload (all M(M-1)/2 partition pairs, all N
cells): label packing, transposition to the compact label layout, per-partition sort, the wave kernel (tables + per-element
ECS), reduction and finish, result copied to the host.  The default workload is BASELINE.json's config 5 (500 partitions x
1M cells, up to 2000 clusters, Louvain-like family), which fits one GPU; with --gpus N > 1 the SAME job
is sharded over N devices (scaling = strong) by ONE process through the library's own multi-device entry: under torchrun
rank 0 drives all N devices and the other ranks exit without work (SURVEY.md section 8b "Threading", 8e).

Printed JSON line: metric = partition-pairs x cells per second.
  value     inputs (the raw int32 M x N label matrix) already resident in HBM (sharded over the devices when N > 1).
  e2e       the same job through the C-ABI entry the R shim binds (ca_consistency) with HOST buffers:
            pinned host -> device copy of the labels and device -> host copy of ecc inside the timed region.
  roofline  the wave kernel (ca_tile.cu): algorithmic bytes (16 B per pair-cell) / its
            CUDA-event duration, against the measured HBM peak of MEASURED_PEAKS.json.
  roofline_onchip  the same rate against the ceilings that actually bound it (issue slots, shared-memory wavefronts).
  cpu_baseline  the reference's own disjointECS (oracle/_ref, else the C port) on 1 host core over a
            bounded sample of the same workload's pairs.
  parity    (N = 1) the product's results on THIS workload against the oracle: ECS-matrix entries of the full job, the
            consistency of a 48-partition subset at every cell, per-pair ECS bit for bit; the bench fails if > 1e-9.
  --mode sim_matrix   element_sim_matrix only (8 B per unit), e.g. --config C4.
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
    ap.add_argument("--mode", default="consistency", choices=["consistency", "sim_matrix"],
                    help="consistency: element_consistency (16 B per unit); sim_matrix: element_sim_matrix only (8 B per unit)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=25.0, help="target CPU work of the cpu_baseline sample")
    return ap.parse_args()


def workload_name(args, m, n):
    what = "element_sim_matrix" if getattr(args, "mode", "consistency") == "sim_matrix" else "element_consistency"
    return "%s %s: %s over %d partitions x %d cells" % (args.config, args.family, what, m, n)


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
def parity_block(args, lib, dl, host, m, n, cores):
    """Parity at the bench's own workload, inside the driver's view (R/ECS.R:1478-1484 per pair).  A subset of `nsub`
    partitions of the workload is taken; the oracle (the reference's disjointECS when oracle/_ref is present) computes, on all
    host cores, every pair among them: mean(ECS) per pair and the sum over the pairs per cell.  The product then runs
      (1) the FULL job with the ECS matrix: its entries for those pairs must equal the oracle's means,
      (2) the consistency of the subset (partitions named by index in the resident set): every cell must equal the oracle's,
      (3) the per-pair entry on a few of the pairs: bit-identical ECS at 1000 fixed cells.
    Returns the dict printed as `parity` (ok = every error <= 1e-9)."""
    import ctypes as C

    import oracle
    from clustassess_b200 import _capi, synth

    impl = "reference" if oracle.have_ref() else "port"
    rng = np.random.default_rng(2024)
    nsub = min(m, 48 if n >= 500_000 else 64)
    parts = sorted(rng.choice(m, size=nsub, replace=False).tolist())
    sub, _ = synth.config(args.config, family=args.family, m=m, n=n, parts=parts)
    pi, pj = np.triu_indices(nsub, k=1)
    cells = np.sort(rng.choice(n, size=min(n, 1000), replace=False)).astype(np.int64)
    t0 = time.perf_counter()
    secs, acc, means, cv = oracle.pairs_detail(sub, pi, pj, cells=cells, nthreads=cores, impl=impl)
    cpu_s = time.perf_counter() - t0
    # (1) the full job with the matrix (this is also what merge_partitions(return_ecs_matrix = TRUE) pays for)
    ecc = np.empty(n, dtype=np.float64)
    sim = np.empty((m, m), dtype=np.float64)
    weights = synth.config_weights(args.config, m)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    dl.invalidate()
    _capi.check(lib.ca_consistency_resident(dl._h, None, 0, _capi._opt(w), ecc, sim.ctypes.data))  # warm-up of the WANT_SIM kernels
    dl.invalidate()
    t0 = time.perf_counter()
    _capi.check(lib.ca_consistency_resident(dl._h, None, 0, _capi._opt(w), ecc, sim.ctypes.data))
    sim_ms = 1e3 * (time.perf_counter() - t0)
    P = np.asarray(parts)
    err_sim = float(np.abs(sim[P[pi], P[pj]] - means).max())
    err_sym = float(np.abs(sim - sim.T).max())
    # (2) the subset's consistency, all cells (unit weights: ecc = sum over the pairs / number of pairs)
    ecc_sub = np.empty(n, dtype=np.float64)
    idx = np.ascontiguousarray(parts, dtype=np.int32)
    _capi.check(lib.ca_consistency_resident(dl._h, idx.ctypes.data, nsub, None, ecc_sub, None))
    err_ecc = float(np.abs(ecc_sub - acc / float(pi.size)).max())
    # (3) the per-pair entry, bit for bit
    exact = True
    out = np.empty(n, dtype=np.float64)
    for t in rng.choice(pi.size, size=min(4, pi.size), replace=False):
        a, b = np.ascontiguousarray(sub[pi[t]]), np.ascontiguousarray(sub[pj[t]])
        _capi.check(lib.ca_ecs_pair(a, b, n, out.ctypes.data, None))
        exact = exact and bool(np.array_equal(out[cells], cv[t]))
    tol = 1e-9
    return {"ok": bool(err_sim <= tol and err_ecc <= tol and exact and err_sym == 0.0), "tolerance": tol, "oracle": impl,
            "partitions_sampled": nsub, "pairs": int(pi.size), "cells_compared": int(n),
            "max_abs_err_sim": err_sim, "max_abs_err_ecc": err_ecc, "ecs_pair_bit_exact_at_1000_cells": exact,
            "sim_symmetric": err_sym == 0.0, "oracle_seconds": cpu_s, "oracle_threads": int(cores),
            "full_job_with_matrix_ms": sim_ms,
            "what": "sim entries of the FULL job vs mean(ECS) of the oracle for all pairs among the sampled partitions; "
                    "consistency of the sampled partitions (resident subset) vs the oracle at every cell; ca_ecs_pair bit for bit"}


def onchip_model():
    """The on-chip ceilings of the wave kernel: counters of this round's ncu capture."""
    p = os.path.join(ROOT, "profiles", "roofline_onchip.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def run_b200(args):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: clustassess_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    ngpu = max(1, int(args.gpus))
    # ONE process drives all the GPUs through the library's own multi-device entry (ca_init with ngpu devices): under
    # torchrun only rank 0 works, the other ranks have nothing to do (SURVEY.md section 8b "Threading", 8e).
    if world > 1 and rank != 0:
        return 0
    if torch.cuda.device_count() < ngpu:
        raise SystemExit("bench.py --gpus %d: only %d device(s) visible" % (ngpu, torch.cuda.device_count()))

    from clustassess_b200 import _capi, synth

    lib = _capi.lib()
    got = _capi.init(list(range(ngpu)))
    assert got == ngpu
    matrix_mode = args.mode == "sim_matrix"
    bytes_per_unit = 8.0 if matrix_mode else BYTES_PER_UNIT

    c = dict(synth.CONFIGS[args.config])
    m, n = args.m or c["m"], args.n or c["n"]
    pairs = m * (m - 1) // 2
    units = float(pairs) * n

    t0 = time.time()
    labels, weights = synth.config(args.config, family=args.family, m=m, n=n)
    gen_s = time.time() - t0
    # pinned host copy of the inputs (what the R shim would hand over), device-resident copy for `value`
    host = torch.empty((m, n), dtype=torch.int32, pin_memory=True)
    host.numpy()[...] = labels
    del labels
    w = None if (weights is None or matrix_mode) else np.ascontiguousarray(weights, dtype=np.float64)
    dl = _capi.DeviceLabels(n, m).upload_pinned(host.data_ptr())  # sharded over the devices when ngpu > 1
    # results land in pinned host memory (the contract's "device -> host read of the step's result")
    ecc_t = torch.empty(n, dtype=torch.float64, pin_memory=True)
    ecc = ecc_t.numpy()
    sim_t = torch.empty((m, m), dtype=torch.float64, pin_memory=True) if matrix_mode else None
    sim = sim_t.numpy() if matrix_mode else None

    def sync_all():
        for d in range(ngpu):
            torch.cuda.synchronize(d)

    stage_keys = ["engine_ms", "streaming_ms", "resident_ms", "pack_ms", "transpose_ms", "sort_ms", "plan_ms", "finish_ms"]
    per_step = []  # per timed step: wall ms, per-device profiles

    def step_resident(record):
        dl.invalidate()  # every step redoes packing / transposition / sort: nothing is cached across steps
        t = time.perf_counter()
        if matrix_mode:
            _capi.check(lib.ca_sim_matrix_resident(dl._h, None, 0, sim))
        else:
            _capi.check(lib.ca_consistency_resident(dl._h, None, 0, _capi._opt(w), ecc, None))
        wall = 1e3 * (time.perf_counter() - t)
        if record:
            per_step.append((wall, [_capi.profile(d) for d in range(ngpu)]))

    def step_e2e(out):
        # exactly the entries the R shim binds: host buffers in, host result out, all selected devices
        if matrix_mode:
            _capi.check(lib.ca_sim_matrix(C_cast(host.data_ptr()), n, m, out))
        else:
            _capi.check(lib.ca_consistency(C_cast(host.data_ptr()), n, m, _capi._opt(w), out, None))

    # ---- value: inputs resident in HBM ----------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident(False)
    sampler = ClockSampler(0)
    sync_all()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_resident(True)
    sync_all()
    ms_total = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    value = units * args.steps / (ms_total * 1e-3)
    walls = sorted(x[0] for x in per_step)
    result = sim if matrix_mode else ecc
    ecc_mean = float(result.mean())
    import hashlib

    digest = hashlib.sha1(result.tobytes()).hexdigest()
    # stage times: per step the MAX over the devices (the devices run side by side), then the mean over the steps
    stage_ms = {k: float(np.mean([max(pr[k] for pr in profs) for _, profs in per_step])) for k in stage_keys}
    n_launch = int(sum(sum(pr["launches"] for pr in profs) for _, profs in per_step)) + args.steps * (2 if matrix_mode else 1)

    # ---- e2e: host buffers in, host result out ----------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        out_t = torch.empty((m, m) if matrix_mode else n, dtype=torch.float64, pin_memory=True)
        out = out_t.numpy()
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e(out)
        ke = max(3, min(args.steps, 5))
        ts = []
        sync_all()
        for _ in range(ke):
            t = time.perf_counter()
            step_e2e(out)
            ts.append(time.perf_counter() - t)
        sync_all()
        e2e = {"value": units * ke / sum(ts), "unit": UNIT, "h2d_bytes_per_step": int(m) * int(n) * 4,
               "d2h_bytes_per_step": int(out.nbytes), "ms_per_step": 1e3 * sum(ts) / ke, "ms_per_step_median": 1e3 * float(np.median(ts)),
               "steps": ke, "entry": ("ca_sim_matrix" if matrix_mode else "ca_consistency") + " (host buffers; %d device(s) selected by ca_init)" % ngpu,
               "result_mean": float(out.mean()), "same_bits_as_resident": bool(np.array_equal(out, result))}
        # the same entry with PAGEABLE host buffers (what R's own vectors are): the library stages the upload through
        # pinned slots on several host threads; CA_UPLOAD_DIRECT=1 is the plain cudaMemcpyAsync, timed beside it
        try:
            pageable = np.array(host.numpy())  # ordinary malloc'ed memory

            def timed(reps):
                step = lambda: _capi.check(lib.ca_sim_matrix(pageable.reshape(-1), n, m, out) if matrix_mode else
                                           lib.ca_consistency(pageable.reshape(-1), n, m, _capi._opt(w), out, None))
                step()
                t = time.perf_counter()
                for _ in range(reps):
                    step()
                return 1e3 * (time.perf_counter() - t) / reps

            staged_ms = timed(2)
            os.environ["CA_UPLOAD_DIRECT"] = "1"
            direct_ms = timed(2)
            del os.environ["CA_UPLOAD_DIRECT"]
            e2e["pageable"] = {"value": units / (staged_ms * 1e-3), "ms_per_step": staged_ms, "direct_memcpy_ms_per_step": direct_ms,
                               "upload": "staged through pinned 4 MB slots by host threads (ca_api.cu upload_segments)"}
            del pageable
        except Exception as ex:  # the headline numbers above do not depend on this leg
            os.environ.pop("CA_UPLOAD_DIRECT", None)
            e2e["pageable"] = {"error": str(ex)}

    # ---- roofline of the dominant kernel (wave_kernel: tables + per-element ECS), measured live ------------------------------------------
    peak, peak_src = measured_peak()
    rb_ms_per_step = stage_ms["engine_ms"]           # CUDA events on the launching stream; max over the devices, each runs 1/ngpu of the pairs
    units_per_launch_set = units / ngpu
    achieved = units_per_launch_set * bytes_per_unit / (rb_ms_per_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("workload") == workload_name(args, m, n) and tj.get("n_gpus", 1) == ngpu:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "wave_kernel (ca_tile.cu), both launches", "kernel_ms_per_step": rb_ms_per_step,
                "kernel_share_of_step": rb_ms_per_step / (ms_total / args.steps),
                "algorithmic_bytes_per_unit": bytes_per_unit, "peak_source": peak_src,
                "note": "frac > 1: the design does not stream (labels are reused out of L2, tables never leave the SM); see roofline_onchip"}
    # the ceilings that actually bound the kernel: issue slots and shared-memory wavefronts per unit (this round's ncu capture)
    onchip = None
    om = onchip_model()
    if om and not matrix_mode:
        sm_mhz = (clocks or {}).get("sm_mhz") or om.get("sm_mhz", 1965.0)
        sms = om.get("sm_count", 148)
        rate = units_per_launch_set / (rb_ms_per_step * 1e-3)
        issue_ceiling = 4.0 * sms * sm_mhz * 1e6 * 32.0 / om["warp_instructions_per_32_units"]
        smem_ceiling = 1.0 * sms * sm_mhz * 1e6 * 32.0 / om["smem_wavefronts_per_32_units"]
        onchip = {"achieved_units_per_s": rate, "issue_slot_ceiling": issue_ceiling, "shared_memory_ceiling": smem_ceiling,
                  "frac_of_issue_ceiling": rate / issue_ceiling, "frac_of_shared_memory_ceiling": rate / smem_ceiling,
                  "bound": "issue slots" if issue_ceiling < smem_ceiling else "shared-memory wavefronts",
                  "frac": rate / min(issue_ceiling, smem_ceiling), "sm_mhz": sm_mhz,
                  "warp_instructions_per_32_units": om["warp_instructions_per_32_units"],
                  "smem_wavefronts_per_32_units": om["smem_wavefronts_per_32_units"], "source": om.get("source")}

    cpu = None
    parity = None
    if not args.no_cpu_baseline and ngpu == 1 and not matrix_mode:
        import oracle

        oracle.build()
        cpu, _, _, _, _ = cpu_baseline(args, m, n, cores=1, seconds=args.cpu_seconds)
        parity = parity_block(args, lib, dl, host, m, n, os.cpu_count() or 1)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpu, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "ms_per_step_median": float(np.median(walls)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, m, n), "mode": args.mode, "partitions": m, "cells": n, "pairs": pairs,
                   "family": args.family, "l2": "inputs (%.2f GB of labels) exceed the 126 MB L2; no flush needed" % (m * n * 4 / 1e9),
                   "sharding": "one process, %d device(s): label matrix sharded by 32-partition spans, pairs of a device's own partitions, "
                               "one ncclAllReduce of int64 sums" % ngpu,
                   "timing": "host clock around K synchronous library calls with device synchronisation on both sides (one process drives "
                             "all devices, so there is no rank barrier); the stage times are CUDA events on each device's stream, max over devices; "
                             "a step ends with the 8 MB result in host memory",
                   "generate_s": gen_s},
        "roofline": roofline, "roofline_onchip": onchip, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": n_launch,
        "clocks": clocks, "stage_ms_per_step": stage_ms, "result_mean": ecc_mean, "result_sha1": digest,
    }
    print(json.dumps(line))
    dl.close()
    if parity is not None and not parity["ok"]:
        print("bench.py: PARITY FAILED: %s" % json.dumps(parity), file=sys.stderr)
        return 1
    return 0


import ctypes as _C  # noqa: E402

C_int = _C.c_int


def C_cast(ptr):
    """int address -> something the ndpointer-typed prototype accepts (a zero-copy int32 view)."""
    return np.ctypeslib.as_array(_C.cast(ptr, _C.POINTER(_C.c_int32)), shape=(1,))


if __name__ == "__main__":
    a = parse_args()
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
