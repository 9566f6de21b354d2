"""SURVEY.md section 8d's small measurements: C1 and C2 as LATENCY (host buffers in, host result out, through the Python
mirror of the R API), and the identical-partition pre-pass (merge_identical_partitions, R/ECS.R:1021-1054) timed
separately, each with the reference's CPU path (oracle/_ref, else the C port) on one host core beside it.
usage: python tools/run_c1_c2_dedup.py [--cpu-only]      (round_measure.sh runs it on the GPU box)"""
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import oracle
from clustassess_b200 import synth

cpu_only = "--cpu-only" in sys.argv
if not cpu_only:
    from clustassess_b200 import ecs


def med_us(f, reps, warm=3):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t)
    return 1e6 * statistics.median(ts)


pair = oracle.ref_disjoint_ecs if oracle.have_ref() else oracle.disjoint_ecs
kind = "reference build (oracle/_ref)" if oracle.have_ref() else "C port (oracle)"

# ---- C2: one pair, 100k cells, 20 vs 35 clusters ---------------------------------------------------------------------
(a, b), _ = synth.config("C2")
cpu = med_us(lambda: pair(a, b), 20)
line = "C2 element_sim_elscore (100k cells, 20 vs 35 clusters): CPU %s %.0f us" % (kind, cpu)
if not cpu_only:
    gpu = med_us(lambda: ecs.element_sim_elscore(a, b), 50)
    gpu_avg = med_us(lambda: ecs.element_sim(a, b), 50)
    assert np.array_equal(ecs.element_sim_elscore(a, b), oracle.disjoint_ecs(a, b))
    line += "; B200 %.0f us per call (element_sim: %.0f us), upload of 0.8 MB and download of 0.8 MB included, bit-identical" % (gpu, gpu_avg)
print(line)

# ---- C1: 30 partitions x 2700 cells -------------------------------------------------------------------------------------
L1, _ = synth.config("C1")
cpu = med_us(lambda: oracle.element_consistency(L1), 5, warm=1)
line = "C1 element_consistency (30 x 2700): CPU oracle (C port of the R loops) %.0f us" % cpu
if not cpu_only:
    cols = list(L1)
    gpu = med_us(lambda: ecs.element_consistency(cols), 30)
    err = np.abs(ecs.element_consistency(cols) - oracle.element_consistency(L1)).max()
    line += "; B200 %.0f us per call (dedup + consistency), max |err| %.2g" % (gpu, err)
print(line)

# ---- the dedup pre-pass on its own: C3 shape with every partition present twice ----------------------------------------
L3, _ = synth.config("C3")
dup = np.concatenate([L3, L3[::-1]])  # 200 partitions, 100 distinct
t = time.perf_counter()
keep_ref, freq_ref = oracle.merge_identical_partitions(dup)
cpu_s = time.perf_counter() - t
line = "merge_identical_partitions (200 x 100k, 100 distinct): CPU port of R/ECS.R:1021-1054 %.3f s" % cpu_s
if not cpu_only:
    recs = [{"mb": r, "freq": 1} for r in dup]
    gpu = med_us(lambda: ecs.merge_identical_partitions(recs), 5, warm=1)
    got = ecs.merge_identical_partitions(recs)
    same = len(got) == len(keep_ref) and all(np.array_equal(g["mb"], dup[k]) and g["freq"] == f for g, k, f in zip(got, keep_ref, freq_ref))
    line += "; B200 %.4f s per call (upload + canonical labelling + signatures + exact confirmation), same partitions and freq: %s" % (gpu / 1e6, same)
print(line)

# ---- the dedup pre-pass at the headline shape (GPU only: the reference's M x U contingency tables at 500 x 1M are core-hours) ----
if not cpu_only:
    from clustassess_b200 import _capi

    L5, _ = synth.config("C5")
    m5, n5 = L5.shape
    cols5 = [np.ascontiguousarray(r) for r in L5]
    t = time.perf_counter()
    dev = _capi.DeviceLabels(n5, m5, primary=True).upload_columns(cols5)
    up_s = time.perf_counter() - t
    ts = []
    for _ in range(4):
        dev.invalidate()
        t = time.perf_counter()
        keep, freq = dev.merge_identical(None)
        ts.append(time.perf_counter() - t)
    print("merge_identical_partitions on C5 (500 x 1M, all distinct): upload of 2 GB from pageable memory %.3f s (once: the consistency that "
          "follows uses the same resident set), dedup pre-pass %.4f s per call (min/max + ranks + canonical labelling + signatures), kept %d"
          % (up_s, statistics.median(ts[1:]), len(keep)))
    t = time.perf_counter()
    ecc = ecs.element_consistency(cols5)
    t1 = time.perf_counter() - t
    t = time.perf_counter()
    ecc = ecs.element_consistency(cols5)
    t2 = time.perf_counter() - t
    print("element_consistency(C5) through the Python mirror of the R API, host list in, host vector out (one upload, dedup, consistency): "
          "%.3f s, then %.3f s; mean ecc %.12f" % (t1, t2, float(np.mean(ecc))))
    dev.close()
