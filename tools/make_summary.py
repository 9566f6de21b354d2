#!/usr/bin/env python
"""tools/make_summary.py TAG1 [MULTI_TAG] -- profiles/r2_summary.txt: the round's bench lines in readable form.
TAG1: tools/round_measure.sh run on one GPU (gpurun_out/TAG1_*); MULTI_TAG: the --gpus 8 call that ran bench.py at 8 / 4 / 2 devices."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
tag1 = sys.argv[1]
tagm = sys.argv[2] if len(sys.argv) > 2 else None


def last_json(path):
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def line(label, d):
    e = d.get("e2e") or {}
    st = {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}
    return ("%-34s value %.4g pair*cells/s  %.2f ms/step (median %.2f)  e2e %.4g (%.2f ms)  wave %.2f ms = %.2f x HBM roofline  launches %d  sha1 %s\n%34s stages %s" %
            (label, d["value"], d["ms_per_step"], d.get("ms_per_step_median", float("nan")), e.get("value", float("nan")), e.get("ms_per_step", float("nan")),
             d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d.get("result_sha1", "")[:12], "", st))


out = ["# Round 2, B200: bench.py lines (gpurun_out/%s_*%s).  value = labels resident in HBM, result to pinned host memory; e2e = host buffers in," % (tag1, (", gpurun_out/%s_*" % tagm) if tagm else ""),
       "# host result out through ca_consistency / ca_sim_matrix.  One process drives all devices (ca_init).", ""]
d1 = last_json(os.path.join(G, tag1 + "_bench_n1.json"))
if d1:
    out.append(line("C5 louvain, 1 GPU", d1))
    oc = d1.get("roofline_onchip") or {}
    if oc:
        out.append("%34s on-chip: %.3g units/s = %.1f %% of the issue-slot ceiling (%.3g), %.1f %% of the shared-memory ceiling (%.3g); %.1f warp-instr and %.2f smem wavefronts per 32 units" %
                   ("", oc["achieved_units_per_s"], 100 * oc["frac_of_issue_ceiling"], oc["issue_slot_ceiling"], 100 * oc["frac_of_shared_memory_ceiling"], oc["shared_memory_ceiling"],
                    oc["warp_instructions_per_32_units"], oc["smem_wavefronts_per_32_units"]))
    if d1.get("parity"):
        p = d1["parity"]
        out.append("%34s parity vs %s: ok=%s  %d pairs of the FULL job's matrix max|err| %.3g; consistency of %d sampled partitions at %d cells max|err| %.3g; ca_ecs_pair bit-exact %s; full job WITH the matrix %.1f ms" %
                   ("", p["oracle"], p["ok"], p["pairs"], p["max_abs_err_sim"], p["partitions_sampled"], p["cells_compared"], p["max_abs_err_ecc"], p["ecs_pair_bit_exact_at_1000_cells"], p["full_job_with_matrix_ms"]))
    if d1.get("cpu_baseline"):
        c = d1["cpu_baseline"]
        out.append("%34s cpu_baseline (%s, %d core): %.4g pair*cells/s  (%s)" % ("", c["kind"], c["cores"], c["value"], c["sample"]))
    pg = (d1.get("e2e") or {}).get("pageable")
    if pg and "ms_per_step" in pg:
        out.append("%34s e2e from PAGEABLE host memory: %.2f ms (library-staged upload); plain cudaMemcpyAsync of the same buffers: %.2f ms" % ("", pg["ms_per_step"], pg["direct_memcpy_ms_per_step"]))
ref = last_json(os.path.join(G, tag1 + "_reference_arm.json"))
if ref:
    out.append("%-34s value %.4g pair*cells/s on %d host threads (%s)" % ("reference arm (--impl reference)", ref["value"], ref["cpu_baseline"]["cores"], ref["cpu_baseline"]["sample"]))
base = d1["ms_per_step"] if d1 else None
base_e = (d1.get("e2e") or {}).get("ms_per_step") if d1 else None
if tagm:
    out.append("")
    for n in (2, 4, 8):
        d = last_json(os.path.join(G, "%s_bench_n%d.json" % (tagm, n)))
        if d:
            out.append(line("C5 louvain, %d GPUs" % n, d))
            if base:
                out.append("%34s strong-scaling efficiency: value %.3f, e2e %.3f" % ("", base / (n * d["ms_per_step"]), (base_e / (n * d["e2e"]["ms_per_step"])) if base_e and d.get("e2e") else float("nan")))
    for name, label in (("_bench_c4_sim_n8.json", "C4 element_sim_matrix, 8 GPUs"), ("_bench_c3_n8.json", "C3 weighted, 8 GPUs")):
        d = last_json(os.path.join(G, tagm + name))
        if d:
            out.append(line(label, d))
out.append("")
for name, label in (("_bench_uniform.json", "C5 uniform labels, 1 GPU"), ("_bench_c4shape.json", "C4 shape, consistency, 1 GPU"),
                    ("_bench_c4_sim_matrix.json", "C4 element_sim_matrix (8 B/unit), 1 GPU"), ("_bench_c3shape.json", "C3 weighted, 1 GPU")):
    d = last_json(os.path.join(G, tag1 + name))
    if d:
        out.append(line(label, d))
open(os.path.join(ROOT, "profiles", "r2_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
