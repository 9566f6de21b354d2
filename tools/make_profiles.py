#!/usr/bin/env python
"""Builds the tracked summaries under profiles/ from one tools/round_measure.sh run (gpurun_out/TAG_*).
usage: python tools/make_profiles.py TAG [ROUND_PREFIX]   (needs ncu on PATH to read the .ncu-rep)"""
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1]
rp = sys.argv[2] if len(sys.argv) > 2 else "r2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def last_json(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


# ---- bench lines
bench = last_json(os.path.join(G, tag + "_bench_n1.json"))
json.dump(bench, open(os.path.join(P, rp + "_bench_c5_n1.json"), "w"))
ref = os.path.join(G, tag + "_reference_arm.json")
if os.path.exists(ref) and last_json(ref):
    json.dump(last_json(ref), open(os.path.join(P, rp + "_bench_c5_reference_arm.json"), "w"))

# ---- launch list
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline  (C5, 3 steps captured; tools/profile_c5.sh)",
       "# wave_kernel<ECC,SIM,UNITW,STREAMING,MODE>: <..,1,m> = streaming items (clusters > 3456 cells), <..,0,m> = resident items; MODE 0 = per cell (C5), 1 = ratio tables, 2 = matrix from the tables.",
       "# bench.py of the same build (CUDA events, no profiler): engine (qinfo + both wave launches) %.1f ms of a %.1f ms step = %.1f %%" %
       (bench["stage_ms_per_step"]["engine_ms"], bench["ms_per_step"], 100 * bench["stage_ms_per_step"]["engine_ms"] / bench["ms_per_step"])]
out.append(run([sys.executable, os.path.join(ROOT, "tools", "ncu_launch_summary.py"), os.path.join(G, tag + "_launches.csv")]))
open(os.path.join(P, rp + "_ncu_launches_c5.txt"), "w").write("\n".join(out))

# ---- full capture: launch 0 = streaming items, launch 1 = resident items
rep = os.path.join(G, tag + "_wave.ncu-rep")
WANT = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_op_shared_atom.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__cycles_active.avg',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_red.sum']
KEEP = ("wave_kernel", "Duration", "Throughput", "Ipc", "Hit Rate", "Eligible", "Issued Warp", "Registers Per", "Shared Memory Config",
        "Dynamic Shared", "Block Size", "Grid Size", "Theoretical Occ", "Achieved Occ", "Local Speedup")
txt = ["# ncu --set full --clock-control none --import-source on -k regex:wave_kernel --launch-skip 2 -c 2 : C5 (500 x 1M), second step's two launches (tools/profile_c5.sh)"]
traffic = 0.0
parts = []
onchip = {"smsp__inst_executed.sum": [], "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": [], "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": [],
          "smsp__inst_executed_op_shared_atom.sum": []}
for k in (1, 0):
    det = run(["ncu", "-i", rep, "--launch-skip", str(k), "--launch-count", "1", "--page", "details"])
    txt += [l[:120] for l in det.splitlines() if any(w in l for w in KEEP)]
    raw = list(csv.reader(run(["ncu", "-i", rep, "--launch-skip", str(k), "--launch-count", "1", "--page", "raw", "--csv"]).splitlines()))
    txt += ["", "# raw counters"]
    vals = {}
    for h, u, v in zip(*raw[:3]):
        vals[h] = (u, v)
        if h in WANT:
            txt.append("%-70s %-16s %s" % (h, u, v))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    b = sum(float(vals[h][1].replace(",", "")) * scale[vals[h][0]] for h in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    ms = float(vals["gpu__time_duration.sum"][1].replace(",", "")) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[vals["gpu__time_duration.sum"][0]]
    traffic += b
    for h in onchip:
        onchip[h].append(float(vals[h][1].replace(",", "")))
    parts.append("%s items %.2f GB in %.1f ms" % ("resident" if k == 1 else "streaming", b / 1e9, ms))
    src = os.path.join("/tmp", "%s_src_%d.csv" % (tag, k))
    open(src, "w").write(run(["ncu", "-i", rep, "--launch-skip", str(k), "--launch-count", "1", "--page", "source", "--csv"]))
    txt += ["", "# SASS-level summary (tools/ncu_source_summary.py)", run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_summary.py"), src, "30"]),
            "=" * 100]
open(os.path.join(P, rp + "_ncu_wave_c5_summary.txt"), "w").write("\n".join(txt))
json.dump({"workload": bench["config"]["workload"], "n_gpus": 1, "dram_bytes_per_launch": int(traffic),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of the wave kernel's two launches of one step (%s), ncu --set full, profiles/%s_ncu_wave_c5_summary.txt"
                   % ("; ".join(parts), rp)}, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)

# ---- the on-chip ceilings bench.py reports (roofline_onchip): warp-instructions and shared-memory wavefronts per 32 units
units32 = bench["config"]["pairs"] * bench["config"]["cells"] / 32.0
json.dump({"workload": bench["config"]["workload"], "n_gpus": 1,
           "warp_instructions_per_32_units": sum(onchip["smsp__inst_executed.sum"]) / units32,
           "smem_wavefronts_per_32_units": sum(onchip["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]) / units32,
           "sm_count": 148, "sm_mhz": (bench.get("clocks") or {}).get("sm_mhz", 1965.0),
           "counters": dict(onchip, launches=["wave_kernel resident", "wave_kernel streaming"]),
           "source": "profiles/%s_ncu_wave_c5_summary.txt (ncu --set full of both launches of one C5 step)" % rp,
           "model": "issue ceiling = 4 warp-instructions/clk/SM x SMs x clock x 32 / (warp-instructions per 32 units); shared-memory ceiling = "
                    "1 wavefront/clk/SM x SMs x clock x 32 / (wavefronts per 32 units)"},
          open(os.path.join(P, "roofline_onchip.json"), "w"), indent=1)

# ---- ablation, other configs
abl = os.path.join(G, tag + "_ablate.txt")
if os.path.exists(abl):
    lines = [l for l in open(abl).read().splitlines() if l.startswith("#") or l.startswith("CA_ABLATE")]
    hdr = ["# wave kernel ablation on C5 (tuning build tools/build_variant.sh abl -DCA_ABLATE, which also carries the buffer-not-ready counters;",
           "# tools/ablate_c5.sh).  engine = qinfo + both wave launches, ms per step; the streaming launch (about 44 ms) ignores the switches.",
           "# The switched-off variants compute garbage (ecc_mean shows it); only their time is meaningful."]
    extra = []
    o0 = os.path.join(G, "abl_out_0.txt")
    if os.path.exists(o0):
        seen = sorted(set(l for l in open(o0).read().splitlines() if l.startswith("[kind")))
        res = [l for l in seen if "resident" in l]
        stre = [l for l in seen if "streaming" in l]
        extra = ["", "# per-block counters of the full run (blocks 3 and 77; one line per step):"] + [l[:230] for l in res[-8:] + stre[-4:]]
    open(os.path.join(P, rp + "_ablation_wave_c5.txt"), "w").write("\n".join(hdr + lines + extra) + "\n")
other = []
for name, label in (("_bench_uniform.json", "C5 uniform labels (worst-case family)"), ("_bench_c4shape.json", "C4 shape (1000 x 50k) through the consistency path"),
                    ("_bench_c4_sim_matrix.json", "C4 element_sim_matrix (1000 x 50k), matrix only: 8 B per unit"),
                    ("_bench_c3shape.json", "C3 shape (100 x 100k, weighted)")):
    pth = os.path.join(G, tag + name)
    d = last_json(pth) if os.path.exists(pth) else None
    if d:
        other.append("%-62s value %.4g pair*cells/s (%.2f ms/step)  e2e %.4g (%.2f ms)  roofline frac %.2f (%.0f B/unit)  stages %s" %
                     (label, d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", float("nan")), (d.get("e2e") or {}).get("ms_per_step", float("nan")),
                      d["roofline"]["frac"], d["roofline"]["algorithmic_bytes_per_unit"], {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}))
c34 = os.path.join(G, tag + "_c3_c4.txt")
if os.path.exists(c34):
    other += ["", "# tools/run_c3_c4.py: C4 element_sim_matrix and C3 weighted consistency + matrix through the Python mirror of the R API, checked against the oracle"]
    other += [l[:300] for l in open(c34).read().splitlines() if l.strip()]
if other:
    open(os.path.join(P, rp + "_bench_other_configs.txt"), "w").write("\n".join(other) + "\n")
dd = os.path.join(G, tag + "_c1_c2_dedup.txt")
if os.path.exists(dd):
    open(os.path.join(P, rp + "_c1_c2_dedup.txt"), "w").write(
        "# tools/run_c1_c2_dedup.py on the GPU box: C1 / C2 as latencies, the identical-partition pre-pass on its own (SURVEY.md section 8d)\n" + open(dd).read())
print("profiles written for", tag)
