#!/usr/bin/env python
"""Per-CUDA-source-line instruction / sample shares from `ncu --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_lines.py report.ncu-rep [min_share_percent]"""
import csv, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--launch-skip", skip, "--launch-count", "1", "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No") or len(r) < 9: continue
    if r[2] == "-" and r[0].isdigit():
        try: agg[(cur, int(r[0]))] = (r[1], float(r[6]), float(r[7]), float(r[8]))
        except ValueError: pass
ti = sum(v[2] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp-instructions %.4g  samples %d" % (ti, ts))
for k in sorted(agg):
    src, smp, ie, te = agg[k]
    if 100 * ie / ti > thr or 100 * smp / ts > thr:
        print("%-14s %4d inst %5.1f%% smp %5.1f%% thr/inst %4.1f | %s" % (k[0][:14], k[1], 100 * ie / ti, 100 * smp / ts, te / max(ie, 1), src.strip()[:105]))
