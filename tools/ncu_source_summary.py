#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): stall reasons, opcode mix, hottest instructions.
usage: python tools/ncu_source_summary.py source.csv [top_n]"""
import csv
import collections
import sys

path = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
seen = set()  # some ncu versions print every SASS line twice (once per view): keep the first of each address
data = [r for r in data if not (r[0] in seen or seen.add(r[0]))]


def f(r, name):
    try:
        return float(r[col[name]])
    except ValueError:
        return 0.0


tot_samples = sum(f(r, "# Samples") for r in data)
tot_inst = sum(f(r, "Instructions Executed") for r in data)
tot_thr = sum(f(r, "Thread Instructions Executed") for r in data)
print("SASS lines %d  samples %d  warp-instructions %.4g  thread-instructions %.4g (avg %.1f threads/instr)" %
      (len(data), tot_samples, tot_inst, tot_thr, tot_thr / max(tot_inst, 1)))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(f(r, n) for r in data) for n in stalls}
print("\nstall reasons (all samples):")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-24s %8d  %5.1f%%" % (n, v, 100 * v / max(tot_samples, 1)))
ops = collections.Counter()
ops_s = collections.Counter()
for r in data:
    parts = r[col["Source"]].split()
    op = parts[0] if parts and not parts[0].startswith("@") else (parts[1] if len(parts) > 1 else "?")
    op = op.split(".")[0]
    ops[op] += f(r, "Instructions Executed")
    ops_s[op] += f(r, "# Samples")
print("\nopcode mix (warp-instructions executed, samples):")
for op, v in ops.most_common(28):
    print("  %-10s %12.4g %5.1f%%   samples %5.1f%%" % (op, v, 100 * v / tot_inst, 100 * ops_s[op] / max(tot_samples, 1)))
print("\nhottest instructions by samples:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top_n]:
    top = max(stalls, key=lambda n: f(r, n))
    print("  %6d %5.2f%%  exec %10.4g  %-18s %s" % (f(r, "# Samples"), 100 * f(r, "# Samples") / max(tot_samples, 1),
                                                 f(r, "Instructions Executed"), top, r[col["Source"]][:110]))
