#!/usr/bin/env python
"""Per-kernel totals from `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv`.
usage: python tools/ncu_launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
col = {n: i for i, n in enumerate(rows[hi])}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[col["Kernel Name"]].split("(")[0]
    tot[name] += float(r[col["Metric Value"]].replace(",", "")) / 1e6
    cnt[name] += 1
all_ms = sum(tot.values())
for name, ms in tot.most_common():
    print("%-70s launches %3d %10.3f ms %5.1f%%" % (name[:70], cnt[name], ms, 100 * ms / all_ms))
print("total %.3f ms, %d launches" % (all_ms, sum(cnt.values())))
