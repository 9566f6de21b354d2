#!/usr/bin/env python
"""Reads bench.py's JSON line from stdin and prints the few numbers a tuning loop needs."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
lines = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
if not lines:
    print(tag, "no JSON line"); sys.exit(0)
d = json.loads(lines[-1])
st = d.get("stage_ms_per_step", {})
e = d.get("e2e") or {}
print(tag, "value %.4g" % d["value"], "ms/step %.1f" % d["ms_per_step"], "engine %.1f" % st.get("engine_ms", float("nan")),
      "e2e %.4g" % e["value"] if e else "", "ecc_mean %.15f" % d.get("ecc_mean", float("nan")))
