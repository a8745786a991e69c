#!/usr/bin/env python
"""Reads bench.py's JSON line on stdin and prints a one-line summary (tools/perf_sweep.sh)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    try:
        d = json.loads(line)
    except Exception:
        continue
    st = {k: round(v, 1) for k, v in d.get("roofline", {}).get("stage_ms", {}).items()}
    print(tag, "value %.4g" % d["value"], "ms/step %.1f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], st)
