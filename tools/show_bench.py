#!/usr/bin/env python
"""One-line digest of a bench.py JSON line read from stdin: tools/show_bench.py <label>"""
import json
import sys

lines = [q for q in sys.stdin.read().strip().splitlines() if q.startswith("{")]
if not lines:
    print(sys.argv[1], ": no JSON line")
    sys.exit(0)
d = json.loads(lines[-1])
print(sys.argv[1], ":", round(d["value"], 1), "Gcell/s  e2e", round(d["e2e"]["value"], 1), " ms/step", round(d["ms_per_step"], 4),
      "kernel_ms", round(d["roofline"]["kernel_ms_mean"], 4), "frac", round(d["roofline"]["frac"], 3), d["config"].get("launch_shape"))
