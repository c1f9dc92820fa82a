#!/usr/bin/env python
"""Device-timed us/step: step-by-step marching kernel (CUDA graph) vs the step-pipelined launch (K6), mid-sized grids."""
from __future__ import annotations

import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
from cases import c2_case  # noqa: E402
from resident_bench import time_chunk  # noqa: E402
from strata_fdtd_b200 import _lib  # noqa: E402
from util import build_b200_solver  # noqa: E402


def main():
    out = []
    sizes = [int(a) for a in sys.argv[1:]] or [100, 128, 160, 200, 256, 300, 320, 400]
    for n in sizes:
        for label, opts in (("march_graph", {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_USE_GRAPH: 1}),
                            ("pipeline", {_lib.OPT_KERNEL: _lib.KERNEL_PIPELINE}),
                            ("auto", {})):
            s = build_b200_solver(c2_case(n, steps=0))
            for k, v in opts.items():
                s.set_kernel_option(k, v)
            try:
                us = time_chunk(s, 100, reps=3)
                out.append({"n": n, "kernel": label, "us_per_step": us, "gcells": n ** 3 / us / 1e3,
                            "frac_hbm": 32.0 * n ** 3 / (us * 1e-6) / 6540.8e9})
            except Exception as e:          # noqa: BLE001
                out.append({"n": n, "kernel": label, "error": str(e)[:160]})
            print(json.dumps(out[-1]), flush=True)
            s.close()
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "pipeline_bench.jsonl").write_text("\n".join(json.dumps(o) for o in out) + "\n")


if __name__ == "__main__":
    main()
