#!/usr/bin/env python
"""Device-timed us/step of the resident kernel (K5) against the streaming kernel (K1, CUDA graph) on small grids."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from cases import c2_case  # noqa: E402
from strata_fdtd_b200 import _lib  # noqa: E402
from util import build_b200_solver  # noqa: E402


def time_chunk(s, n_steps, reps=5):
    dev = s._sync_to_device()
    lib, h = dev.lib, dev.handle
    src = torch.zeros(n_steps * max(1, len(s._sources)), dtype=torch.float64, device=dev.device)
    rec = torch.zeros(n_steps * max(1, len(s._probes)), dtype=torch.float32, device=dev.device)
    for _ in range(2):
        _lib.check(lib.sb_step_n_async(h, n_steps, src.data_ptr(), rec.data_ptr()))
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(dev.stream)
        _lib.check(lib.sb_step_n_async(h, n_steps, src.data_ptr(), rec.data_ptr()))
        e1.record(dev.stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n_steps)
    _lib.check(lib.sb_synchronize(h))
    return best


def main():
    import os
    out = []
    fast = os.environ.get("RB_FAST") == "1"          # one line per size: the default resident configuration only
    for n in ((64, 100) if fast else (64, 100, 110)):
        for geometry in ((False,) if fast else (False, True)):
            case = c2_case(n, steps=0, with_geometry=geometry)
            for label, opts in (("march_graph", {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_USE_GRAPH: 1}),
                                ("resident_nosplit", {_lib.OPT_KERNEL: _lib.KERNEL_RESIDENT, _lib.OPT_RESIDENT_SPLIT: 0}),
                                ("resident_split", {_lib.OPT_KERNEL: _lib.KERNEL_RESIDENT, _lib.OPT_RESIDENT_SPLIT: 1})):
                if fast and label != "resident_nosplit":
                    continue
                for chunk in ((512,) if fast else (32, 512)):
                    s = build_b200_solver(case)
                    for k, v in opts.items():
                        s.set_kernel_option(k, v)
                    try:
                        us = time_chunk(s, chunk)
                    except Exception as e:           # noqa: BLE001
                        out.append({"n": n, "geometry": geometry, "kernel": label, "chunk": chunk, "error": str(e)[:200]})
                        print(json.dumps(out[-1]), flush=True)
                        s.close()
                        continue
                    out.append({"n": n, "geometry": geometry, "kernel": label, "chunk": chunk, "us_per_step": us,
                                "gcells": n ** 3 / us / 1e3, "hbm_equiv_frac": 32.0 * n ** 3 / (us * 1e-6) / 6540.8e9})
                    print(json.dumps(out[-1]), flush=True)
                    s.close()
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "resident_bench.jsonl").write_text("\n".join(json.dumps(o) for o in out) + "\n")


if __name__ == "__main__":
    main()
