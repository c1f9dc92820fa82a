#!/usr/bin/env python
"""Kernel-variant sweep on one GPU: device-timed Gcell-updates/s for the fused step (CUDA events)."""
from __future__ import annotations

import argparse
import itertools
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from cases import c2_case  # noqa: E402
from strata_fdtd_b200 import _lib  # noqa: E402
from util import build_b200_solver  # noqa: E402


def time_steps(s, n_steps, reps=3):
    dev = s._sync_to_device()
    lib, h = dev.lib, dev.handle
    n_src = max(1, len(s._sources)); n_rec = max(1, len(s._probes) + len(s._microphones))
    src = torch.zeros(n_steps * n_src, dtype=torch.float64, device=dev.device)
    rec = torch.zeros(n_steps * n_rec, dtype=torch.float32, device=dev.device)
    _lib.check(lib.sb_step_n_async(h, 3, src.data_ptr(), rec.data_ptr()))
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(dev.stream)
        _lib.check(lib.sb_step_n_async(h, n_steps, src.data_ptr(), rec.data_ptr()))
        e1.record(dev.stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n_steps)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--geometry", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    a = ap.parse_args()
    case = c2_case(a.n, steps=0, with_geometry=a.geometry)
    s = build_b200_solver(case)
    if a.graph:
        s.set_kernel_option(_lib.OPT_USE_GRAPH, 1)
    cells = a.n ** 3
    rows = []
    combos = [("naive", 0, 0, 0, 0)]
    rjs, wjs, wks, chunks = ([2], [8], [1], [0]) if a.quick else ([1, 2, 4], [2, 4, 8], [1, 2, 4], [0, 16, 64])
    if a.small:
        rjs, wjs, wks, chunks = [1, 2], [1, 2, 4, 8], [1], [0, 2, 4, 8, 16, 32]
    for rj, wj, wk, ch in itertools.product(rjs, wjs, wks, chunks):
        if wj * wk > 8:
            continue
        combos.append(("march", rj, wj, wk, ch))
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    with open(a.out, "a") as f:
        for kind, rj, wj, wk, ch in combos:
            s.set_kernel_option(_lib.OPT_KERNEL, _lib.KERNEL_NAIVE if kind == "naive" else _lib.KERNEL_MARCH)
            if kind == "march":
                s.set_kernel_option(_lib.OPT_ROWS_PER_THREAD, rj); s.set_kernel_option(_lib.OPT_WARPS_J, wj)
                s.set_kernel_option(_lib.OPT_WARPS_K, wk); s.set_kernel_option(_lib.OPT_CHUNK_I, ch)
            ms = time_steps(s, a.steps)
            row = dict(n=a.n, geometry=a.geometry, kind=kind, rj=rj, wj=wj, wk=wk, chunk=ch, ms_per_step=ms,
                       gcells=cells / ms / 1e6, frac_hbm=cells * 32 / (ms * 1e-3) / 6540.8e9)
            rows.append(row)
            f.write(json.dumps(row) + "\n"); f.flush()
            print(json.dumps(row), flush=True)
    best = max(rows, key=lambda r: r["gcells"])
    print("BEST", json.dumps(best))


if __name__ == "__main__":
    main()
