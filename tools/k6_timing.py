#!/usr/bin/env python
"""us/step of the step-pipelined kernel K6 on n^3 grids with PML(10) (BASELINE config 2 is n = 200), optionally with the
solid block of profile_fdtd.py.  STRATA_B200_LIB_OVERRIDE selects a library build (launch-bound experiments)."""
import sys
import time

import numpy as np
import torch

import strata_fdtd_b200 as sb
from strata_fdtd_b200 import _lib


def run(n, geom, steps=1000):
    s = sb.FDTDSolver(shape=(n, n, n), resolution=1e-3, backend="b200", chunk_steps=250)
    s.add_boundary(sb.PML(depth=10))
    if geom:
        g = np.ones((n, n, n), dtype=bool)
        g[9 * n // 20: 11 * n // 20, 9 * n // 20: 11 * n // 20, 9 * n // 20: 11 * n // 20] = False
        s.set_geometry(g)
    s.add_source(sb.GaussianPulse(position=(n // 4, n // 2, n // 2), frequency=1e3))
    s.add_probe("a", (3 * n // 4, n // 2, n // 2))
    s.set_kernel_option(_lib.OPT_KERNEL, _lib.KERNEL_PIPELINE)
    s.run(steps=250)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run(steps=steps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    s.close()
    return dt * 1e6


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "library"
    out = []
    for n in (200, 256, 300):
        for geom in (False, True):
            us = run(n, geom)
            out.append(f"{n}^3{'+block' if geom else ''} {us:.2f} us ({n ** 3 / us / 1e3:.1f})")
    print(tag, " | ".join(out), flush=True)
