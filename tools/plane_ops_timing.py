#!/usr/bin/env python
"""us/step of grids with Mur planes on all six faces: the step-by-step path (K1 + one K4 launch per plane + K3, in a
CUDA graph) against the step-pipelined kernel K6 with the planes applied inside its tiles.  Prints one line per size."""
import sys
import time

import torch

import strata_fdtd_b200 as sb
from strata_fdtd_b200 import _lib


def run(n, kernel, steps=600, planes=True):
    s = sb.FDTDSolver(shape=(n, n, n), resolution=1e-3, backend="b200", chunk_steps=100)
    if planes:
        s.add_boundary(sb.boundaries.ABCFirstOrder())
    s.add_source(sb.GaussianPulse(position=(n // 2, n // 2, n // 2), frequency=20e3))
    s.add_probe("a", (n // 4, n // 2, n // 2))
    s.set_kernel_option(_lib.OPT_KERNEL, kernel)
    s.run(steps=200)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run(steps=steps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    v = s.device_stats()["kernel_variant"]
    s.close()
    return dt * 1e6, v


if __name__ == "__main__":
    for n in [int(q) for q in sys.argv[1:]] or [64, 100, 160, 200, 256, 300]:
        a, va = run(n, _lib.KERNEL_MARCH)
        b, vb = run(n, _lib.KERNEL_PIPELINE)
        c, vc = run(n, _lib.KERNEL_AUTO)
        d, vd = run(n, _lib.KERNEL_AUTO, planes=False)
        print(f"{n}^3 six Mur planes: step-by-step {a:8.2f} us  pipelined {b:8.2f} us  auto {c:8.2f} us (variant {vc});  "
              f"no planes, auto {d:8.2f} us (variant {vd})   [{n**3 / b / 1e3:.1f} Gcell/s pipelined]", flush=True)
