#!/usr/bin/env python
"""Where does FDTDSolver.run() spend host time on a small grid?  cProfile over repeated short runs (config 1)."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from strata_fdtd_b200.workloads import build_solver, c1_case  # noqa: E402

s = build_solver(c1_case(0), distributed=False)
s.run(steps=200); s.run(steps=200)
n, reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200, 100
t0 = time.perf_counter()
for _ in range(reps):
    s.run(steps=n)
dt = (time.perf_counter() - t0) / reps
print(f"run(steps={n}): {dt * 1e6:.0f} us per call")
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
    s.run(steps=n)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
