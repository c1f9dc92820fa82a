"""BASELINE config 1 as worded: 100^3 cells at 1 mm, PML thickness 10, 1 kHz Gaussian pulse, 1 probe, 1000 steps.

Written against the reference's public API only (``from strata_fdtd import ...``); run it unchanged with

    python -m strata_fdtd_b200 examples/basic_pulse.py            # b200 backend, this package alone
    python examples/basic_pulse.py                                # where the reference package is installed
"""
import numpy as np

from strata_fdtd import PML, FDTDSolver, GaussianPulse

solver = FDTDSolver(shape=(100, 100, 100), resolution=1e-3)
solver.add_boundary(PML(depth=10, axis="all"))
solver.add_source(GaussianPulse(position=(25, 50, 50), frequency=1000.0))
solver.add_probe("downstream", position=(75, 50, 50))

print(f"grid {solver.grid.shape}, dt = {solver.dt * 1e9:.1f} ns, native backend: {solver.using_native}")
solver.run(duration=1000 * solver.dt * 0.9999, output_file="basic_pulse_results.h5")

trace = solver.get_probe_data("downstream")["downstream"]
print(f"steps: {solver.step_count}, probe samples: {len(trace)}, peak |p| at probe: {np.abs(trace).max():.4e} Pa "
      f"at t = {np.argmax(np.abs(trace)) * solver.dt * 1e3:.3f} ms")
if hasattr(solver, "last_run_stats"):
    st = solver.last_run_stats
    print(f"{st['cell_updates_per_s'] / 1e9:.1f} Gcell-updates/s end to end, {st['kernel_launches']} kernel launches")
