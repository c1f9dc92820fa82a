"""BASELINE config 3 in miniature: PML + a frequency-dependent (ADE) material sphere + a lattice of probes.

The reference's own example assigns a pole-less SimpleMaterial, which is a no-op in every backend
(SURVEY.md F8); this one uses two Debye poles and one Lorentz pole that stay bounded.
Usage:  python -m strata_fdtd_b200 examples/material_sphere.py [N]      (default N = 128; 512 = config 3)
"""
import sys

import numpy as np

from strata_fdtd import PML, FDTDSolver, GaussianPulse
from strata_fdtd.materials import Pole, PoleType, SimpleMaterial

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
solver = FDTDSolver(shape=(n, n, n), resolution=1e-3)
solver.add_boundary(PML(depth=10))

absorber = SimpleMaterial(name="benign_absorber", _rho=1.2, _c=343.0, _poles=[
    Pole(PoleType.DEBYE, delta_chi=0.1, target="density", tau=1e-4),
    Pole(PoleType.DEBYE, delta_chi=0.1, target="modulus", tau=1e-5),
    Pole(PoleType.LORENTZ, delta_chi=0.05, target="modulus", omega_0=2 * np.pi * 2000.0, gamma=2 * np.pi * 200.0),
])
mat_id = solver.register_material(absorber)
c, r = n // 2, n / 10.0
i, j, k = np.ogrid[:n, :n, :n]
solver.set_material_region(((i - c) ** 2 + (j - c) ** 2 + (k - c) ** 2) < r * r, mat_id)

solver.add_source(GaussianPulse(position=(int(0.15 * n), c, c), frequency=40e3))
for a in range(8):
    for b in range(8):
        solver.add_probe(f"p{a}{b}", position=(3 * n // 4, (n * (1 + 2 * a)) // 16, (n * (1 + 2 * b)) // 16))

solver.run(duration=400 * solver.dt * 0.9999)
peak = max(float(np.abs(v).max()) for v in solver.get_probe_data().values())
print(f"{n}^3 cells, {solver.step_count} steps, {len(solver.get_probe_data())} probes, peak |p| = {peak:.3e} Pa")
if hasattr(solver, "last_run_stats"):
    print(f"{solver.last_run_stats['cell_updates_per_s'] / 1e9:.1f} Gcell-updates/s end to end")
