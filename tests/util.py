"""Build the b200 solver from a tests/cases.py case through its public API."""
from __future__ import annotations

import hashlib

import numpy as np


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_membrane_case():
    """Membrane case rebuilt from the fixture (weights produced by the reference's own membrane classes)."""
    from pathlib import Path
    from cases import MEMBRANE_SPEC, membrane_case
    g = np.load(Path(__file__).parent / "golden" / "membranes.npz")
    weights = []
    for q in range(len(MEMBRANE_SPEC["membranes"])):
        w = np.zeros(int(np.prod(MEMBRANE_SPEC["shape"])), dtype=np.float64)
        w[g[f"weights_idx_{q}"]] = g[f"weights_val_{q}"]
        weights.append(w.reshape(MEMBRANE_SPEC["shape"]))
    return membrane_case(weights), g


from strata_fdtd_b200.workloads import WeightedSource as FixtureMembrane  # noqa: E402,F401
from strata_fdtd_b200.workloads import build_distributed_solver  # noqa: E402,F401
from strata_fdtd_b200.workloads import build_solver as build_b200_solver  # noqa: E402,F401


def assert_same_as_oracle(s, o, what=""):
    """Bit-exact comparison of fields and traces between a b200 solver and an OracleSolver."""
    for f in ("p", "vx", "vy", "vz"):
        a, b = s.get_field(f), getattr(o, f)
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError(f"{what}: field {f} differs at {len(bad)} cells, first {bad[0]}, "
                                 f"got {a[tuple(bad[0])]!r} want {b[tuple(bad[0])]!r}, "
                                 f"max|d|={np.abs(a - b).max():.3e} max|ref|={np.abs(b).max():.3e}")
    for name, _ in o.probes:
        a, b = s.get_probe_data(name)[name], o.probe_array(name)
        assert np.array_equal(a, b), f"{what}: probe {name} differs (max|d|={np.abs(a - b).max():.3e})"
    for name, *_ in o.mics:
        a, b = s.microphones[name].get_waveform(), o.mic_array(name)
        assert np.array_equal(a, b), f"{what}: mic {name} differs (max|d|={np.abs(a - b).max():.3e})"
