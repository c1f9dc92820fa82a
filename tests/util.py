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


class FixtureMembrane:
    """Duck-typed stand-in for the reference's MembraneSource (solver.py:210-370): what the solver touches."""
    source_type = "membrane"

    def __init__(self, src: dict):
        import strata_fdtd_b200 as sb
        m = src["spec"]
        self.center, self.normal_axis, self.injection_type = m["center"], m["normal_axis"], m["injection_type"]
        self.waveform = sb.GaussianPulse(position=(0, 0, 0), frequency=src["frequency"], amplitude=src["amplitude"])
        self._weights = src["weights"]
        self._cached_weights = None
        self._cached_mask = None

    def _check_grid_alignment(self, grid):
        pass

    def get_injection_weights(self, grid):
        return self._weights


def build_b200_solver(case: dict, **solver_kw):
    import strata_fdtd_b200 as sb
    kw = dict(c=case.get("c", 343.0), rho=case.get("rho", 1.2), courant=case.get("courant", 0.95), backend="b200")
    kw.update(solver_kw)
    nu = case.get("nonuniform")
    if nu is None:
        s = sb.FDTDSolver(shape=tuple(case["shape"]), resolution=case["resolution"], **kw)
    else:
        s = sb.FDTDSolver(grid=sb.NonuniformGrid(nu["x_coords"], nu["y_coords"], nu["z_coords"]), **kw)
    if case.get("geometry") is not None:
        g = case["geometry"]
        s.set_geometry(g if callable(g) else np.asarray(g, dtype=bool))
    for b in case.get("pml", []):
        axes = tuple(b.get("axes", ("x", "y", "z")))
        s.add_boundary(sb.PML(depth=b.get("depth", 10), axis="all" if axes == ("x", "y", "z") else axes,
                              max_sigma=b.get("max_sigma"), order=b.get("order", 3)))
    for b in case.get("plane_bcs", []):
        if b["kind"] == "mur":
            s.add_boundary(sb.boundaries.ABCFirstOrder(axis=tuple(b.get("axes", ("x", "y", "z")))))
        else:
            s.add_boundary(sb.boundaries.RadiationImpedance(axis=b["axis"], side=b["side"],
                                                            reflection_coeff=b.get("reflection_coeff"),
                                                            pipe_radius=b.get("pipe_radius")))
    for src in case.get("sources", []):
        kind = src.get("kind", "point")
        if kind == "weighted":
            s.add_source(FixtureMembrane(src))
            continue
        pos = src["position"] if kind == "point" else {"axis": src["axis"], "index": src["index"]}
        s.add_source(sb.GaussianPulse(position=pos, frequency=src["frequency"], bandwidth=src.get("bandwidth"),
                                      amplitude=src.get("amplitude", 1.0), source_type=kind))
    for name, pos in case.get("probes", []):
        s.add_probe(name, position=pos)
    for name, pos, *opt in case.get("mics", []):
        s.add_microphone(position=pos, name=name, **(opt[0] if opt else {}))
    for m in case.get("materials", []):
        poles = []
        for p in m["poles"]:
            if p["type"] == "debye":
                poles.append(sb.Pole(sb.PoleType.DEBYE, p["delta_chi"], p["target"], tau=p["tau"]))
            else:
                poles.append(sb.Pole(sb.PoleType.LORENTZ, p["delta_chi"], p["target"], omega_0=p["omega_0"],
                                     gamma=p["gamma"]))
        s.register_material(sb.PoleMaterial(m.get("name", f"mat{m['id']}"), m["rho_inf"], m["K_inf"], poles),
                            material_id=m["id"])
    if case.get("materials"):
        mid = np.asarray(case["material_id"], dtype=np.uint8)
        for m in case["materials"]:
            s.set_material_region(mid == m["id"], material_id=m["id"])
    return s


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


def build_distributed_solver(case: dict, **kw):
    """DistributedFDTDSolver (one slab per rank, torch.distributed already initialised) from a case dict."""
    import strata_fdtd_b200 as sb
    from strata_fdtd_b200.multi import DistributedFDTDSolver
    base = dict(c=case.get("c", 343.0), rho=case.get("rho", 1.2), courant=case.get("courant", 0.95))
    base.update(kw)
    nu = case.get("nonuniform")
    if nu is None:
        d = DistributedFDTDSolver(shape=tuple(case["shape"]), resolution=case["resolution"], **base)
    else:
        d = DistributedFDTDSolver(grid=sb.NonuniformGrid(nu["x_coords"], nu["y_coords"], nu["z_coords"]), **base)
    if case.get("geometry") is not None:
        g = case["geometry"]
        d.set_geometry(g if callable(g) else np.asarray(g, dtype=bool))
    for b in case.get("pml", []):
        axes = tuple(b.get("axes", ("x", "y", "z")))
        d.add_boundary(sb.PML(depth=b.get("depth", 10), axis="all" if axes == ("x", "y", "z") else axes,
                              max_sigma=b.get("max_sigma"), order=b.get("order", 3)))
    for src in case.get("sources", []):
        kind = src.get("kind", "point")
        pos = src["position"] if kind == "point" else {"axis": src["axis"], "index": src["index"]}
        d.add_source(sb.GaussianPulse(position=pos, frequency=src["frequency"], bandwidth=src.get("bandwidth"),
                                      amplitude=src.get("amplitude", 1.0), source_type=kind))
    for name, pos in case.get("probes", []):
        d.add_probe(name, position=pos)
    return d
