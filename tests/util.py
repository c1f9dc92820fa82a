"""Build the b200 solver from a tests/cases.py case through its public API."""
from __future__ import annotations

import hashlib

import numpy as np


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


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
        s.set_geometry(np.asarray(case["geometry"], dtype=bool))
    for b in case.get("pml", []):
        axes = tuple(b.get("axes", ("x", "y", "z")))
        s.add_boundary(sb.PML(depth=b.get("depth", 10), axis="all" if axes == ("x", "y", "z") else axes,
                              max_sigma=b.get("max_sigma"), order=b.get("order", 3)))
    for src in case.get("sources", []):
        kind = src.get("kind", "point")
        pos = src["position"] if kind == "point" else {"axis": src["axis"], "index": src["index"]}
        s.add_source(sb.GaussianPulse(position=pos, frequency=src["frequency"], bandwidth=src.get("bandwidth"),
                                      amplitude=src.get("amplitude", 1.0), source_type=kind))
    for name, pos in case.get("probes", []):
        s.add_probe(name, position=pos)
    for name, pos in case.get("mics", []):
        s.add_microphone(position=pos, name=name)
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
    for name, _ in o.mics:
        a, b = s.microphones[name].get_waveform(), o.mic_array(name)
        assert np.array_equal(a, b), f"{what}: mic {name} differs (max|d|={np.abs(a - b).max():.3e})"
