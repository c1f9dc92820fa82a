"""Workloads of the hot path as plain dicts, and the builders that turn a dict into a solver.

A *case* names everything ``FDTDSolver`` is told through its public API (grid, geometry, sponge, sources,
probes, microphones, materials).  The BASELINE.json configurations live here (``c1_case`` .. ``c5_case``) so that
``bench.py``, the examples and the tests construct them the same way; the small parity cases stay in
tests/cases.py.  Builders only call the public surface mirrored from /root/reference/src/strata_fdtd/core/solver.py
(constructor :1507, set_geometry :1682, add_boundary :1983, add_source :1782, add_probe :1843,
add_microphone :1887, register_material :2836, set_material_region :2877).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"


def _block_geometry(shape, lo, hi):
    g = np.ones(shape, dtype=bool)
    g[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = False
    return g



# benign dispersive material of SURVEY.md F9 (bounded for >= 1000 steps)
BENIGN_POLES = [
    {"type": "debye", "target": "density", "delta_chi": 0.1, "tau": 1e-4},
    {"type": "debye", "target": "modulus", "delta_chi": 0.1, "tau": 1e-5},
    {"type": "lorentz", "target": "modulus", "delta_chi": 0.05, "omega_0": 2 * np.pi * 2000.0, "gamma": 2 * np.pi * 200.0},
]
SECOND_POLES = [
    {"type": "lorentz", "target": "density", "delta_chi": 0.02, "omega_0": 2 * np.pi * 5000.0, "gamma": 2 * np.pi * 500.0},
    {"type": "debye", "target": "density", "delta_chi": 0.05, "tau": 5e-5},
    {"type": "debye", "target": "modulus", "delta_chi": 0.08, "tau": 2e-5},
]


def enclosure_air_mask(X, Y, Z, extent):
    """Air mask (True = air) of a ported loudspeaker enclosure filling the central 60 % of the domain.

    Own closed-form construction with the proportions of the reference's CSG example
    (examples/sdf_csg/ported_enclosure.py:34-121: 200x200x300 mm box, 18 mm walls, 130 mm driver cut-out,
    50 mm flared port through the front wall), evaluated at cell centres X, Y, Z (metres; any sub-range of
    planes along X) so that slabs can voxelise only what they own.  Box axis = grid axis 0, front wall at low x.
    """
    Lx, Ly, Lz = extent
    cx, cy, cz = Lx / 2, Ly / 2, Lz / 2
    hx, hy, hz = 0.3 * Lx, 0.3 * Ly, 0.3 * Lz                    # half sizes of the outer box
    t = 0.06 * 2 * min(hx, hy, hz)                               # wall thickness (18/300 of the smallest side)
    x = np.asarray(X)[:, None, None]; y = np.asarray(Y)[None, :, None]; z = np.asarray(Z)[None, None, :]
    outer = (np.abs(x - cx) <= hx) & (np.abs(y - cy) <= hy) & (np.abs(z - cz) <= hz)
    inner = (np.abs(x - cx) < hx - t) & (np.abs(y - cy) < hy - t) & (np.abs(z - cz) < hz - t)
    solid = outer & ~inner
    depth = x - (cx - hx)                                        # distance behind the front face
    in_front_wall = (depth >= 0) & (depth <= t)
    r_driver = 0.325 * 2 * hy
    r2_d = (y - cy) ** 2 + (z - cz) ** 2
    solid &= ~(in_front_wall & (r2_d < r_driver ** 2))
    r_port = 0.125 * 2 * hy
    flare = 1.5 * r_port - 0.5 * r_port * np.clip(depth / (1.6667 * t), 0.0, 1.0)   # 30 mm flare vs 18 mm wall
    r2_p = (y - cy) ** 2 + (z - (cz + 0.6 * hz)) ** 2
    solid &= ~(in_front_wall & (r2_p < flare ** 2))
    return ~solid


def reference_enclosure_mask_path(shape) -> Path:
    return GOLDEN / "c4_enclosure_mask_{}x{}x{}.npz".format(*shape)


def load_reference_enclosure_mask(shape):
    """Air mask of the REFERENCE's own CSG enclosure (examples/sdf_csg/ported_enclosure.py:34-121, voxelised with
    geometry/sdf.py:99-125 in x-chunks by oracle/make_golden_large.py), stored bit-packed; None if not generated."""
    f = reference_enclosure_mask_path(shape)
    if not f.exists():
        return None
    with np.load(f) as z:
        assert tuple(int(q) for q in z["shape"]) == tuple(shape)
        n = int(np.prod(shape, dtype=np.int64))
        return np.unpackbits(z["bits"])[:n].reshape(shape).astype(bool)


def c4_probes(shape, enclosure: str):
    nx, ny, nz = shape
    if enclosure == "reference":        # the example's front baffle (driver + port) is its z = 0 face
        return [("in_centre", (nx // 2, ny // 2, nz // 2)), ("in_back", (nx // 2, ny // 2, int(0.72 * nz))),
                ("in_corner", (int(0.3 * nx), int(0.3 * ny), int(0.3 * nz))),
                ("driver_mouth", (nx // 2, ny // 2, int(0.19 * nz))), ("port_mouth", (nx // 2, int(0.68 * ny), int(0.19 * nz))),
                ("front_far", (nx // 2, ny // 2, int(0.08 * nz))), ("side", (nx // 2, int(0.1 * ny), nz // 2)),
                ("behind", (int(0.92 * nx), ny // 2, nz // 2))]
    return [("in_centre", (nx // 2, ny // 2, nz // 2)), ("in_back", (int(0.72 * nx), ny // 2, nz // 2)),
            ("in_corner", (int(0.3 * nx), int(0.3 * ny), int(0.3 * nz))),
            ("driver_mouth", (int(0.19 * nx), ny // 2, nz // 2)), ("port_mouth", (int(0.19 * nx), ny // 2, int(0.68 * nz))),
            ("front_far", (int(0.08 * nx), ny // 2, nz // 2)), ("side", (nx // 2, int(0.1 * ny), nz // 2)),
            ("behind", (int(0.92 * nx), ny // 2, nz // 2))]


def c4_case(shape=(1024, 512, 512), steps: int = 1000, stretch_x: float = 1.002, materialise: bool = True,
            enclosure: str = "closed_form") -> dict:
    """BASELINE config 4: nonuniform grid (axis 0 stretched from the centre), ported-enclosure
    rigid geometry, PML(10), source inside the box, 8 probes.

    ``enclosure="reference"``: the reference's own CSG enclosure scaled to the central 60 % of the domain (bit-packed
    fixture, see ``load_reference_enclosure_mask``); ``"closed_form"``: this repo's own construction with the same
    proportions, for sizes without a fixture.  ``materialise=False`` returns the geometry as a callable
    f(i_lo, i_hi) so that a slab only ever holds its own planes."""
    from .grid import NonuniformGrid
    g = NonuniformGrid.from_stretch(shape=shape, base_resolution=1e-3, stretch_x=stretch_x, center_fine=True)
    extent = g.physical_extent()
    X, Y, Z = g.x_coords, g.y_coords, g.z_coords
    nx, ny, nz = shape
    if enclosure == "reference":
        mask = load_reference_enclosure_mask(shape)
        if mask is None:
            raise FileNotFoundError(f"{reference_enclosure_mask_path(shape)} missing: run oracle/make_golden_large.py c4mask")

        def geom(i_lo, i_hi):
            return mask[i_lo:i_hi]
    elif enclosure == "closed_form":
        def geom(i_lo, i_hi, _chunk=16):
            out = np.empty((i_hi - i_lo,) + tuple(shape[1:]), dtype=bool)
            for a in range(i_lo, i_hi, _chunk):
                b = min(a + _chunk, i_hi)
                out[a - i_lo:b - i_lo] = enclosure_air_mask(X[a:b], Y, Z, extent)
            return out
    else:
        raise ValueError(f"unknown enclosure {enclosure!r}")
    return dict(nonuniform=dict(x_coords=X, y_coords=Y, z_coords=Z), shape=tuple(shape), steps=steps,
                geometry=geom(0, nx) if materialise else geom, pml=[dict(depth=10)],
                sources=[dict(kind="point", position=(int(0.6 * nx), ny // 2, nz // 2), frequency=2000.0)],
                probes=c4_probes(shape, enclosure))


def c5_case(n_gpus: int = 1, steps: int = 0, planes_per_gpu: int = 256, ny: int = 2048, nz: int = 2048) -> dict:
    """BASELINE config 5, weak scaling: every GPU owns a planes_per_gpu x ny x nz slab."""
    nx = planes_per_gpu * n_gpus
    probes = [(f"p{q}", (min(nx - 1, (2 * q + 1) * nx // 16), ny // 2 + (ny // 32) * (q - 4), nz // 2 - nz // 85))
              for q in range(8)]
    return dict(shape=(nx, ny, nz), resolution=1e-3, steps=steps, pml=[dict(depth=10)],
                sources=[dict(kind="point", position=(nx // 2, ny // 2, nz // 2), frequency=1000.0)], probes=probes)


def c1_case(steps: int = 1000) -> dict:
    """BASELINE config 1 as worded: 100^3, 1 mm, PML 10, 1 kHz Gaussian pulse, 1 probe."""
    return dict(shape=(100, 100, 100), resolution=1e-3, steps=steps, pml=[dict(depth=10)],
                sources=[dict(kind="point", position=(25, 50, 50), frequency=1000.0)],
                probes=[("probe", (75, 50, 50))])


def c2_case(n: int = 200, steps: int = 1000, with_geometry: bool = False) -> dict:
    """BASELINE config 2: N^3 uniform + PML(10), source at (N/4, N/2, N/2)."""
    c = dict(shape=(n, n, n), resolution=1e-3, steps=steps, pml=[dict(depth=10)],
             sources=[dict(kind="point", position=(n // 4, n // 2, n // 2), frequency=1000.0)],
             probes=[("probe", (3 * n // 4, n // 2, n // 2))])
    if with_geometry:
        a, b = int(0.45 * n), int(0.55 * n)
        c["geometry"] = _block_geometry((n, n, n), (a, a, a), (b, b, b))
    return c


def c3_case(n: int = 512, steps: int = 1000, slab: bool = False) -> dict:
    """BASELINE config 3: N^3 + PML + ADE material sphere (radius N/10) + 64 probes."""
    sh = (n, n, n)
    mid = np.zeros(sh, dtype=np.uint8)
    if slab:
        mid[3 * n // 4:, :, :] = 1
    else:
        c = n // 2
        r = n / 10.0
        # cell-centre SDF < 0, evaluated plane by plane to bound temporaries
        j, k = np.ogrid[:n, :n]
        for i in range(max(0, int(c - r) - 1), min(n, int(c + r) + 2)):
            mid[i][((i - c) ** 2 + (j - c) ** 2 + (k - c) ** 2) < r * r] = 1
    s = n / 512.0
    probes = [(f"p{a}{b}", (int(384 * s), int((32 + 64 * a) * s), int((32 + 64 * b) * s)))
              for a in range(8) for b in range(8)]
    return dict(shape=sh, resolution=1e-3, steps=steps, pml=[dict(depth=10)],
                sources=[dict(kind="point", position=(int(77 * s), n // 2, n // 2), frequency=40e3)],
                probes=probes,
                materials=[dict(id=1, rho_inf=1.2, K_inf=1.2 * 343.0 ** 2, poles=BENIGN_POLES)],
                material_id=mid)


class WeightedSource:
    """Duck-typed stand-in for the reference's MembraneSource (solver.py:210-370): what the solver touches.
    Built from a case entry of kind "weighted" that carries the injection-weight array."""
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


def build_solver(case: dict, **solver_kw):
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
            s.add_source(WeightedSource(src))
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


def build_distributed_solver(case: dict, **kw):
    """DistributedFDTDSolver (one slab per rank, torch.distributed already initialised) from a case dict."""
    import strata_fdtd_b200 as sb
    from .multi import DistributedFDTDSolver
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
