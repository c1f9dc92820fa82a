"""Load the UNMODIFIED reference for oracle validation / CPU baseline (TEST INFRASTRUCTURE ONLY).

Two levels:

* ``load_ref_kernels()`` -- the reference's compiled C++/OpenMP backend
  (pybind11 module ``_kernels`` built by oracle/build_ref.sh into oracle/_ref/).
  The ``.so`` travels to the GPU box, so this works there too.  Used by
  ``RefKernelSolver`` (a step driver that calls the reference kernels in the
  order of core/solver.py:2003-2126) for bench.py's ``--impl reference`` arm.

* ``load_reference_package()`` -- the reference's Python package imported from
  /root/reference/src (only in the build container).  Works around the shipped
  packaging faults documented in SURVEY.md F2/F3: the ``.so`` is loaded once and
  aliased under every sub-package name the stale relative imports look for, and
  ``h5py`` (absent here, only used by the HDF5 writer) is stubbed.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import sysconfig
import types
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
REF_ROOT = Path(os.environ.get("STRATA_REFERENCE", "/root/reference"))
_KERNELS = None


def ref_kernels_path() -> Path:
    return _HERE / "_ref" / ("_kernels" + sysconfig.get_config_var("EXT_SUFFIX"))


def have_ref_kernels() -> bool:
    return ref_kernels_path().exists()


def have_reference_package() -> bool:
    return (REF_ROOT / "src" / "strata_fdtd" / "__init__.py").exists() and have_ref_kernels()


def load_ref_kernels():
    global _KERNELS
    if _KERNELS is None:
        so = ref_kernels_path()
        if not so.exists():
            raise ImportError(f"{so} not built -- run oracle/build_ref.sh where /root/reference exists")
        spec = importlib.util.spec_from_file_location("_kernels", so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _KERNELS = mod
    return _KERNELS


def load_reference_package():
    """import strata_fdtd from /root/reference/src with the native backend actually loaded."""
    if "strata_fdtd" in sys.modules and getattr(sys.modules["strata_fdtd"], "_oracle_seeded", False):
        return sys.modules["strata_fdtd"]
    k = load_ref_kernels()
    try:
        import h5py  # noqa: F401
    except ImportError:
        sys.modules["h5py"] = types.ModuleType("h5py")
    for name in ("strata_fdtd._kernels", "strata_fdtd.core._kernels",
                 "strata_fdtd.boundaries._kernels", "strata_fdtd.io._kernels"):
        sys.modules[name] = k
    src = str(REF_ROOT / "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import strata_fdtd
    from strata_fdtd.core import solver as _s
    assert _s.has_native_kernels(), "reference native backend failed to load"
    strata_fdtd._oracle_seeded = True
    return strata_fdtd


# --------------------------------------------------------------------------- case -> reference solver
def build_reference_solver(case: dict):
    """Construct the reference's own FDTDSolver(backend='native') from a case dict.

    With materials, applies the 'fixed-native ADE' arrangement of SURVEY.md F4/F5:
    pole lists assigned wholesale (pybind def_readwrite copies make .append a no-op)
    and the reference's native kernels called in the documented order
    (core/solver.py:2015-2022, 2135-2193).
    """
    sf = load_reference_package()
    from strata_fdtd.core.solver import FDTDSolver, GaussianPulse
    from strata_fdtd.boundaries import PML
    kw = dict(c=case.get("c", 343.0), rho=case.get("rho", 1.2), courant=case.get("courant", 0.95),
              backend="native")
    nu = case.get("nonuniform")
    if nu is None:
        s = FDTDSolver(shape=tuple(case["shape"]), resolution=case["resolution"], **kw)
    else:
        g = sf.NonuniformGrid(x_coords=nu["x_coords"], y_coords=nu["y_coords"], z_coords=nu["z_coords"])
        s = FDTDSolver(grid=g, **kw)
    if case.get("geometry") is not None:
        s.set_geometry(np.asarray(case["geometry"], dtype=bool))
    for b in case.get("pml", []):
        axes = tuple(b.get("axes", ("x", "y", "z")))
        s.add_boundary(PML(depth=b.get("depth", 10), axis="all" if axes == ("x", "y", "z") else axes,
                           max_sigma=b.get("max_sigma"), order=b.get("order", 3)))
    for b in case.get("plane_bcs", []):
        from strata_fdtd.boundaries import ABCFirstOrder, RadiationImpedance
        if b["kind"] == "mur":
            s.add_boundary(ABCFirstOrder(axis=tuple(b.get("axes", ("x", "y", "z")))))
        else:
            s.add_boundary(RadiationImpedance(axis=b["axis"], side=b["side"],
                                              reflection_coeff=b.get("reflection_coeff"),
                                              pipe_radius=b.get("pipe_radius")))
    for src in case.get("sources", []):
        kind = src.get("kind", "point")
        if kind == "weighted":
            s.add_source(_reference_membrane(src["spec"]))
            continue
        pos = src["position"] if kind == "point" else {"axis": src["axis"], "index": src["index"]}
        s.add_source(GaussianPulse(position=pos, frequency=src["frequency"], bandwidth=src.get("bandwidth"),
                                   amplitude=src.get("amplitude", 1.0), source_type=kind))
    for name, pos in case.get("probes", []):
        s.add_probe(name, position=pos)
    for name, pos, *opt in case.get("mics", []):
        s.add_microphone(position=pos, name=name, **(opt[0] if opt else {}))
    if case.get("materials"):
        _install_fixed_native_ade(s, case)
    return s


def _reference_membrane(m: dict):
    """The reference's own membrane source object for a tests/cases.py MEMBRANE_SPEC entry."""
    from strata_fdtd.core.solver import CircularMembraneSource, GaussianPulse, RectangularMembraneSource
    wf = GaussianPulse(position=(0, 0, 0), frequency=m["frequency"], amplitude=m["amplitude"])
    if m["shape"] == "circular":
        return CircularMembraneSource(center=m["center"], radius=m["radius"], normal_axis=m["normal_axis"],
                                      waveform=wf, mode=tuple(m["mode"]), injection_type=m["injection_type"])
    return RectangularMembraneSource(center=m["center"], size=tuple(m["size"]), normal_axis=m["normal_axis"],
                                     waveform=wf, mode=tuple(m["mode"]), injection_type=m["injection_type"])


def _install_fixed_native_ade(s, case: dict):
    from strata_fdtd.materials.base import Pole, PoleType
    k = load_ref_kernels()
    for m in case["materials"]:
        poles = []
        for p in m["poles"]:
            if p["type"] == "debye":
                poles.append(Pole(pole_type=PoleType.DEBYE, delta_chi=p["delta_chi"], tau=p["tau"], target=p["target"]))
            else:
                poles.append(Pole(pole_type=PoleType.LORENTZ, delta_chi=p["delta_chi"], omega_0=p["omega_0"],
                                  gamma=p["gamma"], target=p["target"]))
        mat = _make_material(m, poles)
        s.register_material(mat, material_id=m["id"])
    mid = np.asarray(case["material_id"], dtype=np.uint8)
    for m in case["materials"]:
        s.set_material_region(mid == m["id"], material_id=m["id"])
    s._initialize_ade_fields()
    # wholesale assignment (F5)
    d = s._ade_native_data
    deb, lor = [], []
    di = li = 0
    rho_inf = [0.0] * (max(s._materials) + 1)
    K_inf = [0.0] * (max(s._materials) + 1)
    for mat_id, material in s._materials.items():
        rho_inf[mat_id] = float(material.rho_inf)
        K_inf[mat_id] = float(material.K_inf)
        for pole in material.poles:
            if pole.is_debye:
                dp = k.ADEDebyePole(); dp.material_id = mat_id; dp.target = 0 if pole.target == "density" else 1
                dp.field_index = di; co = k.DebyePoleCoeffs(); a, b = pole.fdtd_coefficients(s.dt)
                co.alpha = float(a); co.beta = float(b); dp.coeffs = co; deb.append(dp); di += 1
            else:
                lp = k.ADELorentzPole(); lp.material_id = mat_id; lp.target = 0 if pole.target == "density" else 1
                lp.field_index = li; co = k.LorentzPoleCoeffs(); a, b, dd = pole.fdtd_coefficients(s.dt)
                co.a = float(a); co.b = float(b); co.d = float(dd); lp.coeffs = co; lor.append(lp); li += 1
    d.debye_poles = deb
    d.lorentz_poles = lor
    d.rho_inf = rho_inf
    d.K_inf = K_inf
    d.n_debye_fields = di
    d.n_lorentz_fields = li
    s._ade_native_data = d
    assert len(s._ade_native_data.debye_poles) == len(deb)

    uniform = s._grid.is_uniform

    def step_native_with_ade():
        """Reference native kernels in the order of core/solver.py:2135-2193."""
        s._update_ade_density_poles()
        if uniform:
            k.update_velocity(s.p, s.vx, s.vy, s.vz, s._coeff_v)
        else:
            k.update_velocity_nonuniform(s.p, s.vx, s.vy, s.vz, s._native_grid_data, s._coeff_v_base)
        s._apply_ade_velocity_correction()
        if s._boundary_cells is not None:
            k.apply_rigid_boundaries(s.vx, s.vy, s.vz, s._boundary_cells)
        s._update_ade_modulus_poles()
        if uniform:
            k.update_pressure(s.p, s.vx, s.vy, s.vz, s.geometry, s._coeff_p)
        else:
            k.update_pressure_nonuniform(s.p, s.vx, s.vy, s.vz, s.geometry, s._native_grid_data, s._coeff_p_base)
        s._apply_ade_pressure_correction()
        s.p[~s.geometry] = 0.0

    s._step_native = step_native_with_ade
    s._step_native_nonuniform = step_native_with_ade


def _make_material(m: dict, poles: list):
    """A minimal AcousticMaterial with explicit rho_inf / K_inf / poles."""
    from strata_fdtd.materials.base import AcousticMaterial

    class _CaseMaterial(AcousticMaterial):
        def __init__(self):
            super().__init__(name=m.get("name", f"mat{m['id']}"))

        @property
        def rho_inf(self): return m["rho_inf"]

        @property
        def K_inf(self): return m["K_inf"]

        @property
        def poles(self): return poles

    return _CaseMaterial()


# --------------------------------------------------------------------------- kernel-level driver (travels to the GPU box)
class RefKernelSolver:
    """Drives the reference's compiled kernels in FDTDSolver.step() order without the Python package.

    Used as bench.py's ``--impl reference`` / ``cpu_baseline(kind="reference")`` arm on the GPU box,
    where /root/reference does not exist but oracle/_ref/_kernels*.so does.  Host-side set-up numbers
    (dt, coefficients, sigma) come from oracle.OracleSolver's restatement; every per-step array
    operation is the reference's own compiled code (core/solver.py:2079-2102, 2044-2047).
    """

    def __init__(self, case: dict):
        from . import oracle as _o
        self.k = load_ref_kernels()
        self.o = o = _o.OracleSolver(case)          # host numbers + state arrays
        assert o.uniform and not o.materials, "RefKernelSolver covers the uniform, ADE-free path"
        self.bc = self.k.precompute_boundary_cells(o.geometry) if o.rigid else None
        nx, ny, nz = o.shape
        self.pml = [self.k.initialize_pml(sp["sigma"][0], sp["sigma"][1], sp["sigma"][2], nx, ny, nz, o.dt)
                    for sp in o.sponges]

    def step(self):
        o, k = self.o, self.k
        k.update_velocity(o.p, o.vx, o.vy, o.vz, float(o.cv))
        if self.bc is not None:
            k.apply_rigid_boundaries(o.vx, o.vy, o.vz, self.bc)
        k.update_pressure(o.p, o.vx, o.vy, o.vz, o.geometry, float(o.cp))
        for d in self.pml:
            k.apply_pml_velocity(o.vx, o.vy, o.vz, d)
        for d in self.pml:
            k.apply_pml_pressure(o.p, d)
        o._inject()
        for name, (i, j, kk) in o.probes:
            o.probe_data[name].append(float(o.p[i, j, kk]))
        o._record_mics()
        o.step_count += 1
        o.time += o.dt
