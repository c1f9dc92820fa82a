"""Run scripts written for ``strata_fdtd`` unchanged on a box that only has this package.

``install_as_strata_fdtd()`` registers importable modules named ``strata_fdtd``, ``strata_fdtd.boundaries``
and ``strata_fdtd.materials`` whose names resolve to this package's mirrors, with ``FDTDSolver`` defaulting to
``backend="b200"`` ("auto", "native" and "python" requests are mapped to it as well, with a note, because no
other backend exists here).  When the real reference package is importable it is used instead and only gains
the extra backend (``shim.install_into_reference``), so geometry/SDF, the material library, DSP etc. keep working.

``python -m strata_fdtd_b200 script.py [args...]`` executes a script under that arrangement -- the stand-in
for ``fdtd-compute`` on the GPU box (the reference CLI itself needs ``grid.num_cells``, which the shim adds).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import warnings


def _reference_available() -> bool:
    if "strata_fdtd" in sys.modules and not getattr(sys.modules["strata_fdtd"], "_b200_alias", False):
        return True
    try:
        spec = importlib.util.find_spec("strata_fdtd")
    except (ImportError, ValueError):
        spec = None
    return spec is not None


class _Solid:
    """Protocol of the reference's SDF primitives (geometry/sdf.py:49-125): ``sdf(points)`` on an (N, 3) array of metres,
    negative inside; ``contains``; ``voxelize(grid)`` = cell centres with sdf <= 0.  Only the handful of shapes the
    reference's example scripts name live here; horns, transforms, smooth CSG etc. stay with the reference package."""

    def sdf(self, points):
        raise NotImplementedError

    @staticmethod
    def _points(points):
        import numpy as np
        points = np.asarray(points, dtype=np.float64)
        if points.ndim != 2 or points.shape[1] != 3:
            raise ValueError(f"points must be Nx3 array, got shape {points.shape}")
        return points

    def contains(self, points):
        return self.sdf(points) <= 0

    def voxelize(self, grid):
        import numpy as np
        X, Y, Z = np.meshgrid(grid.x_coords, grid.y_coords, grid.z_coords, indexing="ij")
        return (self.sdf(np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)) <= 0).reshape(grid.shape)


class Sphere(_Solid):
    """Ball: distance to the centre minus the radius (geometry/sdf.py:302-338)."""

    def __init__(self, center, radius):
        import numpy as np
        self.center = np.array(center, dtype=np.float64)
        self.radius = float(radius)
        if self.radius <= 0:
            raise ValueError(f"Sphere radius must be positive, got {self.radius}")

    def sdf(self, points):
        import numpy as np
        return np.linalg.norm(self._points(points) - self.center, axis=1) - self.radius

    @property
    def bounding_box(self):
        return self.center - self.radius, self.center + self.radius


class Box(_Solid):
    """Axis-aligned box from (center, size) or (min_corner, max_corner) (geometry/sdf.py:232-300): outside, the length of
    the positive part of |x - c| - size/2; inside, its largest (negative) component."""

    def __init__(self, center=None, size=None, min_corner=None, max_corner=None):
        import numpy as np
        if center is not None and size is not None:
            self.center, self.size = np.array(center, dtype=np.float64), np.array(size, dtype=np.float64)
        elif min_corner is not None and max_corner is not None:
            lo, hi = np.array(min_corner, dtype=np.float64), np.array(max_corner, dtype=np.float64)
            self.center, self.size = (lo + hi) / 2, hi - lo
        else:
            raise ValueError("Must provide either (center, size) or (min_corner, max_corner)")
        if np.any(self.size <= 0):
            raise ValueError(f"Box size must be positive, got {self.size}")

    def sdf(self, points):
        import numpy as np
        q = np.abs(self._points(points) - self.center) - self.size / 2
        return np.linalg.norm(np.maximum(q, 0), axis=1) + np.minimum(np.max(q, axis=1), 0)

    @property
    def bounding_box(self):
        return self.center - self.size / 2, self.center + self.size / 2


class Union(_Solid):
    """min over the children (geometry/sdf.py:549-610)."""

    def __init__(self, *children):
        self.children = children

    def sdf(self, points):
        import numpy as np
        out = self.children[0].sdf(points)
        for c in self.children[1:]:
            out = np.minimum(out, c.sdf(points))
        return out


class Intersection(_Solid):
    """max over the children (geometry/sdf.py:612-683)."""

    def __init__(self, *children):
        self.children = children

    def sdf(self, points):
        import numpy as np
        out = self.children[0].sdf(points)
        for c in self.children[1:]:
            out = np.maximum(out, c.sdf(points))
        return out


class Difference(_Solid):
    """base minus the others: max(base, -s1, -s2, ...) (geometry/sdf.py:685-735)."""

    def __init__(self, base, *subtracted):
        self.base, self.subtracted = base, subtracted

    def sdf(self, points):
        import numpy as np
        out = self.base.sdf(points)
        for c in self.subtracted:
            out = np.maximum(out, -c.sdf(points))
        return out

    @property
    def bounding_box(self):
        return self.base.bounding_box


def install_as_strata_fdtd(force_alias: bool = False):
    """Make ``import strata_fdtd`` work; returns the module that scripts will see."""
    if not force_alias and _reference_available():
        try:
            import strata_fdtd
            from .shim import install_into_reference
            os.environ.setdefault("STRATA_FDTD_BACKEND", "b200")
            return install_into_reference(strata_fdtd)
        except Exception as e:                       # e.g. h5py missing: fall back to the alias
            warnings.warn(f"reference package present but not importable ({e}); using the b200 mirrors")
            for k in [k for k in sys.modules if k == "strata_fdtd" or k.startswith("strata_fdtd.")]:
                del sys.modules[k]
    import strata_fdtd_b200 as sb
    from . import boundaries, materials, sources

    class FDTDSolver(sb.FDTDSolver):
        __doc__ = sb.FDTDSolver.__doc__
        _coerce_backend = True

        def __init__(self, *args, backend="auto", **kw):
            if backend not in ("b200", "auto"):
                warnings.warn(f"backend={backend!r} is not available in this installation; using 'b200'", stacklevel=2)
            super().__init__(*args, backend="b200", **kw)

    root = types.ModuleType("strata_fdtd")
    root.__doc__ = "strata_fdtd API served by strata_fdtd_b200 (backend='b200')"
    root._b200_alias = True
    root.__path__ = []                               # lets 'import strata_fdtd.materials' resolve via sys.modules
    names = dict(FDTDSolver=FDTDSolver, GaussianPulse=sb.GaussianPulse, Probe=sb.Probe, Microphone=sb.Microphone,
                 PML=sb.PML, RigidBoundary=sb.RigidBoundary, ABCFirstOrder=boundaries.ABCFirstOrder,
                 RadiationImpedance=boundaries.RadiationImpedance, UniformGrid=sb.UniformGrid,
                 NonuniformGrid=sb.NonuniformGrid, Pole=sb.Pole, PoleType=sb.PoleType, SimpleMaterial=sb.SimpleMaterial,
                 Sphere=Sphere, Box=Box, Union=Union, Intersection=Intersection, Difference=Difference,
                 POLAR_PATTERNS=sources.POLAR_PATTERNS, AudioFileWaveform=sb.AudioFileWaveform, MembraneSource=sb.MembraneSource,
                 CircularMembraneSource=sb.CircularMembraneSource, RectangularMembraneSource=sb.RectangularMembraneSource,
                 has_native_kernels=lambda: False, has_gpu_backend=lambda: True,
                 get_native_info=lambda: {"available": False, "version": None, "has_openmp": False, "num_threads": 1},
                 __version__=sb.__version__)
    for k, v in names.items():
        setattr(root, k, v)
    b = types.ModuleType("strata_fdtd.boundaries")
    for k in ("PML", "RigidBoundary", "ABCFirstOrder", "RadiationImpedance"):
        setattr(b, k, getattr(boundaries, k))
    m = types.ModuleType("strata_fdtd.materials")
    for k in ("Pole", "PoleType", "SimpleMaterial", "PoleMaterial", "WATER_20C"):
        setattr(m, k, getattr(materials, k))
    core = types.ModuleType("strata_fdtd.core")
    core.__path__ = []
    solver_mod = types.ModuleType("strata_fdtd.core.solver")
    for k in ("FDTDSolver", "GaussianPulse", "Probe", "Microphone", "POLAR_PATTERNS", "MembraneSource", "CircularMembraneSource",
              "RectangularMembraneSource", "has_native_kernels", "has_gpu_backend", "get_native_info"):
        setattr(solver_mod, k, names[k])
    grid_mod = types.ModuleType("strata_fdtd.core.grid")
    grid_mod.UniformGrid, grid_mod.NonuniformGrid = sb.UniformGrid, sb.NonuniformGrid
    core.solver, core.grid = solver_mod, grid_mod
    from . import io as sbio
    io_mod = types.ModuleType("strata_fdtd.io")                  # the result-file reader / writer interface (io/hdf5.py)
    io_mod.__path__ = []
    hdf5_mod = types.ModuleType("strata_fdtd.io.hdf5")
    for mod in (io_mod, hdf5_mod):
        mod.HDF5ResultWriter, mod.HDF5ResultReader = sbio.HDF5ResultWriter, sbio.HDF5ResultReader
    io_mod.hdf5 = hdf5_mod
    root.io = io_mod
    sys.modules.update({"strata_fdtd.io": io_mod, "strata_fdtd.io.hdf5": hdf5_mod})
    geo = types.ModuleType("strata_fdtd.geometry")               # the few shapes above; the toolkit itself stays upstream
    geo.__path__ = []
    for k in ("Sphere", "Box", "Union", "Intersection", "Difference"):
        setattr(geo, k, names[k])
    prim = types.ModuleType("strata_fdtd.geometry.primitives")
    prim.SPEED_OF_SOUND = 343.0                                  # geometry/primitives.py:19
    geo.primitives = prim
    root.boundaries, root.materials, root.core, root.geometry = b, m, core, geo
    sys.modules.update({"strata_fdtd": root, "strata_fdtd.boundaries": b, "strata_fdtd.materials": m,
                        "strata_fdtd.core": core, "strata_fdtd.core.solver": solver_mod,
                        "strata_fdtd.core.grid": grid_mod, "strata_fdtd.geometry": geo,
                        "strata_fdtd.geometry.primitives": prim})
    return root


def run_script(path: str, argv: list[str] | None = None) -> dict:
    """Execute ``path`` as __main__ with ``strata_fdtd`` importable; returns the script's globals."""
    import runpy
    install_as_strata_fdtd()
    old_argv = sys.argv
    sys.argv = [path] + list(argv or [])
    try:
        return runpy.run_path(path, run_name="__main__")
    finally:
        sys.argv = old_argv
