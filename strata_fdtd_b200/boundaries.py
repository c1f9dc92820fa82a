"""Boundary objects accepted by ``FDTDSolver.add_boundary`` on the b200 backend.

Host-side mirror of /root/reference/src/strata_fdtd/boundaries/_boundaries.py:75-417
(``RigidBoundary``, ``PML``).  The reference's "PML" is a sponge: per-axis multipliers
``expf(-sigma*dt)`` applied to v and p after every step; no auxiliary
recursion exists, so none is allocated here.  On this backend the multipliers are folded
into the fused step kernel, hence ``apply_velocity`` / ``apply_pressure`` do nothing.
"""
from __future__ import annotations

import numpy as np

from . import _lib

_AXES = ("x", "y", "z")


class RigidBoundary:
    """Perfectly reflecting walls -- implicit in the geometry mask (reference :75-101)."""

    def initialize(self, solver) -> None: ...
    def apply_velocity(self, solver) -> None: ...
    def apply_pressure(self, solver) -> None: ...
    def reset(self) -> None: ...


class PML:
    """Absorbing sponge layer of ``depth`` cells on the chosen axes (reference :104-417)."""

    def __init__(self, depth: int = 10, axis="all", max_sigma: float | None = None, order: int = 3):
        self.depth = depth
        self.order = order
        if axis == "all":
            self.axes = _AXES
        elif isinstance(axis, str):
            self.axes = (axis,)
        else:
            self.axes = tuple(axis)
        self._max_sigma = max_sigma
        self._initialized = False
        self._sigma_x = self._sigma_y = self._sigma_z = None
        self._decay = (None, None, None)
        self._solver = None

    # -- profiles ------------------------------------------------------------------------
    def _profile_uniform(self, n: int) -> np.ndarray:
        """sigma_max * (distance into the layer / depth)**order on both ends (reference :228-255)."""
        sig = np.zeros(n, dtype=np.float32)
        d = self.depth
        for m in range(d):
            sig[m] = self._max_sigma * (((d - m) / d) ** self.order)
        for m in range(n - d, n):
            sig[m] = self._max_sigma * (((m - (n - d - 1)) / d) ** self.order)
        return sig

    def _profile_nonuniform(self, coords: np.ndarray, sizes: np.ndarray) -> np.ndarray:
        """Same polynomial in physical distance; depth shrinks to n//4 if it would cover half (reference :257-315)."""
        n = len(coords)
        sig = np.zeros(n, dtype=np.float32)
        d = self.depth
        if d >= n // 2:
            d = max(1, n // 4)
        thick_lo, thick_hi = float(np.sum(sizes[:d])), float(np.sum(sizes[-d:]))
        face_lo = coords[d] - sizes[d] / 2 if d < n else coords[-1]
        face_hi = coords[n - d - 1] + sizes[n - d - 1] / 2 if n - d > 0 else coords[0]
        for m in range(d):
            frac = min(1.0, (face_lo - coords[m]) / thick_lo) if thick_lo > 0 else 0.0
            sig[m] = self._max_sigma * (frac ** self.order)
        for m in range(n - d, n):
            frac = min(1.0, (coords[m] - face_hi) / thick_hi) if thick_hi > 0 else 0.0
            sig[m] = self._max_sigma * (frac ** self.order)
        return sig

    def initialize(self, solver) -> None:
        self._solver = solver
        grid = solver.grid
        if self._max_sigma is None:                      # reference :182-187, R = 1e-6
            thickness = self.depth * grid.min_spacing
            self._max_sigma = -(self.order + 1) * solver.c * np.log(1e-6) / (2 * thickness)
        sig = []
        for a, (coords, sizes) in zip(_AXES, ((grid.x_coords, grid.dx), (grid.y_coords, grid.dy),
                                              (grid.z_coords, grid.dz))):
            if a not in self.axes:
                sig.append(None)
            elif grid.is_uniform:
                sig.append(self._profile_uniform(len(coords)))
            else:
                sig.append(self._profile_nonuniform(coords, sizes))
        self._sigma_x, self._sigma_y, self._sigma_z = sig
        self._decay = tuple(None if s is None else decay_table(s, solver.dt) for s in sig)
        self._initialized = True

    # fused into the step kernel on this backend
    def apply_velocity(self, solver) -> None: ...
    def apply_pressure(self, solver) -> None: ...
    def reset(self) -> None: ...

    def get_interior_slice(self):
        if self._solver is None:
            raise RuntimeError("PML not initialized")
        d = self.depth
        return tuple(slice(d, n - d) if a in self.axes else slice(None)
                     for a, n in zip(_AXES, self._solver.shape))

    @property
    def is_initialized(self) -> bool:
        return self._initialized


class ABCFirstOrder:
    """First-order Mur absorbing boundary on the chosen axes (reference _boundaries.py:420-526).

    On this backend the six plane updates run on the device after the sponges (k4_plane_op); the
    previous-plane state is kept there, so ``apply_pressure`` is a no-op."""

    def __init__(self, axis="all"):
        self.axes = _AXES if axis == "all" else ((axis,) if isinstance(axis, str) else tuple(axis))
        self._solver = None
        self._coeff = 0.0

    def initialize(self, solver) -> None:
        self._solver = solver
        self._coeff = (solver.c * solver.dt - solver.dx) / (solver.c * solver.dt + solver.dx)   # reference :461-463

    def apply_velocity(self, solver) -> None: ...
    def apply_pressure(self, solver) -> None: ...
    def reset(self) -> None: ...


class RadiationImpedance:
    """Partially reflecting open end on one face (reference _boundaries.py:529-793):
    p_b = R*p_i + (1-R)*(prev + mur*(p_i - p_b)); R constant or from the pipe radius (ka ~ 1 kHz, clamped 0.95)."""

    def __init__(self, axis, side, reflection_coeff: float | None = None, pipe_radius: float | None = None):
        self.axis, self.side = axis, side
        if reflection_coeff is not None:
            if not 0.0 <= reflection_coeff <= 1.0:
                raise ValueError("reflection_coeff must be in [0, 1]")
            self._reflection_coeff, self._frequency_dependent = reflection_coeff, False
        elif pipe_radius is not None:
            if pipe_radius <= 0:
                raise ValueError("pipe_radius must be positive")
            self._pipe_radius, self._frequency_dependent, self._reflection_coeff = pipe_radius, True, None
        else:
            raise ValueError("Must specify either reflection_coeff or pipe_radius")
        self._solver = None
        self._initialized = False
        self._coeff_a = self._coeff_b = 0.0

    def initialize(self, solver) -> None:
        self._solver = solver
        c, dt, dx = solver.c, solver.dt, solver.dx
        mur = (c * dt - dx) / (c * dt + dx)
        if not self._frequency_dependent:
            R = self._reflection_coeff
        else:                                            # reference :650-663
            ka = (2 * np.pi * 1000 / c) * self._pipe_radius
            z_ratio = (ka ** 2) / 4
            R = min(abs((z_ratio - 1) / (z_ratio + 1)), 0.95)
        self._coeff_a, self._coeff_b = (1 - R) * mur, R
        self._initialized = True

    def apply_velocity(self, solver) -> None: ...
    def apply_pressure(self, solver) -> None: ...
    def reset(self) -> None: ...

    @property
    def reflection_coefficient(self) -> float:
        return self._coeff_b

    @property
    def is_initialized(self) -> bool:
        return self._initialized


def plane_ops(boundary, solver):
    """[(axis, side, kind, mur, R, weak_r)] for a Mur / RadiationImpedance object (ours or the reference's),
    in the order the reference applies them; None if the object is not of that family."""
    name = type(boundary).__name__
    mur = (solver.c * solver.dt - solver.dx) / (solver.c * solver.dt + solver.dx)
    if name == "ABCFirstOrder":
        return [("xyz".index(a), side, 0, float(boundary._coeff), 0.0, 0)
                for a in "xyz" if a in boundary.axes for side in (0, 1)]
    if name == "RadiationImpedance":
        R = boundary._coeff_b
        weak = isinstance(R, float) and not isinstance(R, np.floating)     # Python float -> NumPy multiplies in fp32
        return [("xyz".index(boundary.axis), 0 if boundary.side == "low" else 1, 1, float(mur), float(R), int(weak))]
    return None


def decay_table(sigma: np.ndarray, dt) -> np.ndarray:
    """``expf(-sigma*float(dt))`` through the host libm, as pml.cpp:13-45 computes it."""
    sigma = np.ascontiguousarray(sigma, dtype=np.float32)
    out = np.empty_like(sigma)
    _lib.check(_lib.load().sb_sponge_decay(_lib.ptr(sigma), len(sigma), float(dt), _lib.ptr(out)))
    return out


def sponge_tables(boundary, solver):
    """(decay_x, decay_y, decay_z) for our PML or any object exposing _sigma_x/_y/_z (the reference's PML)."""
    if isinstance(boundary, PML):
        return boundary._decay
    sig = [getattr(boundary, "_sigma_" + a, None) for a in _AXES]
    if all(s is None for s in sig) and not hasattr(boundary, "_sigma_x"):
        return None
    return tuple(None if s is None else decay_table(s, solver.dt) for s in sig)
