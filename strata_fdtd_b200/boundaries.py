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
