"""Membrane sources: a plane of cells driven with a modal weight per cell (reference core/solver.py:210-755).

Mirrors of the reference's ``MembraneSource`` / ``CircularMembraneSource`` / ``RectangularMembraneSource`` for boxes where the
reference package is absent -- same constructor arguments, attributes, messages and, cell for cell, the same weights (the CPU
tests compare them with the reference live, and the reference's own membrane tests run against them).  The solver only needs
``source_type == "membrane"``, ``injection_type``, ``normal_axis``, ``waveform`` and ``get_injection_weights(grid)``, so the
reference's own objects are accepted just as well (solver._build_source_table).

Everything is evaluated on the membrane's plane only: the two in-plane offsets from the centre are outer "sums" of two
coordinate vectors, not slices of a whole-grid meshgrid; the whole-grid arrays the reference API promises are assembled from
that plane.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field

import numpy as np

_AXIS = {"x": 0, "y": 1, "z": 2}
# zeros alpha_mn of the Bessel functions J_m for the common drum modes (n-th positive zero of J_m)
_DRUM_ZEROS = {(0, 1): 2.4048255576957727, (0, 2): 5.5200781102863115, (1, 1): 3.8317059702075125,
               (1, 2): 7.0155866698156190, (2, 1): 5.1356223018406826, (2, 2): 8.4172441403998649}


def _unit_peak(shape):
    """|shape| scaled so that its largest entry is 1 (all zeros if there is nothing above 1e-10)."""
    shape = np.abs(shape)
    top = np.max(shape) if shape.size else 0.0
    return shape / top if top > 1e-10 else np.zeros_like(shape)


@dataclass
class MembraneSource:
    """Base class: where the plane lies and how whole-grid masks / weights are put together; the outline (``_inside``) and
    the modal pattern (``mode_shape``) come from the subclasses."""

    center: tuple
    normal_axis: str
    waveform: object                      # anything with .waveform(t, dt)
    mode: tuple = (0, 1)
    injection_type: str = "pressure"      # or "velocity": the face velocity along normal_axis
    source_type: str = field(default="membrane", init=False)
    _cached_weights: object = field(default=None, repr=False, init=False)
    _cached_mask: object = field(default=None, repr=False, init=False)

    # ---- geometry of the plane ---------------------------------------------------------------------------
    def _in_plane(self, grid):
        """(normal axis, index of the grid plane nearest to the centre, offsets A[:, None] and B[None, :] from the centre
        along the two remaining axes in x, y, z order)."""
        n = _AXIS[self.normal_axis]
        coords = (np.asarray(grid.x_coords), np.asarray(grid.y_coords), np.asarray(grid.z_coords))
        first, second = [a for a in range(3) if a != n]
        plane = int(np.argmin(np.abs(coords[n] - self.center[n])))
        return n, plane, (coords[first] - self.center[first])[:, None], (coords[second] - self.center[second])[None, :]

    @staticmethod
    def _whole_grid(grid, n, plane, values, dtype):
        out = np.zeros(grid.shape, dtype=dtype)
        sel = [slice(None)] * 3
        sel[n] = plane
        out[tuple(sel)] = values
        return out

    def _along_normal(self, grid, n, plane2d):
        """The plane's values repeated along the normal axis (the reference's coordinate arrays do not depend on it)."""
        return np.ascontiguousarray(np.broadcast_to(np.expand_dims(plane2d, n), grid.shape))

    def _grid_to_membrane_coords(self, grid):
        """Distance from the centre line and angle around it for every cell, and the plane index (reference :296-340)."""
        n, plane, a, b = self._in_plane(grid)
        return self._along_normal(grid, n, np.sqrt(a ** 2 + b ** 2)), self._along_normal(grid, n, np.arctan2(b + 0 * a, a + 0 * b)), plane

    def _check_grid_alignment(self, grid) -> None:
        """Warn when the centre lies more than half a cell away from the plane it is snapped to (reference :342-372)."""
        n = _AXIS[self.normal_axis]
        coords = (grid.x_coords, grid.y_coords, grid.z_coords)[n]
        plane = int(np.argmin(np.abs(np.asarray(coords) - self.center[n])))
        offset = abs(coords[plane] - self.center[n])
        sizes = (grid.dx, grid.dy, grid.dz)[n]
        spacing = sizes[plane] if hasattr(sizes, "__getitem__") else grid.min_spacing
        if offset > 0.5 * spacing:
            warnings.warn(f"MembraneSource center {self.center} is {offset:.4f}m ({offset / spacing:.1f} cells) off-grid in "
                          f"{self.normal_axis}-direction. Source will be placed at nearest grid plane.", stacklevel=3)

    # ---- what subclasses provide -----------------------------------------------------------------------------
    def mode_shape(self, r, theta):
        raise NotImplementedError("Subclasses must implement mode_shape()")

    def get_injection_mask(self, grid):
        raise NotImplementedError("Subclasses must implement get_injection_mask()")

    def get_injection_weights(self, grid):
        raise NotImplementedError("Subclasses must implement get_injection_weights()")


@dataclass(kw_only=True)
class CircularMembraneSource(MembraneSource):
    """Clamped circular membrane (drum head, cone): |J_m(alpha_mn r / a) cos(m theta)| inside the radius, scaled to a peak of
    1 over the cells it covers (reference :375-552)."""

    radius: float
    _BESSEL_ZEROS: dict = field(default_factory=lambda: dict(_DRUM_ZEROS), repr=False, init=False)
    _alpha: float = field(default=0.0, repr=False, init=False)

    def __post_init__(self) -> None:
        m, n = self.mode
        if m < 0:
            raise ValueError(f"Azimuthal mode m must be non-negative, got {m}")
        if n < 1:
            raise ValueError(f"Radial mode n must be positive, got {n}")
        if (m, n) in self._BESSEL_ZEROS:
            alpha = self._BESSEL_ZEROS[(m, n)]
        else:
            from scipy.special import jn_zeros
            alpha = jn_zeros(m, n)[-1]
        object.__setattr__(self, "_alpha", alpha)

    def mode_shape(self, r, theta):
        from scipy.special import j0, jv
        m = self.mode[0]
        rho = r / self.radius
        pattern = j0(self._alpha * rho) if m == 0 else jv(m, self._alpha * rho) * np.cos(m * theta)
        return _unit_peak(np.where(rho <= 1.0, pattern, 0.0))

    def _covered(self, grid):
        n, plane, a, b = self._in_plane(grid)
        r = np.sqrt(a ** 2 + b ** 2)
        return n, plane, a, b, r, r <= self.radius

    def get_injection_mask(self, grid):
        n, plane, _a, _b, _r, inside = self._covered(grid)
        return self._whole_grid(grid, n, plane, inside, bool)

    def get_injection_weights(self, grid):
        self._check_grid_alignment(grid)
        n, plane, a, b, r, inside = self._covered(grid)
        w = np.zeros(inside.shape, dtype=np.float64)
        w[inside] = self.mode_shape(r[inside], np.arctan2(b + 0 * a, a + 0 * b)[inside])
        return self._whole_grid(grid, n, plane, w, np.float64)


@dataclass(kw_only=True)
class RectangularMembraneSource(MembraneSource):
    """Clamped rectangular membrane (panel, planar driver): |sin(m pi (u / Lx + 1/2)) sin(n pi (v / Ly + 1/2))| inside the
    rectangle, scaled to a peak of 1 over the cells it covers; (1, 1) peaks at the centre (reference :555-755)."""

    size: tuple
    mode: tuple = (1, 1)

    def __post_init__(self) -> None:
        m, n = self.mode
        if m < 1:
            raise ValueError(f"Mode index m must be >= 1, got {m}")
        if n < 1:
            raise ValueError(f"Mode index n must be >= 1, got {n}")

    def _grid_to_rectangular_coords(self, grid):
        n, plane, a, b = self._in_plane(grid)
        return self._along_normal(grid, n, a + 0 * b), self._along_normal(grid, n, b + 0 * a), plane

    def mode_shape(self, u, v):
        m, n = self.mode
        lx, ly = self.size
        pattern = np.sin(m * np.pi * (u / lx + 0.5)) * np.sin(n * np.pi * (v / ly + 0.5))
        return _unit_peak(np.where((np.abs(u) <= lx / 2) & (np.abs(v) <= ly / 2), pattern, 0.0))

    def _covered(self, grid):
        n, plane, a, b = self._in_plane(grid)
        u, v = a + 0 * b, b + 0 * a
        return n, plane, u, v, (np.abs(u) <= self.size[0] / 2) & (np.abs(v) <= self.size[1] / 2)

    def get_injection_mask(self, grid):
        n, plane, _u, _v, inside = self._covered(grid)
        return self._whole_grid(grid, n, plane, inside, bool)

    def get_injection_weights(self, grid):
        self._check_grid_alignment(grid)
        n, plane, u, v, inside = self._covered(grid)
        w = np.zeros(inside.shape, dtype=np.float64)
        w[inside] = self.mode_shape(u[inside], v[inside])
        return self._whole_grid(grid, n, plane, w, np.float64)
