"""Grid descriptions accepted by the b200 FDTDSolver.

Host-side mirror of the reference's ``UniformGrid`` / ``NonuniformGrid``
(/root/reference/src/strata_fdtd/core/grid.py:50-145, 148-540): same constructor
arguments, attributes and numerical definitions, so that spacing arrays -- and with
them every update coefficient -- are bit-identical (checked in tests/test_host_mirror.py).
The reference's own grid objects are accepted as well (duck typing on these attributes).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def _centres(n: int, h: float) -> np.ndarray:
    return np.arange(n) * h + h / 2


class UniformGrid:
    """Equal spacing ``resolution`` on every axis (reference grid.py:50-145)."""

    is_uniform = True

    def __init__(self, shape: tuple[int, int, int], resolution: float):
        self.shape = tuple(int(n) for n in shape)
        self.resolution = resolution
        self._spacing = [np.full(n, resolution, dtype=np.float64) for n in self.shape]
        self._coords = [_centres(n, resolution) for n in self.shape]

    dx = property(lambda self: self._spacing[0])
    dy = property(lambda self: self._spacing[1])
    dz = property(lambda self: self._spacing[2])
    x_coords = property(lambda self: self._coords[0])
    y_coords = property(lambda self: self._coords[1])
    z_coords = property(lambda self: self._coords[2])

    @property
    def min_spacing(self) -> float:
        return self.resolution

    @property
    def max_spacing(self) -> float:
        return self.resolution

    @property
    def num_cells(self) -> int:
        """Total cell count (the reference CLI reads this attribute: cli/progress.py:144)."""
        return int(np.prod(self.shape, dtype=np.int64))

    def physical_extent(self) -> tuple[float, float, float]:
        return tuple(n * self.resolution for n in self.shape)

    def __repr__(self):
        return f"UniformGrid(shape={self.shape}, resolution={self.resolution})"


class NonuniformGrid:
    """Per-axis cell-centre coordinates with variable spacing (reference grid.py:148-540)."""

    is_uniform = False

    def __init__(self, x_coords, y_coords, z_coords):
        self._coords = []
        for name, q in (("x_coords", x_coords), ("y_coords", y_coords), ("z_coords", z_coords)):
            q = np.asarray(q, dtype=np.float64).ravel()
            if q.size < 2:
                raise ValueError(f"{name} must have at least 2 points")
            if not np.all(np.diff(q) > 0):
                raise ValueError(f"{name} must be monotonically increasing")
            self._coords.append(q)
        self._spacing = [self._cell_sizes(q) for q in self._coords]
        self._shape = tuple(q.size for q in self._coords)

    @staticmethod
    def _cell_sizes(q: np.ndarray) -> np.ndarray:
        # interior: half the distance between the two neighbours; ends: one-sided (grid.py:219-237)
        sp = np.zeros(q.size, dtype=np.float64)
        for i in range(1, q.size - 1):
            sp[i] = (q[i + 1] - q[i - 1]) / 2
        sp[0] = q[1] - q[0]
        sp[-1] = q[-1] - q[-2]
        return sp

    shape = property(lambda self: self._shape)
    dx = property(lambda self: self._spacing[0])
    dy = property(lambda self: self._spacing[1])
    dz = property(lambda self: self._spacing[2])
    x_coords = property(lambda self: self._coords[0])
    y_coords = property(lambda self: self._coords[1])
    z_coords = property(lambda self: self._coords[2])

    @property
    def min_spacing(self) -> float:
        return float(min(np.min(s) for s in self._spacing))

    @property
    def max_spacing(self) -> float:
        return float(max(np.max(s) for s in self._spacing))

    @property
    def num_cells(self) -> int:
        return int(np.prod(self._shape, dtype=np.int64))

    @property
    def stretch_ratio(self) -> tuple[float, float, float]:
        return tuple(float(np.max(s) / np.min(s)) for s in self._spacing)

    def physical_extent(self) -> tuple[float, float, float]:
        return tuple(float(q[-1] - q[0] + s[-1] / 2 + s[0] / 2) for q, s in zip(self._coords, self._spacing))

    # ---- constructors ----------------------------------------------------------------
    @classmethod
    def from_stretch(cls, shape, base_resolution, stretch_x=1.0, stretch_y=1.0, stretch_z=1.0,
                     center_fine=True) -> "NonuniformGrid":
        """Geometric growth of the cell size per axis (grid.py:324-410)."""
        axes = [cls._stretched_axis(n, base_resolution, s, center_fine)
                for n, s in zip(shape, (stretch_x, stretch_y, stretch_z))]
        return cls(*axes)

    @staticmethod
    def _stretched_axis(n: int, h: float, ratio: float, center_fine: bool) -> np.ndarray:
        if ratio == 1.0:
            return np.arange(n, dtype=np.float64) * h + h / 2
        if not center_fine:
            sizes = h * (ratio ** np.arange(n, dtype=np.float64))
            return np.cumsum(sizes) - sizes / 2
        half = n // 2
        sizes = h * (ratio ** np.arange(half, dtype=np.float64))
        pos = np.cumsum(sizes) - sizes[0] / 2
        left = -pos[::-1] - sizes[::-1] / 2
        if n % 2 == 0:
            q = np.concatenate([left, pos + sizes / 2])
        else:
            right = pos[1:] + sizes[1:] / 2 if half > 0 else np.array([])
            q = np.concatenate([left, np.array([0.0]), right])
        return q - q[0] + h / 2

    @classmethod
    def from_regions(cls, x_regions: Sequence, y_regions: Sequence | None = None,
                     z_regions: Sequence | None = None) -> "NonuniformGrid":
        """Piecewise-constant resolution from (start, end, resolution) triples (grid.py:412-496)."""
        y_regions = x_regions if y_regions is None else y_regions
        z_regions = x_regions if z_regions is None else z_regions
        return cls(*(cls._region_axis(r) for r in (x_regions, y_regions, z_regions)))

    @staticmethod
    def _region_axis(regions: Sequence) -> np.ndarray:
        if not regions:
            raise ValueError("At least one region must be specified")
        pts: list[float] = []
        for n, (lo, hi, res) in enumerate(regions):
            if hi <= lo:
                raise ValueError(f"Region {n}: end ({hi}) must be > start ({lo})")
            if res <= 0:
                raise ValueError(f"Region {n}: resolution must be positive")
            if n > 0 and abs(lo - regions[n - 1][1]) > 1e-10:
                raise ValueError(f"Regions must be contiguous: region {n-1} ends at {regions[n-1][1]}, "
                                 f"region {n} starts at {lo}")
            count = max(1, int(round((hi - lo) / res)))
            step = (hi - lo) / count
            seg = lo + (np.arange(count) + 0.5) * step
            if pts and seg.size and abs(seg[0] - pts[-1]) < step / 2:
                seg = seg[1:]
            pts.extend(seg.tolist())
        return np.array(pts, dtype=np.float64)

    def get_spacing_arrays_for_stencil(self) -> dict[str, np.ndarray]:
        """fp32 inverse spacings: faces from np.diff(coords), cells from the cell sizes (grid.py:498-532)."""
        out = {}
        for a, q, s in zip("xyz", self._coords, self._spacing):
            out[f"inv_d{a}_face"] = (1.0 / np.diff(q)).astype(np.float32)
            out[f"inv_d{a}_cell"] = (1.0 / s).astype(np.float32)
        return out

    def __repr__(self):
        return (f"NonuniformGrid(shape={self.shape}, min_spacing={self.min_spacing:.4g}, "
                f"max_spacing={self.max_spacing:.4g})")
