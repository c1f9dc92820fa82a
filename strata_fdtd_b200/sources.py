"""Sources, probes and microphones of the b200 backend.

Host-side mirror of /root/reference/src/strata_fdtd/core/solver.py:152-207 (``GaussianPulse``),
:758-781 (``Probe``) and :797-1002 (``Microphone``, omnidirectional part).  Waveforms are
evaluated on the host in float64 exactly as the reference does and shipped to the device as
one table per chunk of steps; probe / microphone samples come back the same way.
The reference's own source objects (anything with ``source_type`` and ``waveform(t, dt)``,
including its membrane sources) are accepted too.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class GaussianPulse:
    """Gaussian-windowed sine: amp * exp(-(t-t0)^2 / 2 sigma^2) * sin(2 pi f (t-t0)),
    sigma = 1/(pi*bandwidth), t0 = 4 sigma (reference solver.py:188-207)."""

    position: tuple | dict
    frequency: float
    bandwidth: float | None = None
    amplitude: float = 1.0
    source_type: str = "point"

    def __post_init__(self):
        if self.bandwidth is None:
            self.bandwidth = 2.0 * self.frequency

    def waveform(self, t, dt):
        sigma = 1.0 / (np.pi * self.bandwidth)
        t0 = 4.0 * sigma
        envelope = np.exp(-((t - t0) ** 2) / (2 * sigma**2))
        carrier = np.sin(2 * np.pi * self.frequency * (t - t0))
        return self.amplitude * envelope * carrier


@dataclass
class Probe:
    """Pressure sample at one cell per step (reference solver.py:758-781)."""

    name: str
    position: tuple
    data: list = field(default_factory=list)

    def record(self, pressure: float) -> None:
        self.data.append(pressure)

    def get_data(self) -> np.ndarray:
        return np.array(self.data, dtype=np.float32)

    def clear(self) -> None:
        self.data.clear()


class Microphone:
    """Omnidirectional virtual microphone at a physical position, trilinear in p
    (reference solver.py:797-1002; weights as microphones.cpp:16-80).

    Directional patterns need the velocity field per step and are not on the device path yet.
    """

    def __init__(self, position, name=None, pattern="omni", direction=None, up=None):
        if pattern != "omni":
            raise NotImplementedError(
                "the b200 backend records omnidirectional microphones only (pattern='omni'); "
                "directional patterns are listed under 'next' in DESIGN.md")
        self.position = position
        self.name = name
        self.pattern = pattern
        self._pattern_name = "omni"
        self._data: list = []
        self._times: list = []
        self._grid_position = None
        self._solver_dt = None

    def is_directional(self) -> bool:
        return False

    def _initialize(self, solver) -> None:
        g = tuple(q / solver.dx for q in self.position)          # reference :941-943
        if not all(0 <= gq < n - 1 for gq, n in zip(g, solver.shape)):
            hi = tuple((n - 1) * solver.dx for n in solver.shape)
            raise ValueError(f"Microphone position {self.position} is outside simulation domain. "
                             f"Valid range: (0, 0, 0) to ({hi[0]:.4f}, {hi[1]:.4f}, {hi[2]:.4f})")
        self._grid_position = g
        self._solver_dt = solver.dt

    def get_waveform(self) -> np.ndarray:
        return np.array(self._data, dtype=np.float32)

    def get_time_axis(self) -> np.ndarray:
        return np.array(self._times, dtype=np.float64)

    def get_sample_rate(self) -> float:
        return 1.0 / self._solver_dt

    def clear(self) -> None:
        self._data.clear()
        self._times.clear()

    def __len__(self) -> int:
        return len(self._data)
