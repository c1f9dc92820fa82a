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


POLAR_PATTERNS = {"omni": 0.0, "subcardioid": 0.3, "cardioid": 0.5, "supercardioid": 0.63,
                  "hypercardioid": 0.75, "figure8": 1.0}           # S(theta) = (1-k) + k cos(theta)


def trilinear_tables(g, shape, clamp: bool = False):
    """8 corner indices and float64 weights of the reference's Python path (solver.py:962-1002), corner
    order (i0,j0,k0),(i1,j0,k0),(i0,j1,k0),(i1,j1,k0),(i0,j0,k1),...; ``clamp`` as in solver.py:1093-1097."""
    gx, gy, gz = g
    i0, j0, k0 = int(gx), int(gy), int(gz)
    fx, fy, fz = gx - i0, gy - j0, gz - k0
    idx, wts = [], []
    for c in range(8):
        bx, by, bz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        i, j, k = i0 + bx, j0 + by, k0 + bz
        if clamp:
            i, j, k = min(i, shape[0] - 1), min(j, shape[1] - 1), min(k, shape[2] - 1)
        idx.append((i * shape[1] + j) * shape[2] + k)
        wts.append((fx if bx else 1 - fx) * (fy if by else 1 - fy) * (fz if bz else 1 - fz))
    return idx, wts


class Microphone:
    """Virtual microphone at a physical position, trilinear in the fields (reference solver.py:797-1439).

    Omnidirectional microphones are gathered on the device with the native backend's fp32 weights
    (microphones.cpp:16-116).  Directional patterns (cardioid ... figure-8, or a callable gain(theta)) follow the
    reference's Python path (solver.py:1004-1173): the device gathers p and the three velocity components with
    that path's weights, the host applies (1-k) p + k rho c (v . d) in the same mixed precision."""

    def __init__(self, position, name=None, pattern="omni", direction=None, up=None):
        self.position = position
        self.name = name
        if isinstance(pattern, str):
            if pattern not in POLAR_PATTERNS:
                raise ValueError(f"Unknown pattern '{pattern}'. Valid patterns: {list(POLAR_PATTERNS.keys())}")
            self._pattern_name, self._pattern_k, self._custom_pattern = pattern, POLAR_PATTERNS[pattern], None
        elif callable(pattern):
            self._pattern_name, self._pattern_k, self._custom_pattern = "custom", None, pattern
        else:
            raise TypeError(f"pattern must be str or callable, got {type(pattern).__name__}")
        self.pattern = pattern
        d = np.array((1.0, 0.0, 0.0) if direction is None else direction, dtype=np.float64)
        if np.linalg.norm(d) < 1e-10:
            raise ValueError("direction vector cannot be zero")
        self._direction = d / np.linalg.norm(d)
        u = np.array((0.0, 0.0, 1.0) if up is None else up, dtype=np.float64)
        if np.linalg.norm(u) < 1e-10:
            raise ValueError("up vector cannot be zero")
        self._up = u / np.linalg.norm(u)
        if abs(np.dot(self._direction, self._up)) > 0.999:
            raise ValueError("direction and up vectors cannot be parallel")
        self._data: list = []
        self._times: list = []
        self._grid_position = None
        self._solver_dt = self._solver_rho = self._solver_c = None

    @property
    def direction(self):
        return tuple(self._direction)

    def is_directional(self) -> bool:
        return self._pattern_name != "omni"

    def _initialize(self, solver) -> None:
        g = tuple(q / solver.dx for q in self.position)          # reference :941-943
        shape = getattr(solver, "global_shape", solver.shape)    # positions are global also on a slab
        if not all(0 <= gq < n - 1 for gq, n in zip(g, shape)):
            hi = tuple((n - 1) * solver.dx for n in shape)
            raise ValueError(f"Microphone position {self.position} is outside simulation domain. "
                             f"Valid range: (0, 0, 0) to ({hi[0]:.4f}, {hi[1]:.4f}, {hi[2]:.4f})")
        self._grid_position = g
        self._solver_dt, self._solver_rho, self._solver_c = solver.dt, solver.rho, solver.c

    # ---- Python-path sampling tables (used whenever any microphone of the solver is directional) ----
    def _gather_tables(self, shape):
        """[(field, idx8, w8 float32)]: p, then vx, vy, vz for a directional microphone.  The reference multiplies
        a float64 weight with an fp32 sample, which NumPy evaluates in fp32 -> the weight is rounded first."""
        out = [(0, *trilinear_tables(self._grid_position, shape))]
        if self.is_directional():
            for axis in range(3):                                 # staggered positions, solver.py:1073-1090
                g = list(self._grid_position)
                g[axis] -= 0.5
                g = [max(0.0, min(q, n - 1.001)) for q, n in zip(g, shape)]
                out.append((1 + axis, *trilinear_tables(g, shape, clamp=True)))
        return [(f, np.array(i, dtype=np.int64), np.array(w, dtype=np.float64).astype(np.float32)) for f, i, w in out]

    def _combine(self, pressure, vel):
        """Directional output from fp32 samples (solver.py:1102-1173); arrays in, float64 array out."""
        if not self.is_directional():
            return pressure
        vx, vy, vz = vel
        d = self._direction
        v_dot_d = vx * d[0] + vy * d[1] + vz * d[2]               # fp32 * np.float64 -> float64
        if self._custom_pattern is not None:
            out = np.empty(len(pressure), dtype=np.float64)
            for q in range(len(pressure)):
                v_mag = np.sqrt(vx[q] ** 2 + vy[q] ** 2 + vz[q] ** 2)
                if v_mag > 1e-20:
                    theta = np.arccos(np.clip(v_dot_d[q] / v_mag, -1.0, 1.0))
                else:
                    theta = 0.0
                out[q] = pressure[q] * self._custom_pattern(theta)
            return out
        k = self._pattern_k
        Z = self._solver_rho * self._solver_c
        return (1 - k) * pressure + k * Z * v_dot_d               # fp32 term + float64 term

    def record(self, pressure_field, time: float, vx=None, vy=None, vz=None) -> None:
        """One sample taken on the host from whole-field arrays (the reference's per-step Python path, solver.py:1004-1058).
        ``run()`` never calls this -- it gathers on the device -- but scripts and the reference's own tests do."""
        if self._grid_position is None:
            raise RuntimeError("Microphone not initialized. Add to solver first.")

        def sample(field, g, clamp):
            idx, wts = trilinear_tables(g, field.shape, clamp=clamp)
            flat, acc = field.reshape(-1), 0.0
            for q, w in zip(idx, wts):
                acc += w * flat[q]
            return acc
        pressure = sample(np.asarray(pressure_field), self._grid_position, False)
        if self.is_directional():
            if vx is None or vy is None or vz is None:
                raise ValueError("Velocity fields (vx, vy, vz) are required for directional microphones")
            vel = []
            for axis, f in enumerate((vx, vy, vz)):               # staggered positions, clamped (solver.py:1073-1097)
                f = np.asarray(f)
                g = list(self._grid_position)
                g[axis] -= 0.5
                vel.append(sample(f, [max(0.0, min(q, n - 1.001)) for q, n in zip(g, f.shape)], True))
            pressure = self._combine(np.array([pressure]), [np.array([q]) for q in vel])[0]
        self._data.append(pressure)
        self._times.append(time)

    def get_waveform(self, weighting=None) -> np.ndarray:
        if weighting not in ("Z", None):
            raise NotImplementedError("frequency weighting is post-processing; use strata_fdtd.analysis on the raw waveform")
        return np.array(self._data, dtype=np.float32)

    def to_wav(self, filepath: str, sample_rate: int = 44100, normalize: bool = True, bit_depth: int = 16) -> None:
        """Mono WAV of the recording, linearly resampled from 1/dt to ``sample_rate`` when they differ by more than 1 Hz,
        optionally scaled to 0.95 of full scale (solver.py:1337-1418)."""
        import wave
        if len(self._data) == 0:
            raise RuntimeError("No data recorded. Run simulation first.")
        if bit_depth not in (16, 32):
            raise ValueError("bit_depth must be 16 or 32")
        w = self.get_waveform()
        if self._solver_dt is not None and abs(1.0 / self._solver_dt - sample_rate) > 1.0:
            n_out = int(len(w) * sample_rate / (1.0 / self._solver_dt))
            if n_out != len(w):
                w = np.interp(np.linspace(0, len(w) - 1, n_out), np.arange(len(w)), w).astype(np.float32)
        if normalize:
            peak = np.max(np.abs(w))
            if peak > 0:
                w = w / peak * 0.95
        ints = (w * (32767 if bit_depth == 16 else 2147483647)).astype(np.int16 if bit_depth == 16 else np.int32)
        with wave.open(filepath, "wb") as f:
            f.setnchannels(1)
            f.setsampwidth(bit_depth // 8)
            f.setframerate(sample_rate)
            f.writeframes(ints.tobytes())

    def get_time_axis(self) -> np.ndarray:
        return np.array(self._times, dtype=np.float64)

    def get_sample_rate(self) -> float:
        if self._solver_dt is None:
            raise RuntimeError("Microphone not initialized. Add to solver first.")
        return 1.0 / self._solver_dt

    def clear(self) -> None:
        self._data.clear()
        self._times.clear()

    def __len__(self) -> int:
        return len(self._data)

    def __repr__(self) -> str:
        who = f"'{self.name}'" if self.name else "unnamed"
        aim = f", direction={self.direction}" if self.is_directional() else ""
        return f"Microphone({who}, position={self.position}, pattern='{self._pattern_name}'{aim}, samples={len(self._data)})"


def combine_corner_samples(mics, gathers, corners: dict, times) -> None:
    """Finish the microphones of a decomposed run: ``corners[(mic, gather, corner)]`` are the raw fp32 samples the
    owning slabs recorded; each gather is summed exactly as microphones.cpp:82-116 / solver.py:1027-1034 do --
    ``sum = 0; sum += w[c] * f[c]`` for c = 0..7 in fp32 -- and handed to the microphone's pattern."""
    if len(times) == 0:
        return
    for mi, mic in enumerate(mics):
        cols = []
        for g, (_f, _idx8, w8) in enumerate(gathers[mi]):
            acc = np.zeros(len(times), dtype=np.float32)
            for c in range(8):
                acc = acc + np.float32(w8[c]) * corners[(mi, g, c)]
            cols.append(acc)
        mic._data.extend(np.asarray(mic._combine(cols[0], cols[1:])).tolist())
        mic._times.extend(np.asarray(times).tolist())
