"""Frequency-dependent (ADE) material description for the b200 backend.

Host-side mirror of /root/reference/src/strata_fdtd/materials/base.py:38-213 (``PoleType``,
``Pole``) and :215-291, 509-560 (``AcousticMaterial`` / ``SimpleMaterial``), reduced to what
the time-stepping path consumes: rho_inf, K_inf and the pole list with its update
coefficients.  Any object exposing those attributes -- in particular every material of the
reference's library -- can be registered instead.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum


class PoleType(Enum):
    DEBYE = "debye"
    LORENTZ = "lorentz"


@dataclass
class Pole:
    pole_type: PoleType
    delta_chi: float
    target: str                      # "density" | "modulus"
    tau: float | None = None         # Debye relaxation time [s]
    omega_0: float | None = None     # Lorentz resonance [rad/s]
    gamma: float | None = None       # Lorentz damping [rad/s]

    def __post_init__(self):
        if self.pole_type == PoleType.DEBYE:
            if self.tau is None:
                raise ValueError("Debye poles require tau parameter")
            if self.tau <= 0:
                raise ValueError("tau must be positive")
        else:
            if self.omega_0 is None or self.gamma is None:
                raise ValueError("Lorentz poles require omega_0 and gamma parameters")
            if self.omega_0 <= 0:
                raise ValueError("omega_0 must be positive")
            if self.gamma < 0:
                raise ValueError("gamma must be non-negative")
        if self.target not in ("density", "modulus"):
            raise ValueError("target must be 'density' or 'modulus'")

    @property
    def is_debye(self) -> bool:
        return self.pole_type == PoleType.DEBYE

    @property
    def is_lorentz(self) -> bool:
        return self.pole_type == PoleType.LORENTZ

    def fdtd_coefficients(self, dt):
        """Debye: J' = alpha J + beta f.  Lorentz: J' = a J + b J_prev + d f  (base.py:157-188)."""
        if self.is_debye:
            ratio = self.tau / dt
            norm = 1 + ratio
            return (ratio / norm, self.delta_chi / norm)
        w0, g = self.omega_0, self.gamma
        dt2 = dt * dt
        norm = 1 + g * dt / 2
        return ((2 - w0**2 * dt2) / norm, -(1 - g * dt / 2) / norm, self.delta_chi * w0**2 * dt2 / norm)


@dataclass
class SimpleMaterial:
    """Constant rho / c plus an optional pole list (base.py:509-560)."""

    name: str = "unnamed"
    _rho: float = field(default=1.2, repr=False)
    _c: float = field(default=343.0, repr=False)
    _poles: list = field(default_factory=list, repr=False)

    @property
    def rho_inf(self) -> float:
        return self._rho

    @property
    def K_inf(self) -> float:
        return self._rho * self._c**2

    @property
    def c_inf(self) -> float:
        return self._c

    @property
    def poles(self) -> list:
        return self._poles


WATER_20C = SimpleMaterial(name="water_20C", _rho=998.2, _c=1482.0)       # materials/library.py:52-56 (a constant scripts import)


@dataclass
class PoleMaterial:
    """Material given directly by rho_inf, K_inf and poles."""

    name: str
    rho_inf: float
    K_inf: float
    poles: list = field(default_factory=list)
