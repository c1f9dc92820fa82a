"""strata_fdtd_b200 -- B200-native backend for strata-fdtd's time-stepping hot path.

``FDTDSolver(..., backend="b200")`` keeps the reference solver's surface (grids, PML,
materials/ADE, add_source, add_probe, run) and executes the step on one B200 through the
C ABI in ``include/strata_b200.h``.  ``install_into_reference()`` registers the backend with
an importable reference package so that ``strata_fdtd.FDTDSolver(backend="b200")`` works too.
"""
from ._lib import B200BackendError, build  # noqa: F401
from .boundaries import PML, RigidBoundary  # noqa: F401
from .grid import NonuniformGrid, UniformGrid  # noqa: F401
from .materials import Pole, PoleMaterial, PoleType, SimpleMaterial  # noqa: F401
from .solver import FDTDSolver  # noqa: F401
from .sources import POLAR_PATTERNS, GaussianPulse, Microphone, Probe  # noqa: F401
from .membranes import CircularMembraneSource, MembraneSource, RectangularMembraneSource  # noqa: F401
from .waveforms import AudioFileWaveform  # noqa: F401
from .shim import install_into_reference  # noqa: F401
from . import io, workloads  # noqa: F401,E402

__version__ = "0.1.0"
__all__ = ["FDTDSolver", "UniformGrid", "NonuniformGrid", "PML", "RigidBoundary", "GaussianPulse", "Probe",
           "Microphone", "POLAR_PATTERNS", "AudioFileWaveform", "MembraneSource", "CircularMembraneSource", "RectangularMembraneSource",
           "Pole", "PoleType", "SimpleMaterial", "PoleMaterial", "B200BackendError", "build", "install_into_reference"]
