"""Register backend="b200" with an importable reference package (``strata_fdtd``).

The reference dispatches on a constructor keyword (core/solver.py:1515, 1536-1574) and on
``step()`` (:2029-2040).  ``install_into_reference()`` wraps ``strata_fdtd.FDTDSolver`` so that
``backend="b200"`` (or env ``STRATA_FDTD_BACKEND=b200`` with backend="auto") constructs this
package's solver instead; every other backend value reaches the reference untouched.  It also
gives the reference grids the ``num_cells`` attribute its CLI reads (cli/progress.py:144, F10).
INTEGRATION.md shows the equivalent three-line patch a maintainer would apply upstream.
"""
from __future__ import annotations

import os


def install_into_reference(module=None):
    if module is None:
        import strata_fdtd as module
    from .solver import FDTDSolver as B200Solver
    ref_cls = module.FDTDSolver
    if getattr(ref_cls, "_b200_dispatch", False):
        return module

    class FDTDSolver(ref_cls):                        # noqa: N801 - keeps the public name
        _b200_dispatch = True

        def __new__(cls, *args, **kwargs):
            backend = kwargs.get("backend", "auto")
            if backend == "auto" and os.environ.get("STRATA_FDTD_BACKEND", "").lower() == "b200":
                backend = "b200"
            if backend == "b200":
                kwargs["backend"] = "b200"
                return B200Solver(*args, **kwargs)
            return super().__new__(cls)

    FDTDSolver.__name__ = ref_cls.__name__
    FDTDSolver.__qualname__ = ref_cls.__qualname__
    FDTDSolver.__doc__ = ref_cls.__doc__
    module.FDTDSolver = FDTDSolver
    core = getattr(module, "core", None)
    if core is not None and hasattr(core, "solver"):
        core.solver.FDTDSolver = FDTDSolver
    for gname in ("UniformGrid", "NonuniformGrid"):
        g = getattr(module, gname, None)
        if g is not None and not hasattr(g, "num_cells"):
            g.num_cells = property(lambda self: int(self.shape[0]) * int(self.shape[1]) * int(self.shape[2]))
    return module
