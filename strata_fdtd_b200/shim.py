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
    _repair_cli_sandbox(module)
    return module


def _repair_cli_sandbox(module) -> None:
    """``fdtd-compute`` restricts what a simulation script may import by replacing ``__import__`` in the process-wide
    builtins while the script runs (cli/executor.py:57-83), so every lazy import made by *library* code in that window
    is refused too: on Python 3.12 ``Path.parent`` imports ``ntpath`` and the CLI dies before the script starts, with any
    backend.  The replacement applies the same allow-list to the import statements of the script itself only (its
    globals are the namespace the CLI gets back)."""
    try:
        from importlib import import_module
        executor = import_module(module.__name__ + ".cli.executor")
    except Exception:                                # CLI extras (click / rich) not installed: nothing to repair
        return
    import builtins
    import sys

    allowed = {"strata_fdtd", "numpy", "np", "scipy", "math", "pathlib"}

    def execute_simulation_script(script_path, script_content, verbose=False):
        namespace = {"__name__": "__main__", "__file__": str(script_path)}
        real_import = builtins.__import__

        def guarded_import(name, globals=None, locals=None, fromlist=(), level=0):
            if globals is namespace and level == 0 and name.split(".")[0] not in allowed:
                raise executor.RestrictedImportError(
                    f"Import of '{name}' is not allowed in simulation scripts. "
                    f"Allowed modules: {', '.join(sorted(allowed))}")
            return real_import(name, globals, locals, fromlist, level)

        namespace["__builtins__"] = {**vars(builtins), "__import__": guarded_import}     # the script's private copy
        script_dir = str(script_path.parent)
        sys.path.insert(0, script_dir)
        try:
            if verbose:
                print(f"Executing script: {script_path}")
            exec(compile(script_content, str(script_path), "exec"), namespace)
        finally:
            if script_dir in sys.path:
                sys.path.remove(script_dir)
        return namespace

    executor.execute_simulation_script = execute_simulation_script
    compute = sys.modules.get(module.__name__ + ".cli.compute")
    if compute is None:
        try:
            compute = import_module(module.__name__ + ".cli.compute")
        except Exception:
            return
    compute.execute_simulation_script = execute_simulation_script
