#!/usr/bin/env python
"""Wall time of the reference's example scripts (oracle/_ref/examples, copied there by build()) run unchanged on
backend="b200": whole script, and the run() call alone -- what a user switching backends sees at these small sizes."""
import contextlib
import io
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from strata_fdtd_b200 import compat  # noqa: E402
import strata_fdtd_b200 as sb  # noqa: E402

names = sys.argv[1:] or ["basic_pulse", "material_sphere", "multiple_probes", "waveguide", "organ_pipes", "pzt_transducer", "frequency_sweep"]
for rep in range(2):                      # the second round shows what the process-wide autotune cache saves
    for name in names:
        path = ROOT / "oracle" / "_ref" / "examples" / f"{name}.py"
        t_run = [0.0]
        orig = sb.FDTDSolver.run

        def timed(self, *a, _orig=orig, **k):
            import torch
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = _orig(self, *a, **k)
            torch.cuda.synchronize()
            t_run[0] += time.perf_counter() - t0
            return out
        sb.FDTDSolver.run = timed
        with tempfile.TemporaryDirectory() as tmp:
            cwd = os.getcwd(); os.chdir(tmp)
            try:
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    ns = compat.run_script(str(path))
                total = time.perf_counter() - t0
            finally:
                os.chdir(cwd); sb.FDTDSolver.run = orig
        s = ns["solver"]
        cells = s.shape[0] * s.shape[1] * s.shape[2]
        print(f"pass {rep} {name:16s} {s.shape} {s.step_count:5d} steps: script {total * 1e3:8.1f} ms, run() {t_run[0] * 1e3:8.1f} ms "
              f"= {t_run[0] / s.step_count * 1e6:7.1f} us/step ({cells * s.step_count / t_run[0] / 1e9:6.1f} Gcell/s), "
              f"kernel variant {s.device_stats()['kernel_variant']}", flush=True)
        s.close()
