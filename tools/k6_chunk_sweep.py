import sys, json
from pathlib import Path
ROOT = Path("/root/repo") if Path("/root/repo/tools").exists() else Path(".").resolve()
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
from cases import c2_case
from resident_bench import time_chunk
from strata_fdtd_b200 import _lib
from util import build_b200_solver
for n in (200, 300):
    for wj in (4, 8, 2):
        for chunk in (4, 6, 8, 12, 16, 25):
            s = build_b200_solver(c2_case(n, steps=0))
            for k, v in {_lib.OPT_KERNEL: _lib.KERNEL_PIPELINE, _lib.OPT_ROWS_PER_THREAD: 1, _lib.OPT_WARPS_J: wj, _lib.OPT_CHUNK_I: chunk}.items():
                s.set_kernel_option(k, v)
            us = time_chunk(s, 100, reps=3)
            print(json.dumps({"n": n, "wj": wj, "chunk": chunk, "us": round(us, 2), "frac": round(32.0 * n ** 3 / (us * 1e-6) / 6540.8e9, 3)}), flush=True)
            s.close()
