import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from cases import make_cases
from strata_fdtd_b200 import _lib
from test_multi_gpu import _group_from_case
from util import build_b200_solver
CASES = make_cases()
for name in ("ade_sphere", "ade_two_materials_nonuniform"):
    for graph in (1, 0):
        for steps in (16, 48, 33):
            case = CASES[name]
            one = build_b200_solver(case); one.run(steps=steps)
            grp = _group_from_case(case, 2, {_lib.OPT_USE_GRAPH: graph}, halo="p2p")
            t0 = time.time()
            try:
                grp.run(steps)
                ok = all(np.array_equal(grp.get_field(f), one.get_field(f)) for f in ("p", "vx", "vy", "vz"))
                msg = f"equal={ok}"
            except Exception as e:
                msg = "ERR " + str(e)[:80]
            print(name, "graph", graph, "steps", steps, f"{time.time()-t0:.2f}s", msg,
                  [s.device_stats()["kernels_launched"] for s in grp.slabs], flush=True)
            grp.close(); one.close()
