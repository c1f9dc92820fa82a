"""Oracle vs the UNMODIFIED reference, run live (only where /root/reference exists)."""
import numpy as np
import pytest

from cases import make_cases
from oracle import oracle as O
from oracle import ref_loader as R

pytestmark = pytest.mark.skipif(not R.have_reference_package(),
                                reason="reference sources / oracle/_ref not available on this box")
CASES = make_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_bit_exact_vs_reference_native(name):
    case = CASES[name]
    o = O.OracleSolver(case)
    r = R.build_reference_solver(case)
    assert r.using_native
    assert float(o.dt) == float(r.dt)
    n = min(case["steps"], 120)
    for _ in range(n):
        o.step(); r.step()
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(getattr(o, f), getattr(r, f)), f
    for pname, _ in o.probes:
        assert np.array_equal(o.probe_array(pname), r.get_probe_data(pname)[pname])
    for mname, *_ in o.mics:
        assert np.array_equal(o.mic_array(mname), np.array(r.microphones[mname]._data, dtype=np.float32))


def test_ref_kernel_driver_matches_reference_solver():
    """RefKernelSolver (what bench.py --impl reference times) == reference FDTDSolver.step()."""
    case = CASES["block_pml"]
    d = R.RefKernelSolver(case)
    r = R.build_reference_solver(case)
    for _ in range(60):
        d.step(); r.step()
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(getattr(d.o, f), getattr(r, f)), f


@pytest.mark.skipif(not R.have_reference_package(), reason="reference sources not on this box")
def test_compiled_reference_passes_the_references_own_pinned_tests():
    """the reference holds no golden files -- its pins for this path are live tests (native vs NumPy backend,
    microphone kernels, solver physics, nonuniform grids).  They must pass on the build of the reference that the oracle is
    checked against and that bench.py times as the CPU baseline (oracle/_ref, compiled by oracle/build_ref from the sources
    where they lie)."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    ref_tests = Path("/root/reference/tests")
    files = [str(ref_tests / f) for f in ("test_native_extension.py", "test_microphone_native.py", "test_fdtd.py", "test_grid.py")]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(root / "tests"), str(root), os.environ.get("PYTHONPATH", "")]))
    res = subprocess.run([sys.executable, "-m", "pytest", *files, "-p", "ref_native_plugin", "-q", "--no-header", "-p", "no:cacheprovider",
                          "--rootdir", str(root / "tests"), "-k", "not throughput"],          # (those two need pytest-benchmark)
                         cwd=root / "tests", env=env, capture_output=True, text=True, timeout=900)
    tail = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
    assert res.returncode == 0 and " failed" not in tail and "error" not in tail, res.stdout[-3000:] + res.stderr[-2000:]
    assert int(tail.split(" passed")[0].split()[-1]) >= 118, tail
