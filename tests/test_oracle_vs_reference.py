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
