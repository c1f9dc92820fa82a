"""Pin the CPU oracle (oracle/oracle_step.c + oracle/oracle.py) against the committed
fixtures generated from the UNMODIFIED reference (oracle/make_golden.py).  Bit-exact."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

from cases import c1_case, make_cases
from oracle import oracle as O

GOLDEN = Path(__file__).parent / "golden"
CASES = make_cases()


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check_against_golden(solver_fields, probes, mics, g):
    for f in ("p", "vx", "vy", "vz"):
        arr = solver_fields[f]
        assert sha(arr) == str(g["sha_" + f]), f"final {f} differs from the reference"
        if "final_" + f in g:
            assert np.array_equal(arr, g["final_" + f])
        else:
            assert np.array_equal(arr[::3, ::3, ::3], g["sample_" + f])
    for name, trace in probes.items():
        assert np.array_equal(trace, g["probe_" + name]), f"probe {name}"
    for name, trace in mics.items():
        assert np.array_equal(trace, g["mic_" + name]), f"mic {name}"


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    case = CASES[name]
    g = np.load(GOLDEN / f"{name}.npz")
    o = O.OracleSolver(case)
    # host-side numbers: dt, coefficients, sponge profiles, indices -- exact
    assert float(o.dt) == float(g["dt"])
    assert np.float32(o.cv) == g["cv"] and np.float32(o.cp) == g["cp"]
    for bi, sp in enumerate(o.sponges):
        assert float(sp["max_sigma"]) == float(g[f"pml{bi}_max_sigma"])
        for a, ax in enumerate("xyz"):
            if sp["sigma"][a] is None:
                assert f"pml{bi}_sigma_{ax}" not in g
            else:
                assert np.array_equal(sp["sigma"][a], g[f"pml{bi}_sigma_{ax}"])
                assert np.array_equal(sp["decay"][a], g[f"pml{bi}_decay_{ax}"])
    if not o.uniform:
        for a, ax in enumerate("xyz"):
            assert np.array_equal(o.inv_face[a], g[f"sp_inv_d{ax}_face"])
            assert np.array_equal(o.inv_cell[a], g[f"sp_inv_d{ax}_cell"])
    ny, nz = o.shape[1], o.shape[2]
    for pname, (i, j, k) in o.probes:
        assert (i * ny + j) * nz + k == int(g["probe_idx_" + pname])
    for si, s in enumerate(o.sources):
        if s.get("kind", "point") == "point":
            i, j, k = s["position"]
            assert (i * ny + j) * nz + k == int(g[f"source_idx_{si}"])
    o.run_steps(case["steps"])
    if o.mics and "mic_flat_indices" in g:
        assert np.array_equal(o._mic_tables[0], g["mic_flat_indices"])
        assert np.array_equal(o._mic_tables[1], g["mic_weights"])
    check_against_golden({f: getattr(o, f) for f in ("p", "vx", "vy", "vz")},
                         {n: o.probe_array(n) for n, _ in o.probes},
                         {m[0]: o.mic_array(m[0]) for m in o.mics}, g)


@pytest.mark.slow
def test_oracle_matches_golden_c1_full():
    """BASELINE config 1 (100^3, PML 10, 1 kHz, 1000 steps) -- probe trace + field digests."""
    g = np.load(GOLDEN / "c1_100cubed_1000.npz")
    o = O.OracleSolver(c1_case(1000))
    o.run_steps(1000)
    check_against_golden({f: getattr(o, f) for f in ("p", "vx", "vy", "vz")},
                         {"probe": o.probe_array("probe")}, {}, g)


def test_ade_changes_the_answer():
    """Guard against a vacuous ADE fixture (the reference's own ADE tests pass vacuously)."""
    case = dict(CASES["ade_sphere"])
    with_mat = O.OracleSolver(case)
    plain = dict(case); plain.pop("materials"); plain.pop("material_id")
    without = O.OracleSolver(plain)
    with_mat.run_steps(150); without.run_steps(150)
    a, b = with_mat.probe_array("behind"), without.probe_array("behind")
    assert np.abs(a - b).max() > 1e-3 * np.abs(b).max()


def test_oracle_membrane_sources_match_reference_fixture():
    """Pressure- and velocity-injecting membrane sources (solver.py:2389-2412) vs the reference's output."""
    from util import load_membrane_case
    case, g = load_membrane_case()
    o = O.OracleSolver(case)
    assert float(o.dt) == float(g["dt"])
    o.run_steps(case["steps"])
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(getattr(o, f), g["final_" + f]), f
    for n, _ in o.probes:
        assert np.array_equal(o.probe_array(n), g["probe_" + n])
    assert np.abs(g["probe_a"]).max() > 0 and np.abs(g["probe_b"]).max() > 0
