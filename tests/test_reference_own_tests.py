"""The reference's OWN solver and grid test files, unmodified, against backend="b200".

/root/reference/tests/test_fdtd.py, test_grid.py, test_microphone.py, test_microphone_directional.py, test_membrane_source.py,
test_circular_membrane.py, test_rectangular_membrane.py, test_waveforms.py, test_hdf5_output.py and test_fdtd_output.py (273 tests -- 257 pass, 11 skip themselves for want of the reference's C++ kernels, 5 are listed below: wave speed, symmetry, rigid walls, PML absorption, energy
conservation in a closed pipe over 2000 steps, probes, Gaussian pulses, radiation impedance, nonuniform grids, trilinear and
directional microphones, WAV export, Bessel / sinusoidal membrane modes and their injection, audio-file waveforms, the result-file writer and reader, ...) travel to
the GPU box as byte-identical copies in oracle/_ref/tests/ (put there by __graft_entry__.build(); oracle/_ref is git-ignored,
reference files never enter this repository).  They are collected with tests/ref_alias_plugin.py, which makes
``import strata_fdtd`` resolve to this package (compat.install_as_strata_fdtd) -- the situation of a user who switches a
script over -- and must pass except for the handful listed below, each of which asserts something about the reference's own
backend bookkeeping that cannot hold for another backend.
"""
import os
import subprocess
import sys
import xml.etree.ElementTree as ET
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF_TESTS = ROOT / "oracle" / "_ref" / "tests"

FILES = ["test_grid.py", "test_fdtd.py", "test_microphone.py", "test_microphone_directional.py", "test_membrane_source.py",
         "test_circular_membrane.py", "test_rectangular_membrane.py", "test_waveforms.py", "test_hdf5_output.py", "test_fdtd_output.py"]

# test id -> why it cannot pass on any backend but the reference's own
EXPECTED_DIFFERENCES = {
    "test_grid.py::TestFDTDWithNonuniformGrid::test_native_backend_with_nonuniform":
        "expects ImportError for backend='native' where the C++ kernels are missing; here the request is served by b200 with a warning",
    "test_fdtd.py::TestGPUBackendSelection::test_backend_native_raises_if_unavailable": "same",
    "test_fdtd.py::TestGPUBackendSelection::test_backend_gpu_raises_if_unavailable":
        "expects the reference's PyTorch backend and its 'limited feature support' warning",
    "test_fdtd.py::TestGPUBackendSelection::test_backend_python_forces_python":
        "expects the NumPy backend (using_gpu False); there is no CPU path here by contract",
    "test_fdtd_output.py::test_compression_ratio":
        "compares file sizes with and without h5py's gzip filter; the in-memory h5py stand-in of this image stores datasets as they come",
}


@pytest.mark.gpu
def test_reference_solver_and_grid_tests_pass_on_b200(tmp_path):
    if not all((REF_TESTS / f).exists() for f in FILES):
        pytest.skip(f"{REF_TESTS} not present (build() copies the files where /root/reference exists)")
    xml = tmp_path / "ref.xml"
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(ROOT / "tests"), str(ROOT), os.environ.get("PYTHONPATH", "")]))
    res = subprocess.run([sys.executable, "-m", "pytest", *FILES, "-p", "ref_alias_plugin", "-q", "--no-header",
                          "-p", "no:cacheprovider", "--tb=short", f"--junitxml={xml}"], cwd=REF_TESTS, env=env,
                         capture_output=True, text=True, timeout=1500)
    assert xml.exists(), res.stdout[-3000:] + res.stderr[-3000:]
    outcome = {}
    for case in ET.parse(xml).getroot().iter("testcase"):
        cls = case.get("classname", "").split(".")
        tid = f"{cls[0]}.py::" + "::".join(cls[1:] + [case.get("name")])
        outcome[tid] = "failed" if case.find("failure") is not None or case.find("error") is not None else \
                       "skipped" if case.find("skipped") is not None else "passed"
    failed = {t for t, o in outcome.items() if o == "failed"}
    passed = {t for t, o in outcome.items() if o == "passed"}
    unexpected = failed - set(EXPECTED_DIFFERENCES)
    assert not unexpected, f"reference tests failing on b200: {sorted(unexpected)}\n" + res.stdout[-6000:]
    assert len(passed) >= 257, f"only {len(passed)} of the reference's tests passed: {res.stdout[-2000:]}"
    fixed = set(EXPECTED_DIFFERENCES) & passed
    assert not fixed, f"listed as expected differences but passing: {sorted(fixed)}"


def test_reference_waveform_tests_pass_on_the_mirror():
    """The one of those files that needs no device (test_waveforms.py: loading, trimming, resampling, looping of audio-file
    sources) also runs in the CPU suite, from the reference tree or its copy."""
    src = next((d for d in (REF_TESTS, Path("/root/reference/tests")) if (d / "test_waveforms.py").exists()), None)
    if src is None:
        pytest.skip("the reference's tests are not on this box")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(ROOT / "tests"), str(ROOT), os.environ.get("PYTHONPATH", "")]))
    res = subprocess.run([sys.executable, "-m", "pytest", str(src / "test_waveforms.py"), "-p", "ref_alias_plugin", "-q", "--no-header",
                          "-p", "no:cacheprovider", "--rootdir", str(ROOT / "tests")], cwd=ROOT / "tests", env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "24 passed" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
