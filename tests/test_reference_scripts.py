"""The reference's own example scripts and CLI on backend="b200" (BASELINE north_star: "example scripts and fdtd-compute
work unchanged"), and the HDF5 result schema.

* GPU: /root/reference/examples/{basic_pulse, material_sphere, multiple_probes, waveguide, organ_pipes, pzt_transducer,
  frequency_sweep}.py (ADE sphere, rigid geometry, a 16-element phased array, a user-defined chirp source class, 3127
  steps for the pipes), byte-identical copies of which travel in
  oracle/_ref/examples/ (put there by __graft_entry__.build(); never committed), run through
  ``python -m strata_fdtd_b200 <script>``; their probe traces and final fields must equal, bit for bit, what the
  UNMODIFIED reference produced with its native backend in the build container (tests/golden/script_*.npz, written
  by oracle/make_golden_scripts.py) -- metre-valued positions, ``duration=`` and all.  examples/nonuniform_grid.py
  fails upstream (a probe position outside the grid); it must fail the same way here.  sealed_subwoofer.py needs the
  reference's enclosure builder (geometry toolkit, out of scope) and is not run.
* CPU: our ResultWriter, given an h5py (tests/fake_h5py.py -- the real one is installed nowhere here), produces the
  tree the reference's HDF5ResultWriter produced for the same script, and the reference's HDF5ResultReader
  (io/hdf5.py:236-375) reads it back; ``fdtd-compute --dry-run`` gets past print_simulation_info.
"""
import hashlib
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import fake_h5py
from util import sha

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"
REF_EXAMPLES = ROOT / "oracle" / "_ref" / "examples"
SCRIPTS = ["basic_pulse", "material_sphere", "multiple_probes", "waveguide", "organ_pipes", "pzt_transducer", "frequency_sweep"]


def _tree(node, prefix=""):
    out = []
    for name, child in node.items():
        path = f"{prefix}/{name}"
        if isinstance(child, fake_h5py.Dataset):
            out.append(f"D {path} {tuple(child.shape)} {child.dtype} attrs={sorted(child.attrs)}")
        else:
            out.append(f"G {path} attrs={sorted(child.attrs)}")
            out += _tree(child, path)
    return out


def _script(name):
    path = REF_EXAMPLES / f"{name}.py"
    if not path.exists():
        pytest.skip(f"{path} not present (build() copies it where /root/reference exists)")
    g = np.load(GOLDEN / f"script_{name}.npz")
    assert hashlib.sha256(path.read_bytes()).hexdigest() == str(g["script_sha"]), "not the script the fixture was made from"
    return path, g


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCRIPTS)
def test_reference_script_verbatim_through_the_module_runner(name, tmp_path):
    """python -m strata_fdtd_b200 <reference script>: the file it writes holds the reference's traces."""
    path, g = _script(name)
    env = dict(os.environ, PYTHONPATH=str(ROOT) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run([sys.executable, "-m", "strata_fdtd_b200", str(path)], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert "Simulation complete" in res.stdout
    out = tmp_path / "results.h5"
    assert out.exists()
    from strata_fdtd_b200 import io as sbio
    assert not sbio.HAVE_H5PY, "with a real h5py the file is HDF5: read it with h5py here"
    z = np.load(out)
    attrs = json.loads(str(z["__attrs__"]))
    assert attrs["simulation@num_steps"] == int(g["steps"]) and attrs["metadata@backend"] == "b200"
    for key in g.files:
        if key.startswith("probe_") and not key.startswith("probe_pos_"):
            pname = key[6:]
            assert np.array_equal(z["probes/" + pname], g[key]), f"{name}: probe {pname} differs from the reference"
            assert attrs[f"probes/{pname}@position"] == [int(q) for q in g["probe_pos_" + pname]]


@pytest.mark.gpu
def test_reference_script_that_fails_upstream_fails_the_same_way(tmp_path):
    """examples/nonuniform_grid.py: the reference raises ValueError from add_probe (core/solver.py:1863-1884 converts
    metres with the minimum spacing); same exception, same message, before any step."""
    want = json.loads((GOLDEN / "script_failures.json").read_text())["nonuniform_grid"]
    path = REF_EXAMPLES / "nonuniform_grid.py"
    if not path.exists():
        pytest.skip(f"{path} not present")
    assert hashlib.sha256(path.read_bytes()).hexdigest() == want["script_sha"]
    env = dict(os.environ, PYTHONPATH=str(ROOT) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run([sys.executable, "-m", "strata_fdtd_b200", str(path)], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode != 0
    assert res.stderr.strip().splitlines()[-1] == f"{want['type']}: {want['message']}", res.stderr[-2000:]


@pytest.mark.gpu
def test_reference_script_unchanged_on_two_gpus(tmp_path):
    """python -m torch.distributed.run --nproc-per-node 2 -m strata_fdtd_b200 basic_pulse.py: the script's plain
    ``FDTDSolver(...)`` becomes the slab solver, rank 0 writes results.h5, and the trace is the reference's."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    path, g = _script("basic_pulse")
    env = dict(os.environ, PYTHONPATH=str(ROOT) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29547", "-m", "strata_fdtd_b200", str(path)], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.count("Simulation complete") == 1, "only rank 0 prints"
    z = np.load(tmp_path / "results.h5")
    attrs = json.loads(str(z["__attrs__"]))
    assert attrs["metadata@num_gpus"] == 2 and attrs["simulation@num_steps"] == int(g["steps"])
    assert np.array_equal(z["probes/downstream"], g["probe_downstream"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCRIPTS)
def test_reference_script_fields_traces_and_hdf5_tree(name, tmp_path, monkeypatch):
    """The same scripts in-process: final fields (SHA-256) and traces equal the reference's, and -- with an h5py
    available -- the result file has exactly the groups, datasets, shapes and attribute names the reference wrote."""
    path, g = _script(name)
    from strata_fdtd_b200 import compat
    from strata_fdtd_b200 import io as sbio
    monkeypatch.setattr(sbio, "h5py", fake_h5py)
    monkeypatch.setattr(sbio, "HAVE_H5PY", True)
    monkeypatch.chdir(tmp_path)
    ns = compat.run_script(str(path))
    s = ns["solver"]
    assert int(s.step_count) == int(g["steps"]) and float(s.dt) == float(g["dt"]) and float(s.time) == float(g["time"])
    for key in g.files:
        if key.startswith("probe_") and not key.startswith("probe_pos_"):
            assert np.array_equal(s.get_probe_data(key[6:])[key[6:]], g[key]), key
    for f in ("p", "vx", "vy", "vz"):
        assert sha(s.get_field(f)) == str(g["sha_" + f]), f"{name}: final {f} != reference"
    f = fake_h5py.File(tmp_path / "results.h5", "r")
    assert "\n".join(_tree(f)) == str(g["tree"])
    for key in g.files:
        if key.startswith("probe_") and not key.startswith("probe_pos_"):
            assert np.array_equal(f["probes/" + key[6:]][:], g[key])
    s.close()


def _fake_run(tmp_path, monkeypatch, n_blocks=5, block=40):
    """A solver that never touches a device, fed with synthetic traces through the writer's producer interface."""
    import strata_fdtd_b200 as sb
    from strata_fdtd_b200 import io as sbio
    monkeypatch.setattr(sbio, "h5py", fake_h5py)
    monkeypatch.setattr(sbio, "HAVE_H5PY", True)
    s = sb.FDTDSolver(shape=(12, 10, 8), resolution=1e-3)
    s.add_source(sb.GaussianPulse(position=(3, 5, 4), frequency=40e3))
    s.add_probe("a", (6, 5, 4)); s.add_probe("b", (0.002, 0.003, 0.004))
    g = np.ones(s.shape, dtype=bool); g[4:6, 2:4, 1:3] = False
    s.set_geometry(g)
    w = sbio.ResultWriter(tmp_path / "out.h5", s, script_content="print('hi')")
    rng = np.random.default_rng(3)
    blocks = [rng.standard_normal((block, 2)).astype(np.float32) for _ in range(n_blocks)]
    snaps = [rng.standard_normal(s.shape).astype(np.float32) for _ in range(3)]
    for q, b in enumerate(blocks):
        w.append_probe_block(["a", "b"], b)
        if q < 3:
            w.write_snapshot(snaps[q])
    s._step_count, s._time = n_blocks * block, n_blocks * block * s.dt
    w.finalize(runtime=1.5, backend="b200", num_threads=0)
    return s, g, np.concatenate(blocks), snaps


def test_result_writer_hdf5_branch_has_the_reference_schema(tmp_path, monkeypatch):
    s, g, traces, snaps = _fake_run(tmp_path, monkeypatch)
    f = fake_h5py.File(tmp_path / "out.h5", "r")
    want = str(np.load(GOLDEN / "script_basic_pulse.npz")["tree"]).splitlines()
    got = _tree(f)
    # same groups and attribute names as the reference's writer produced (shapes / probe names differ with the case)
    strip = lambda lines: sorted(q.split(" ")[0] + " " + q.split(" ")[1] + " " + q[q.index("attrs="):] for q in lines
                                 if "/probes/" not in q and "/materials/geometry" not in q and "/fields/pressure" not in q)
    ours = strip(got)
    theirs = strip(want)
    theirs[theirs.index(next(q for q in theirs if q.startswith("G /metadata")))] = \
        "G /metadata attrs=['backend', 'created_at', 'num_threads', 'script_content', 'script_hash', 'solver_version', 'total_runtime_seconds']"
    assert ours == sorted(theirs)
    assert f["probes/a"].dtype == np.float32 and f["probes/a"].maxshape == (None,) and f["probes/a"].compression == "gzip"
    assert np.array_equal(f["probes/a"][:], traces[:, 0]) and np.array_equal(f["probes/b"][:], traces[:, 1])
    assert list(f["probes/b"].attrs["position"]) == [2, 3, 4] and f["probes/a"].attrs["units"] == "Pa"
    p = f["fields/pressure"]
    assert p.shape == (3,) + s.shape and p.chunks == (1,) + s.shape and p.attrs["units"] == "Pa"
    assert all(np.array_equal(p[q], snaps[q]) for q in range(3))
    assert np.array_equal(f["materials/geometry"][:].astype(bool), g)
    assert f["metadata"].attrs["script_hash"] == hashlib.sha256(b"print('hi')").hexdigest()
    assert f["simulation"].attrs["num_steps"] == 200


def test_reference_reader_reads_what_our_writer_wrote(tmp_path, monkeypatch):
    """HDF5ResultReader (io/hdf5.py:236-375), imported from the reference itself, on a file written by io.ResultWriter."""
    from oracle import ref_loader as R
    if not R.have_reference_package():
        pytest.skip("needs /root/reference")
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    for k in [k for k in sys.modules if k == "strata_fdtd" or k.startswith("strata_fdtd.")]:
        monkeypatch.delitem(sys.modules, k)              # a copy imported earlier may hold the h5py stub
    R.load_reference_package()
    from strata_fdtd.io.hdf5 import HDF5ResultReader
    s, g, traces, snaps = _fake_run(tmp_path, monkeypatch)
    rd = HDF5ResultReader(tmp_path / "out.h5")
    meta = rd.get_metadata()
    assert list(meta["grid"]["shape"]) == list(s.shape) and meta["grid"]["is_uniform"] and meta["grid"]["resolution"] == s.dx
    assert meta["simulation"]["timestep"] == s.dt and meta["simulation"]["num_steps"] == 200 and meta["simulation"]["c"] == 343.0
    assert meta["metadata"]["backend"] == "b200" and meta["metadata"]["total_runtime_seconds"] == 1.5
    assert meta["sources"][0]["type"] == "point" and list(meta["sources"][0]["position"]) == [3, 5, 4]
    assert meta["sources"][0]["frequency"] == 40e3 and "bandwidth" in meta["sources"][0]
    assert rd.get_probe_names() == ["a", "b"] and list(meta["probes"]["a"]["position"]) == [6, 5, 4]
    assert np.array_equal(rd.load_probe("a"), traces[:, 0]) and np.array_equal(rd.load_probe("b"), traces[:, 1])
    assert rd.get_num_snapshots() == 3 and np.array_equal(rd.load_timestep(2), snaps[2])
    assert np.array_equal(rd.load_geometry(), g)
    with pytest.raises(KeyError):
        rd.load_probe("nope")
    rd.close()
    for k in [k for k in sys.modules if k == "strata_fdtd" or k.startswith("strata_fdtd.")]:
        monkeypatch.delitem(sys.modules, k, raising=False)   # do not leave a copy bound to the fake h5py behind


def test_npz_fallback_appends_in_linear_time(tmp_path):
    """Without h5py the tree goes into an .npz; blocks are kept in a list and joined once at close."""
    from strata_fdtd_b200 import io as sbio
    t = sbio._NpzTree(tmp_path / "x.npz")
    blocks = [np.full(7, q, dtype=np.float32) for q in range(50)]
    for b in blocks:
        t.append("probes/a", b)
    assert isinstance(t.data["probes/a"], list) and len(t.data["probes/a"]) == 50      # no re-concatenation per block
    t.close()
    z = np.load(tmp_path / "x.npz")
    assert np.array_equal(z["probes/a"], np.concatenate(blocks))


def test_fdtd_compute_dry_run_through_the_shim(tmp_path, monkeypatch):
    """The reference's CLI (cli/compute.py:46-207) with backend="b200" up to the point where the device would be
    created: script execution in its sandbox, validate_solver_object, print_simulation_info -- which reads
    grid.num_cells (cli/progress.py:195), missing from the reference's own grids and supplied by the shim."""
    from oracle import ref_loader as R
    if not R.have_reference_package():
        pytest.skip("needs /root/reference")
    for k in [k for k in sys.modules if k == "strata_fdtd" or k.startswith("strata_fdtd.")]:
        monkeypatch.delitem(sys.modules, k)
    sf = R.load_reference_package()
    import strata_fdtd_b200 as sb
    sb.install_into_reference(sf)
    monkeypatch.setenv("STRATA_FDTD_BACKEND", "b200")
    from click.testing import CliRunner
    from strata_fdtd.cli.compute import main
    script = tmp_path / "sim.py"
    script.write_text("from strata_fdtd import FDTDSolver, GaussianPulse\n"
                      "solver = FDTDSolver(shape=(100, 100, 100), resolution=1e-3)\n"
                      "solver.add_source(GaussianPulse(position=(0.05, 0.05, 0.05), frequency=40e3))\n"
                      "solver.add_probe('center', (0.05, 0.05, 0.05))\n"
                      "num_steps = 50\n")
    res = CliRunner().invoke(main, [str(script), "--dry-run", "-o", str(tmp_path / "o.h5")], standalone_mode=False)
    assert res.exception is None, res.output
    assert res.return_value == 0, res.output
    assert "1.0M cells" in res.output and "Dry run" in res.output and "50 steps" in res.output
    # the sandbox still refuses what it refused before -- for the script's own imports only
    bad = tmp_path / "bad.py"
    bad.write_text("import os\nsolver = None\n")
    res = CliRunner().invoke(main, [str(bad), "--dry-run"], standalone_mode=False)
    assert res.return_value == 1 and "Security Error" in res.output and "'os' is not allowed" in res.output
    for k in [k for k in sys.modules if k == "strata_fdtd" or k.startswith("strata_fdtd.")]:
        monkeypatch.delitem(sys.modules, k, raising=False)
