"""Parity of the CUDA path (through FDTDSolver -> C ABI) with the CPU oracle and with the
committed fixtures generated from the unmodified reference.  Bit-exact: every comparison is
np.array_equal on fp32 fields and traces (the stated tolerance of 1e-5 relative is slack)."""
from pathlib import Path

import numpy as np
import pytest

from cases import c1_case, c2_case, c3_case, make_cases
from oracle import oracle as O
from strata_fdtd_b200 import _lib
from util import assert_same_as_oracle, build_b200_solver, sha

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
CASES = make_cases()
VARIANTS = {
    "naive": {_lib.OPT_KERNEL: _lib.KERNEL_NAIVE},
    "march_r1": {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_ROWS_PER_THREAD: 1},
    "march_r2": {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_ROWS_PER_THREAD: 2},
    "march_r1_chunk5_fused": {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_ROWS_PER_THREAD: 1, _lib.OPT_CHUNK_I: 5,
                              _lib.OPT_WARPS_J: 2, _lib.OPT_FUSE_K3: 2},
    "march_r2_graph": {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_ROWS_PER_THREAD: 2, _lib.OPT_USE_GRAPH: 1},
    "march_r2_separate_k3": {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_ROWS_PER_THREAD: 2, _lib.OPT_FUSE_K3: 0},
    "march_r1_flat": {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_ROWS_PER_THREAD: 1, _lib.OPT_PLANE_MAP: 2, _lib.OPT_FUSE_K3: 0,
                      _lib.OPT_CHUNK_I: 7},
    "march_r2_flat_fused": {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_ROWS_PER_THREAD: 2, _lib.OPT_PLANE_MAP: 2, _lib.OPT_FUSE_K3: 2,
                            _lib.OPT_WARPS_J: 2},
    "march_r2_strips": {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_ROWS_PER_THREAD: 2, _lib.OPT_PLANE_MAP: 1, _lib.OPT_FUSE_K3: 0},
    "march_r1_nograph_wk2": {_lib.OPT_KERNEL: _lib.KERNEL_MARCH, _lib.OPT_ROWS_PER_THREAD: 1, _lib.OPT_USE_GRAPH: 0,
                             _lib.OPT_WARPS_K: 2, _lib.OPT_WARPS_J: 2},
}


def _with_options(s, opts):
    for k, v in opts.items():
        s.set_kernel_option(k, v)
    return s


@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("name", sorted(CASES))
def test_matches_oracle_and_golden(name, variant):
    case = CASES[name]
    g = np.load(GOLDEN / f"{name}.npz")
    s = _with_options(build_b200_solver(case, chunk_steps=37), VARIANTS[variant])
    assert float(s.dt) == float(g["dt"])
    s.run(steps=case["steps"])
    # fixtures from the reference itself
    for f in ("p", "vx", "vy", "vz"):
        assert sha(s.get_field(f)) == str(g["sha_" + f]), f"{name}/{variant}: final {f} != reference"
    for pname in s._probes:
        assert np.array_equal(s.get_probe_data(pname)[pname], g["probe_" + pname])
    for mname, mic in s.microphones.items():
        assert np.array_equal(mic.get_waveform(), g["mic_" + mname])
    # and the oracle run live on the same inputs
    o = O.OracleSolver(case)
    o.run_steps(case["steps"])
    assert_same_as_oracle(s, o, f"{name}/{variant}")
    assert s.kernel_launches() > 0
    s.close()


def test_sponge_tables_and_indices_match_golden():
    case = CASES["partial_pml_plane"]
    g = np.load(GOLDEN / "partial_pml_plane.npz")
    s = build_b200_solver(case)
    for bi, b in enumerate(s._boundaries):
        if not hasattr(b, "_max_sigma"):
            continue
        assert float(b._max_sigma) == float(g[f"pml{bi}_max_sigma"])
        for a, sig, dec in zip("xyz", (b._sigma_x, b._sigma_y, b._sigma_z), b._decay):
            if sig is None:
                assert f"pml{bi}_sigma_{a}" not in g
            else:
                assert np.array_equal(sig, g[f"pml{bi}_sigma_{a}"])
                assert np.array_equal(dec, g[f"pml{bi}_decay_{a}"])


def test_step_by_step_with_host_pokes():
    """solver.p is a live view: writes between steps are honoured (reference tests/test_fdtd.py:228)."""
    case = dict(shape=(20, 18, 22), resolution=1e-3, steps=0, pml=[dict(depth=4)])
    s = build_b200_solver(case)
    o = O.OracleSolver(case)
    s.p[10, 9, 11] = 1.0
    o.p[10, 9, 11] = 1.0
    for n in range(25):
        s.step(); o.step()
        if n == 7:
            s.vx[3, 4, 5] = 0.25; o.vx[3, 4, 5] = 0.25
            s.p[2, 2, 2] += 0.5;  o.p[2, 2, 2] += 0.5
    assert_same_as_oracle(s, o, "pokes")
    # last-face velocities are never updated but are damped
    s.vx[19, 5, 5] = 1.0; o.vx[19, 5, 5] = 1.0
    s.step(); o.step()
    assert_same_as_oracle(s, o, "last face")


def test_c1_full_config_against_reference_fixture():
    """BASELINE config 1: 100^3, PML 10, 1 kHz, 1 probe, 1000 steps -- vs the reference's own output."""
    g = np.load(GOLDEN / "c1_100cubed_1000.npz")
    s = build_b200_solver(c1_case(1000))
    s.run(steps=1000)
    assert np.array_equal(s.get_probe_data("probe")["probe"], g["probe_probe"])
    for f in ("p", "vx", "vy", "vz"):
        assert sha(s.get_field(f)) == str(g["sha_" + f])


def test_c2_200cubed_vs_oracle_with_geometry():
    case = c2_case(200, steps=120, with_geometry=True)
    s = build_b200_solver(case)
    o = O.OracleSolver(case)
    s.run(steps=120); o.run_steps(120)
    assert_same_as_oracle(s, o, "c2")


def test_c2_full_config_1000_steps_vs_oracle():
    """BASELINE config 2 as worded (200^3, PML 10, 1000 steps; SURVEY 8d parity gate): final fields and the probe
    trace after 1000 steps equal the oracle bit for bit (the stated tolerance, 1e-5 relative, is slack)."""
    case = c2_case(200, steps=1000)
    s = build_b200_solver(case)
    o = O.OracleSolver(case)
    s.run(steps=1000); o.run_steps(1000)
    assert_same_as_oracle(s, o, "c2/1000")
    ref = np.abs(o.probe_array("probe")).max()
    assert ref > 0 and np.abs(s.get_probe_data("probe")["probe"] - o.probe_array("probe")).max() <= 1e-5 * ref
    s.close()


def test_large_grid_march_equals_naive():
    """Size-independent property at a size the oracle cannot reach quickly: the marching kernel and the
    one-thread-per-cell kernel are two independent implementations and must agree bit-for-bit."""
    case = c2_case(320, steps=60, with_geometry=True)
    a = _with_options(build_b200_solver(case), VARIANTS["march_r2"])
    b = _with_options(build_b200_solver(case), VARIANTS["naive"])
    a.run(steps=60); b.run(steps=60)
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(a.get_field(f), b.get_field(f)), f
    assert np.array_equal(a.get_probe_data("probe")["probe"], b.get_probe_data("probe")["probe"])
    assert np.abs(a.get_field("p")).max() > 0


def test_energy_and_reset():
    case = dict(shape=(24, 24, 24), resolution=1e-3, steps=0)
    s = build_b200_solver(case)
    s.p[12, 12, 12] = 1.0
    e0 = s.compute_energy()
    p = np.zeros((24, 24, 24), np.float32); p[12, 12, 12] = 1.0
    want = 0.5 * float((p.astype(np.float64) ** 2).sum()) / (1.2 * 343.0**2) * (1e-3) ** 3
    assert abs(e0 - want) <= 1e-12 * want
    s.run(steps=40, track_energy=True, energy_sample_interval=10)
    hist = s.get_energy_history()
    assert [h[0] for h in hist] == [0, 10, 20, 30, 40]
    f = {k: s.get_field(k).astype(np.float64) for k in ("p", "vx", "vy", "vz")}   # solver.py:2697-2706 in float64
    want = (0.5 * (f["p"] ** 2).sum() / (1.2 * 343.0**2) + 0.5 * 1.2 * (f["vx"]**2 + f["vy"]**2 + f["vz"]**2).sum()) * 1e-9
    assert abs(hist[-1][2] - want) <= 1e-6 * want
    assert 0.2 < hist[-1][2] / hist[1][2] < 5.0                          # closed rigid box: energy stays bounded
    s.reset()
    assert s.step_count == 0 and s.time == 0.0 and not np.any(s.get_field("p"))


def test_no_cuda_fallback_message():
    import strata_fdtd_b200 as sb
    with pytest.raises(ValueError):
        sb.FDTDSolver(shape=(8, 8, 8), resolution=1e-3, backend="python")


def test_membrane_sources_pressure_and_velocity_injection():
    """Weighted region sources into p and into vx (reference membranes) -- vs the reference fixture and the oracle."""
    from util import load_membrane_case
    case, g = load_membrane_case()
    s = build_b200_solver(case, chunk_steps=50)
    s.run(steps=case["steps"])
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(s.get_field(f), g["final_" + f]), f
    for n in s._probes:
        assert np.array_equal(s.get_probe_data(n)[n], g["probe_" + n])
    o = O.OracleSolver(case); o.run_steps(case["steps"])
    assert_same_as_oracle(s, o, "membranes")


def test_result_writer_and_callbacks(tmp_path):
    """run(output_file=..., snapshot_interval=..., callback=...) -- traces in the file equal the probe data,
    one callback per step, snapshots at the reference's steps (solver.py:2572-2584, io/hdf5.py:164-207)."""
    case = CASES["uniform_pml"]
    s = build_b200_solver(case, chunk_steps=16)
    seen = []
    out = tmp_path / "res.h5"
    s.enable_snapshots(25)
    s.run(steps=60, output_file=str(out), snapshot_interval=20, callback=seen.append, script_content="# test")
    assert seen == list(range(60))
    assert [round(t / s.dt) for t, _ in s.get_snapshots()] == [0, 25, 50]
    o = O.OracleSolver(case)
    for n in range(60):
        o.step()
        if n == 50:
            assert np.array_equal(s.get_snapshots()[2][1], o.p)
    from strata_fdtd_b200 import io as sbio
    if sbio.HAVE_H5PY:
        import h5py
        with h5py.File(out, "r") as f:
            assert np.array_equal(f["probes/a"][:], o.probe_array("a"))
            assert f["fields/pressure"].shape[0] == 3 and f["simulation"].attrs["num_steps"] == 60
    else:
        z = np.load(str(out), allow_pickle=False)
        assert np.array_equal(z["probes/a"], o.probe_array("a"))
        assert z["fields/pressure"].shape == (3,) + tuple(case["shape"])
        import json
        attrs = json.loads(str(z["__attrs__"]))
        assert attrs["simulation@num_steps"] == 60 and attrs["metadata@backend"] == "b200"


def test_example_scripts_run_unchanged_through_the_alias(tmp_path):
    """Scripts that only know ``from strata_fdtd import ...`` run on the b200 backend via
    ``python -m strata_fdtd_b200 script.py`` (the fdtd-compute stand-in on a box without the reference)."""
    import subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    res = subprocess.run([sys.executable, "-m", "strata_fdtd_b200", str(root / "examples" / "basic_pulse.py")],
                         cwd=tmp_path, capture_output=True, text=True, timeout=300,
                         env={**__import__("os").environ, "PYTHONPATH": str(root)})
    assert res.returncode == 0, res.stderr[-2000:]
    assert "steps: 1000" in res.stdout and (tmp_path / "basic_pulse_results.h5").exists()
    g = np.load(GOLDEN / "c1_100cubed_1000.npz")
    peak = float(np.abs(g["probe_probe"]).max())
    assert f"{peak:.4e}" in res.stdout                      # same trace as the reference's native backend
    res = subprocess.run([sys.executable, "-m", "strata_fdtd_b200", str(root / "examples" / "material_sphere.py"), "64"],
                         cwd=tmp_path, capture_output=True, text=True, timeout=300,
                         env={**__import__("os").environ, "PYTHONPATH": str(root)})
    assert res.returncode == 0 and "64 probes" in res.stdout, res.stderr[-2000:]


def test_lifecycle_reset_rerun_late_additions_and_inplace_geometry():
    """reset() reproduces the first run; sources / probes added after stepping take effect; geometry edited in
    place WITHOUT set_geometry zeroes p in solids but leaves faces open, as the native backend does
    (boundary lists exist only after set_geometry, core/solver.py:1779-1780, 2094)."""
    case = CASES["block_pml"]
    s = build_b200_solver(case, chunk_steps=33)
    s.run(steps=90)
    first = {f: s.get_field(f) for f in ("p", "vx", "vy", "vz")}
    tr = s.get_probe_data("shadow")["shadow"].copy()
    s.reset()
    assert len(s.get_probe_data("shadow")["shadow"]) == 0
    s.run(steps=45); s.run(steps=45)
    for f in first:
        assert np.array_equal(s.get_field(f), first[f]), f
    assert np.array_equal(s.get_probe_data("shadow")["shadow"], tr)
    # late additions
    import strata_fdtd_b200 as sb
    o = O.OracleSolver(case); o.run_steps(90)
    s.add_source(sb.GaussianPulse(position=(40, 30, 10), frequency=12e3))
    s.add_probe("late", (41, 30, 10))
    o.sources.append(dict(kind="point", position=(40, 30, 10), frequency=12e3)); o.probes.append(("late", (41, 30, 10)))
    o.probe_data["late"] = []
    s.run(steps=40); o.run_steps(40)
    assert_same_as_oracle(s, o, "late additions")
    # in-place geometry edit without set_geometry
    plain = dict(shape=(20, 18, 22), resolution=1e-3, steps=0, pml=[dict(depth=3)],
                 sources=[dict(kind="point", position=(5, 9, 11), frequency=20e3)], probes=[("q", (15, 9, 11))])
    s2 = build_b200_solver(plain); o2 = O.OracleSolver(plain)
    s2.geometry[9:12, 6:12, 8:14] = False
    o2.geometry[9:12, 6:12, 8:14] = False
    s2.run(steps=80); o2.run_steps(80)
    assert not o2.rigid
    assert_same_as_oracle(s2, o2, "in-place geometry")


# ---------------------------------------------------------------------------------------------------------
# K5: shared-memory-resident chunk kernel (csrc/sb_resident.cuh).  The same cases the CPU emulation checks
# (tests/test_resident_emulation.py), now with the real barriers / step flags / cooperative launch.
def _resident_cases():
    from test_resident_emulation import EMU_CASES
    return EMU_CASES


@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("name", sorted(_resident_cases()))
def test_resident_kernel_matches_oracle(name, split):
    case = _resident_cases()[name]
    s = _with_options(build_b200_solver(case, chunk_steps=37),
                      {_lib.OPT_KERNEL: _lib.KERNEL_RESIDENT, _lib.OPT_RESIDENT_SPLIT: split})
    o = O.OracleSolver(case)
    s.run(steps=case["steps"]); o.run_steps(case["steps"])
    assert_same_as_oracle(s, o, f"resident/{name}/split{split}")
    st = s.device_stats()
    assert st["kernel_variant"] == _lib.KERNEL_RESIDENT
    assert st["kernels_launched"] <= -(-case["steps"] // 37) + 2, "one launch per chunk"
    s.close()


def test_resident_is_the_automatic_choice_for_config_1_and_matches_the_reference_fixture():
    """BASELINE config 1 (100^3, 1000 steps): AUTO keeps the grid in shared memory; output == the reference's."""
    g = np.load(GOLDEN / "c1_100cubed_1000.npz")
    s = build_b200_solver(c1_case(1000))
    s.run(steps=1000)
    assert s.device_stats()["kernel_variant"] == _lib.KERNEL_RESIDENT
    assert s.kernel_launches() <= 6
    assert np.array_equal(s.get_probe_data("probe")["probe"], g["probe_probe"])
    for f in ("p", "vx", "vy", "vz"):
        assert sha(s.get_field(f)) == str(g["sha_" + f])
    # a single step() afterwards goes through K1 and must continue from the stored state
    m = _with_options(build_b200_solver(c1_case(1000)), {_lib.OPT_KERNEL: _lib.KERNEL_MARCH})
    m.run(steps=1000)
    s.step(); m.step()
    assert s.device_stats()["kernel_variant"] == _lib.KERNEL_MARCH
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(s.get_field(f), m.get_field(f)), f


@pytest.mark.parametrize("shape", [(150, 7, 9), (5, 300, 12), (3, 3, 3), (40, 40, 1), (97, 101, 103), (2, 2, 600)])
def test_resident_equals_march_on_awkward_shapes(shape):
    """Two independent implementations (K1 streams through L2, K5 stays in shared memory) on shapes that
    stress the box partition: more SMs than planes, one-cell axes, extents that no box count divides."""
    nx, ny, nz = shape
    geom = np.ones(shape, dtype=bool)
    geom[nx // 3: nx // 3 + 2, ny // 2:, : max(1, nz // 3)] = False
    case = dict(shape=shape, resolution=1e-3, steps=120, geometry=geom,
                pml=[dict(depth=min(3, max(1, min(shape) // 3)))] if min(shape) >= 3 else [],
                sources=[dict(kind="point", position=(nx // 2, ny // 4, nz // 2), frequency=30e3),
                         dict(kind="point", position=(0, 0, 0), frequency=12e3, amplitude=0.3)],
                probes=[("a", (nx - 1, ny - 1, nz - 1)), ("b", (nx // 2, ny // 2, nz // 2)), ("c", (0, ny - 1, 0))])
    a = _with_options(build_b200_solver(case, chunk_steps=50), {_lib.OPT_KERNEL: _lib.KERNEL_RESIDENT})
    b = _with_options(build_b200_solver(case, chunk_steps=50), {_lib.OPT_KERNEL: _lib.KERNEL_MARCH})
    a.run(steps=120); b.run(steps=120)
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(a.get_field(f), b.get_field(f)), f
    for n in ("a", "b", "c"):
        assert np.array_equal(a.get_probe_data(n)[n], b.get_probe_data(n)[n]), n
    assert np.abs(a.get_field("p")).max() > 0
    assert a.device_stats()["kernel_variant"] == _lib.KERNEL_RESIDENT


@pytest.mark.parametrize("split", [0, 1])
@pytest.mark.parametrize("name", ["mur_all", "pml_radiation_mur"])
def test_resident_kernel_applies_mur_and_radiation_planes(name, split):
    """Mur / radiation planes inside K5: after the pressure phase every box updates the face cells it holds in shared memory,
    plane by plane in list order, then adds its sources and only then publishes its faces to the neighbouring boxes."""
    case = CASES[name]
    s = _with_options(build_b200_solver(case, chunk_steps=37), {_lib.OPT_KERNEL: _lib.KERNEL_RESIDENT, _lib.OPT_RESIDENT_SPLIT: split})
    o = O.OracleSolver(case)
    s.run(steps=case["steps"]); o.run_steps(case["steps"])
    assert_same_as_oracle(s, o, f"resident-planes/{name}/split{split}")
    st = s.device_stats()
    assert st["kernel_variant"] == _lib.KERNEL_RESIDENT
    assert st["kernels_launched"] <= -(-case["steps"] // 37) + 2, "one launch per chunk"
    s.close()


@pytest.mark.parametrize("shape", [(64, 64, 64), (100, 100, 100), (33, 13, 3), (5, 7, 130), (148, 3, 9)])
def test_small_grids_with_planes_run_resident_by_default_and_match_the_stepwise_path(shape):
    """AUTO: with planes a grid that fits in shared memory runs K5 (box grids of >= 2 planes / rows where a face pair needs
    them); sources sit on faces, edges and next to them, two of them in one cell; equals K1 + K4 step by step and the oracle."""
    nx, ny, nz = shape
    geom = np.ones(shape, dtype=bool)
    geom[nx // 3: nx // 3 + 2, ny // 2:, : max(1, nz // 3)] = False
    case = dict(shape=shape, resolution=1e-3, steps=70, geometry=geom, pml=[dict(depth=2, axes=("z",))] if nz >= 6 else [],
                plane_bcs=[dict(kind="mur", axes=("y", "x")), dict(kind="radiation", axis="z", side="low", reflection_coeff=0.3),
                           dict(kind="radiation", axis="x", side="high", pipe_radius=0.01), dict(kind="mur", axes=("z",))],
                sources=[dict(kind="point", position=(0, 0, 0), frequency=25e3), dict(kind="point", position=(nx - 1, ny - 2, nz - 1), frequency=18e3),
                         dict(kind="point", position=(nx // 2, ny // 2, nz // 2), frequency=30e3),
                         dict(kind="point", position=(nx // 2, ny // 2, nz // 2), frequency=11e3, amplitude=0.5), dict(kind="point", position=(1, 1, 1), frequency=21e3)],
                probes=[("corner", (nx - 1, ny - 1, nz - 1)), ("origin", (0, 0, 0)), ("edge", (nx - 1, 0, nz - 2)), ("mid", (nx // 2, ny // 2, nz // 3))])
    a = build_b200_solver(case, chunk_steps=35)
    b = _with_options(build_b200_solver(case, chunk_steps=35), {_lib.OPT_KERNEL: _lib.KERNEL_MARCH})
    a.run(steps=70); b.run(steps=70)
    assert a.device_stats()["kernel_variant"] == _lib.KERNEL_RESIDENT
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(a.get_field(f), b.get_field(f)), f
    for n in ("corner", "origin", "edge", "mid"):
        assert np.array_equal(a.get_probe_data(n)[n], b.get_probe_data(n)[n]), n
    if nx * ny * nz <= 300_000:
        o = O.OracleSolver(case)
        o.run_steps(70)
        assert_same_as_oracle(a, o, f"resident-planes/{shape}")
    assert np.abs(a.get_field("p")).max() > 0
    a.close(); b.close()


def test_resident_refuses_what_it_cannot_do_and_auto_falls_back_to_k1():
    case = CASES["directional_mics"]                          # velocity gathers that cross boxes stay on the K1 path
    s = _with_options(build_b200_solver(case), {_lib.OPT_KERNEL: _lib.KERNEL_RESIDENT})
    with pytest.raises(_lib.B200BackendError, match="resident kernel not applicable: microphones"):
        s.run(steps=8)
    s.close()
    big = _with_options(build_b200_solver(c2_case(200, steps=0)), {_lib.OPT_KERNEL: _lib.KERNEL_RESIDENT})
    with pytest.raises(_lib.B200BackendError, match="does not fit in shared memory"):
        big.run(steps=8)
    big.close()
    a = build_b200_solver(case)
    a.run(steps=16)
    assert a.device_stats()["kernel_variant"] == _lib.KERNEL_MARCH
    a.close()


def _array_case(n_src):
    """A line array of point sources (pzt_transducer.py has 16), two of them in the same cell, several in one float4."""
    shape = (30, 26, 44)
    src = [dict(kind="point", position=(4 + (q % 3), 3 + (q * 5) % 20, 6 + (2 * q) % 34), frequency=20e3 + 900.0 * q,
                amplitude=0.5 + 0.05 * q) for q in range(n_src)]
    src[1]["position"] = src[0]["position"]
    src[2]["position"] = (src[0]["position"][0], src[0]["position"][1], src[0]["position"][2] + 1)
    return dict(shape=shape, resolution=1e-3, steps=96, pml=[dict(depth=3)], geometry=_np_block(shape), sources=src,
                probes=[("a", (20, 13, 22)), ("at_src", src[0]["position"]), ("far", (29, 25, 43))])


def _np_block(shape):
    g = np.ones(shape, dtype=bool)
    g[14:17, 8:20, 10:30] = False
    return g


@pytest.mark.parametrize("kernel", [_lib.KERNEL_RESIDENT, _lib.KERNEL_PIPELINE, _lib.KERNEL_MARCH])
def test_chunk_kernels_take_a_phased_array_of_point_sources(kernel):
    """Up to 32 point-source entries are applied by the step / chunk kernels themselves (K5, K6, K1 with K3 fused in)."""
    case = _array_case(24)
    s = _with_options(build_b200_solver(case, chunk_steps=32), {_lib.OPT_KERNEL: kernel})
    o = O.OracleSolver(case)
    s.run(steps=96); o.run_steps(96)
    assert s.device_stats()["kernel_variant"] == kernel
    assert_same_as_oracle(s, o, f"array/{kernel}")
    if kernel != _lib.KERNEL_MARCH:
        assert s.device_stats()["kernels_launched"] <= 5, "one launch per chunk"
    s.close()


def test_more_sources_than_the_chunk_kernels_hold_run_step_by_step():
    case = _array_case(40)
    s = _with_options(build_b200_solver(case, chunk_steps=32), {_lib.OPT_KERNEL: _lib.KERNEL_RESIDENT})
    with pytest.raises(_lib.B200BackendError, match="more than 32 source cells"):
        s.run(steps=8)
    s.close()
    a = build_b200_solver(case, chunk_steps=32)
    o = O.OracleSolver(case)
    a.run(steps=96); o.run_steps(96)
    assert a.device_stats()["kernel_variant"] == _lib.KERNEL_MARCH
    assert_same_as_oracle(a, o, "array/40")
    a.close()


# ---------------------------------------------------------------------------------------------------------
# K6: the marching kernel pipelined across steps in one persistent launch (csrc/sb_pipeline.cuh)
@pytest.mark.parametrize("shape_opts", [{}, {_lib.OPT_ROWS_PER_THREAD: 2, _lib.OPT_CHUNK_I: 3, _lib.OPT_WARPS_J: 2, _lib.OPT_PLANE_MAP: 2},
                                        {_lib.OPT_ROWS_PER_THREAD: 1, _lib.OPT_CHUNK_I: 1, _lib.OPT_WARPS_J: 1, _lib.OPT_PLANE_MAP: 1},
                                        {_lib.OPT_ROWS_PER_THREAD: 1, _lib.OPT_CHUNK_I: 5, _lib.OPT_PLANE_MAP: 2}])
@pytest.mark.parametrize("name", sorted(_resident_cases()))
def test_pipelined_kernel_matches_oracle(name, shape_opts):
    case = _resident_cases()[name]
    s = _with_options(build_b200_solver(case, chunk_steps=37), {_lib.OPT_KERNEL: _lib.KERNEL_PIPELINE, **shape_opts})
    o = O.OracleSolver(case)
    s.run(steps=case["steps"]); o.run_steps(case["steps"])
    assert_same_as_oracle(s, o, f"pipeline/{name}")
    st = s.device_stats()
    assert st["kernel_variant"] == _lib.KERNEL_PIPELINE
    assert st["kernels_launched"] <= -(-case["steps"] // 37) + 2, "one launch per chunk"
    s.close()


@pytest.mark.parametrize("shape_opts", [{}, {_lib.OPT_ROWS_PER_THREAD: 2, _lib.OPT_CHUNK_I: 4, _lib.OPT_WARPS_J: 2},
                                        {_lib.OPT_ROWS_PER_THREAD: 1, _lib.OPT_CHUNK_I: 5, _lib.OPT_WARPS_J: 3, _lib.OPT_PLANE_MAP: 2}])
@pytest.mark.parametrize("name", ["mur_all", "pml_radiation_mur"])
def test_pipelined_kernel_applies_mur_and_radiation_planes(name, shape_opts):
    """Mur / radiation planes inside K6: every tile updates the face cells it owns, plane by plane in list order (edges and
    corners depend on it), before its sources and probes -- one launch per chunk of steps, no K4 launches."""
    case = CASES[name]
    s = _with_options(build_b200_solver(case, chunk_steps=37), {_lib.OPT_KERNEL: _lib.KERNEL_PIPELINE, **shape_opts})
    o = O.OracleSolver(case)
    s.run(steps=case["steps"]); o.run_steps(case["steps"])
    assert_same_as_oracle(s, o, f"pipeline-planes/{name}")
    st = s.device_stats()
    assert st["kernel_variant"] == _lib.KERNEL_PIPELINE
    assert st["kernels_launched"] <= -(-case["steps"] // 37) + 2, "one launch per chunk"
    s.close()


@pytest.mark.parametrize("shape", [(25, 21, 129), (17, 9, 129), (9, 5, 30), (33, 13, 3)])
def test_pipelined_kernel_picks_tiles_that_keep_plane_pairs_together(shape):
    """Extents where the default tiles would separate a high face from its interior neighbour ((n - 1) % tile == 0 along
    i, j or k): the launcher picks another tile shape; results equal the step-by-step path (K1 + K4)."""
    nx, ny, nz = shape
    case = dict(shape=shape, resolution=1e-3, steps=80, pml=[],
                plane_bcs=[dict(kind="mur", axes=("z", "x")), dict(kind="radiation", axis="y", side="high", reflection_coeff=0.4),
                           dict(kind="mur", axes=("y",))],
                sources=[dict(kind="point", position=(nx // 2, ny // 2, nz // 2), frequency=30e3),
                         dict(kind="point", position=(nx - 1, ny - 1, nz - 1), frequency=14e3, amplitude=0.4)],
                probes=[("corner", (nx - 1, ny - 1, nz - 1)), ("origin", (0, 0, 0)), ("edge", (nx - 1, 0, nz - 2)), ("mid", (nx // 2, ny // 2, nz // 3))])
    a = _with_options(build_b200_solver(case, chunk_steps=40), {_lib.OPT_KERNEL: _lib.KERNEL_PIPELINE})
    b = _with_options(build_b200_solver(case, chunk_steps=40), {_lib.OPT_KERNEL: _lib.KERNEL_MARCH})
    o = O.OracleSolver(case)
    a.run(steps=80); b.run(steps=80); o.run_steps(80)
    assert a.device_stats()["kernel_variant"] == _lib.KERNEL_PIPELINE
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(a.get_field(f), b.get_field(f)), f
    assert_same_as_oracle(a, o, f"pipeline-planes/{shape}")
    a.close(); b.close()


def test_pipelined_kernel_refuses_a_fixed_tile_shape_that_separates_a_plane_pair():
    case = dict(shape=(25, 21, 20), resolution=1e-3, steps=8, pml=[], plane_bcs=[dict(kind="mur", axes=("x",))],
                sources=[dict(kind="point", position=(12, 10, 10), frequency=30e3)], probes=[("a", (0, 0, 0))])
    s = _with_options(build_b200_solver(case), {_lib.OPT_KERNEL: _lib.KERNEL_PIPELINE, _lib.OPT_ROWS_PER_THREAD: 1, _lib.OPT_CHUNK_I: 8})
    with pytest.raises(_lib.B200BackendError, match="no tile shape keeps every Mur / radiation face"):
        s.run(steps=8)
    s.close()
    # ... and where no shape at all does (9 rows and 257 columns: 8 % rows and 256 % columns are 0 for every shape of at
    # most 8 warps), the automatic choice quietly stays with the step-by-step path
    case = dict(case, shape=(900, 9, 257), plane_bcs=[dict(kind="mur", axes=("x", "y", "z"))],     # (too large for K5)
                sources=[dict(kind="point", position=(8, 4, 128), frequency=30e3)], steps=24)
    a = build_b200_solver(case)
    o = O.OracleSolver(case)
    a.run(steps=24); o.run_steps(24)
    assert a.device_stats()["kernel_variant"] == _lib.KERNEL_MARCH
    assert_same_as_oracle(a, o, "planes/no-pipeline-shape")
    a.close()


def test_pipelined_kernel_is_the_automatic_choice_for_config_2_and_equals_the_oracle():
    """BASELINE config 2 (200^3 + PML, here with the solid block): too large for shared memory, AUTO pipelines the steps."""
    case = c2_case(200, steps=120, with_geometry=True)
    s = build_b200_solver(case, chunk_steps=64)
    o = O.OracleSolver(case)
    s.run(steps=120); o.run_steps(120)
    assert s.device_stats()["kernel_variant"] == _lib.KERNEL_PIPELINE
    assert_same_as_oracle(s, o, "c2/pipeline")
    s.close()


@pytest.mark.parametrize("shape", [(150, 7, 9), (5, 300, 12), (3, 3, 3), (40, 40, 1), (97, 101, 103), (2, 2, 600), (300, 40, 260)])
def test_pipelined_equals_stepwise_on_awkward_shapes(shape):
    nx, ny, nz = shape
    geom = np.ones(shape, dtype=bool)
    geom[nx // 3: nx // 3 + 2, ny // 2:, : max(1, nz // 3)] = False
    case = dict(shape=shape, resolution=1e-3, steps=90, geometry=geom,
                pml=[dict(depth=min(3, max(1, min(shape) // 3)))] if min(shape) >= 3 else [],
                sources=[dict(kind="point", position=(nx // 2, ny // 4, nz // 2), frequency=30e3),
                         dict(kind="point", position=(0, 0, 0), frequency=12e3, amplitude=0.3)],
                probes=[("a", (nx - 1, ny - 1, nz - 1)), ("b", (nx // 2, ny // 2, nz // 2)), ("c", (0, ny - 1, 0))])
    a = _with_options(build_b200_solver(case, chunk_steps=50), {_lib.OPT_KERNEL: _lib.KERNEL_PIPELINE})
    b = _with_options(build_b200_solver(case, chunk_steps=50), {_lib.OPT_KERNEL: _lib.KERNEL_MARCH})
    a.run(steps=90); b.run(steps=90)
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(a.get_field(f), b.get_field(f)), f
    for n in ("a", "b", "c"):
        assert np.array_equal(a.get_probe_data(n)[n], b.get_probe_data(n)[n]), n
    assert np.abs(a.get_field("p")).max() > 0
    assert a.device_stats()["kernel_variant"] == _lib.KERNEL_PIPELINE


@pytest.mark.parametrize("kernel", [_lib.KERNEL_RESIDENT, _lib.KERNEL_PIPELINE, _lib.KERNEL_AUTO])
@pytest.mark.parametrize("name", ["uniform_pml", "nonuniform_block_pml", "odd_geometry_pml"])
def test_omni_microphones_ride_the_chunk_kernels_as_corner_probes(name, kernel):
    """Omnidirectional microphones = eight raw p samples (probes) summed on the host in the reference's order, so
    cases with microphones run through K5 / K6 too; traces must equal the reference's own (golden) bit for bit."""
    case = CASES[name]
    g = np.load(GOLDEN / f"{name}.npz")
    s = _with_options(build_b200_solver(case, chunk_steps=41), {_lib.OPT_KERNEL: kernel})
    s.run(steps=case["steps"])
    for mname, mic in s.microphones.items():
        assert np.array_equal(mic.get_waveform(), g["mic_" + mname]), mname
        assert len(mic.get_time_axis()) == case["steps"]
    for pname in s._probes:
        assert np.array_equal(s.get_probe_data(pname)[pname], g["probe_" + pname])
    for f in ("p", "vx", "vy", "vz"):
        assert sha(s.get_field(f)) == str(g["sha_" + f])
    want = {_lib.KERNEL_AUTO: _lib.KERNEL_RESIDENT}.get(kernel, kernel)
    assert s.device_stats()["kernel_variant"] == want
    s.close()


def test_c3_ade_sphere_at_128_cubed_vs_oracle():
    """BASELINE config 3 at 1/4 scale (128^3, PML 10, ADE sphere of radius N/10 with 2 Debye + 1 Lorentz poles,
    64 probes): every field and all 64 traces equal the oracle (= the reference's ade.cpp kernels in the documented
    order, oracle/oracle.py) bit for bit."""
    case = c3_case(128, steps=150)
    assert int((np.asarray(case["material_id"]) > 0).sum()) > 8000
    s = build_b200_solver(case, chunk_steps=64)
    o = O.OracleSolver(case)
    s.run(steps=150); o.run_steps(150)
    assert_same_as_oracle(s, o, "c3/128")
    assert len(s._probes) == 64 and np.abs(s.get_field("p")).max() > 0
    s.close()


FUSED_SHAPES = {
    "auto": {},
    "r2_chunk5": {_lib.OPT_ROWS_PER_THREAD: 2, _lib.OPT_CHUNK_I: 5, _lib.OPT_WARPS_J: 2},
    "r1_flat_chunk3": {_lib.OPT_ROWS_PER_THREAD: 1, _lib.OPT_PLANE_MAP: 2, _lib.OPT_CHUNK_I: 3},
    "r2_strips_graph": {_lib.OPT_ROWS_PER_THREAD: 2, _lib.OPT_PLANE_MAP: 1, _lib.OPT_USE_GRAPH: 1},
    "r1_flat_nograph": {_lib.OPT_ROWS_PER_THREAD: 1, _lib.OPT_PLANE_MAP: 2, _lib.OPT_USE_GRAPH: 0, _lib.OPT_WARPS_J: 4},
}


@pytest.mark.parametrize("shape", sorted(FUSED_SHAPES))
@pytest.mark.parametrize("name", ["ade_sphere", "ade_two_materials_nonuniform", "ade_dense_layers"])
def test_fused_ade_kernel_matches_oracle(name, shape):
    """Layout 3: the material cells are updated inside a variant of the step kernel (K1-ADE) while the plain K1 launch
    skips their bounding box.  Every launch shape of the plain kernel (rows per thread, strips / flat plane mapping,
    chunk length, CUDA graph) against the oracle, including Lorentz density poles (three rotating J buffers), two
    materials touching each other, solids inside a material and materials up to the outer faces of the grid."""
    case = CASES[name]
    s = _with_options(build_b200_solver(case, chunk_steps=53), {_lib.OPT_ADE_LAYOUT: 3, **FUSED_SHAPES[shape]})
    o = O.OracleSolver(case)
    s.run(steps=case["steps"]); o.run_steps(case["steps"])
    assert_same_as_oracle(s, o, f"ade/{name}/fused/{shape}")
    assert np.abs(s.get_field("p")).max() > 0
    s.reset(); o2 = O.OracleSolver(case)
    s.run(steps=41); o2.run_steps(41)
    assert_same_as_oracle(s, o2, f"ade/{name}/fused/{shape}/after reset")
    s.close()


def test_fused_ade_survives_a_later_geometry_change():
    """The ADE bits share the mask bytes with the geometry: setting a geometry after the materials (and removing it
    again) must keep them."""
    case = dict(CASES["ade_sphere"])
    s = _with_options(build_b200_solver(case, chunk_steps=32), {_lib.OPT_ADE_LAYOUT: 3})
    s.run(steps=20)
    g = np.ones(case["shape"], dtype=bool); g[18:22, 10:14, 12:20] = False          # a solid inside the sphere
    s.set_geometry(g)
    s.run(steps=60)
    case2 = dict(case, geometry=g)
    o = O.OracleSolver(case); o.run_steps(20)
    o2 = O.OracleSolver(case2)
    for f in ("p", "vx", "vy", "vz"):
        getattr(o2, f)[...] = getattr(o, f)
    for a, b in zip(o2.debye + o2.lorentz, o.debye + o.lorentz):
        a["J"][...] = b["J"]
        if "Jp" in a:
            a["Jp"][...] = b["Jp"]
    o2.time, o2.step_count = o.time, o.step_count
    for name, _ in o2.probes:
        o2.probe_data[name] = list(o.probe_data[name])
    o2.run_steps(60)
    assert_same_as_oracle(s, o2, "fused ade + late geometry")
    s.close()


@pytest.mark.parametrize("layout", [1, 2])
@pytest.mark.parametrize("name", ["ade_sphere", "ade_two_materials_nonuniform", "ade_dense_layers"])
def test_ade_layouts_match_oracle(name, layout):
    """Compact list (1) and dense bounding-box layout (2) of the material cells are two data layouts of the same
    arithmetic: both equal the oracle (= the reference's ade.cpp kernels) bit for bit on every ADE case."""
    case = CASES[name]
    s = _with_options(build_b200_solver(case, chunk_steps=53), {_lib.OPT_ADE_LAYOUT: layout})
    o = O.OracleSolver(case)
    s.run(steps=case["steps"]); o.run_steps(case["steps"])
    assert_same_as_oracle(s, o, f"ade/{name}/layout{layout}")
    s.reset(); o2 = O.OracleSolver(case)                     # J fields are zeroed by reset in either layout
    s.run(steps=40); o2.run_steps(40)
    assert_same_as_oracle(s, o2, f"ade/{name}/layout{layout}/after reset")
    s.close()


@pytest.mark.parametrize("kernel", [_lib.KERNEL_AUTO, _lib.KERNEL_MARCH, _lib.KERNEL_PIPELINE])
@pytest.mark.parametrize("name", ["mur_all", "pml_radiation_mur"])
def test_checkpoint_and_resume_with_plane_boundaries(name, kernel):
    """The previous-plane arrays of Mur / radiation faces are part of the state: get_state() after 90 steps -> set_state()
    on a fresh solver running another kernel -> the continued run equals the uninterrupted one and the oracle."""
    case = CASES[name]
    a = _with_options(build_b200_solver(case, chunk_steps=41), {_lib.OPT_KERNEL: kernel})
    a.run(steps=90)
    st = a.get_state()
    assert len(st["planes"]) >= 4 and all(np.abs(q).max() > 0 for q in st["planes"])
    b = _with_options(build_b200_solver(case, chunk_steps=29), {_lib.OPT_KERNEL: _lib.KERNEL_MARCH if kernel != _lib.KERNEL_MARCH else _lib.KERNEL_AUTO})
    b.set_state(st)
    o = O.OracleSolver(case)
    a.run(steps=70); b.run(steps=70); o.run_steps(160)
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(a.get_field(f), b.get_field(f)), f
        assert np.array_equal(b.get_field(f), getattr(o, f)), f
    a.close(); b.close()


@pytest.mark.parametrize("layout", [1, 2, 3])
@pytest.mark.parametrize("name", ["ade_two_materials_nonuniform", "ade_dense_layers"])
def test_checkpoint_and_resume_with_auxiliary_fields(name, layout):
    """get_state() after 90 steps -> set_state() on a fresh solver (any ADE device layout, here a different one) -> the
    continued run equals the uninterrupted one bit for bit; the auxiliary fields themselves equal the oracle's."""
    case = CASES[name]
    a = _with_options(build_b200_solver(case, chunk_steps=41), {_lib.OPT_ADE_LAYOUT: layout})
    a.run(steps=90)
    st = a.get_state()
    o = O.OracleSolver(case)
    o.run_steps(90)
    want = [(e, "J") for e in o.debye] + [(e, "J") for e in o.lorentz]
    assert len(st["ade"]) == len(want)
    for got, (e, key) in zip(st["ade"], want):
        assert got["material_id"] == e["mat"] and got["target"] == e["target"]
        sel = np.asarray(case["material_id"]) == e["mat"]
        assert np.array_equal(got["J"][sel], e["J"][sel]) and not got["J"][~sel].any()
        if "Jp" in e:
            assert np.array_equal(got["J_prev"][sel], e["Jp"][sel])
    b = _with_options(build_b200_solver(case, chunk_steps=29), {_lib.OPT_ADE_LAYOUT: 1 + layout % 3})
    b.set_state(st)
    a.run(steps=70); b.run(steps=70); o.run_steps(70)
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(a.get_field(f), b.get_field(f)), f
        assert np.array_equal(b.get_field(f), getattr(o, f)), f
    assert b.step_count == 160 and b.time == a.time
    a.close(); b.close()
