"""CPU tests of the host side: the C ABI library loads and exports every declared symbol, and the
host mirrors (grids, sponge profiles, position rounding, pole coefficients, source tables) reproduce
the reference's numbers exactly (fixtures; live reference where available)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import strata_fdtd_b200 as sb
from cases import make_cases
from oracle import oracle as O
from oracle import ref_loader as R
from strata_fdtd_b200 import _lib
from util import build_b200_solver

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"
CASES = make_cases()


def test_library_builds_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    header = (ROOT / "include" / "strata_b200.h").read_text()
    declared = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", header))
    declared -= {"sb_solver"}
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/strata_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.sb_abi_version() == 1
    # no-GPU calls only
    pitch = ctypes.c_int32(0)
    assert lib.sb_choose_pitch(100, ctypes.byref(pitch)) == 0 and pitch.value == 104
    d = _lib.GridDesc(nx=10, ny=7, nz=100, pitch=0, global_nx=10, i_offset=0, has_lower=0, has_upper=0)
    assert lib.sb_field_elems(ctypes.byref(d)) == 12 * 7 * 104


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    s = sb.FDTDSolver(shape=(8, 8, 8), resolution=1e-3)
    with pytest.raises(sb.B200BackendError, match="no CPU fallback"):
        s.step()
    lib = _lib.load()
    d = _lib.GridDesc(nx=8, ny=8, nz=8, pitch=0, global_nx=8, i_offset=0, has_lower=0, has_upper=0)
    h = ctypes.c_void_p()
    assert lib.sb_create(ctypes.byref(d), 0, None, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in lib.sb_last_error()


def test_decay_table_is_libm_expf():
    sig = np.linspace(0, 4e5, 57, dtype=np.float32)
    dt = 1.599075089592258e-06
    got = sb.boundaries.decay_table(sig, dt)
    want = np.empty_like(sig)
    O.lib().orc_decay_table(sig.ctypes.data_as(O._f32p), 57, ctypes.c_float(dt), want.ctypes.data_as(O._f32p))
    assert np.array_equal(got, want)
    assert got[0] == 1.0


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_numbers_match_golden(name):
    case = CASES[name]
    g = np.load(GOLDEN / f"{name}.npz")
    s = build_b200_solver(case)
    assert float(s.dt) == float(g["dt"])
    faces, cells, cp = s._coefficient_tables()
    assert np.float32(cp) == g["cp"]
    if s.grid.is_uniform:
        assert all(np.all(t == g["cv"]) for t in faces)
    else:
        for a, t, c in zip("xyz", faces, cells):
            assert np.array_equal(c, g[f"sp_inv_d{a}_cell"])
            assert np.array_equal(t[:-1], np.float32(g["cv"]) * g[f"sp_inv_d{a}_face"])
    for bi, b in enumerate(s._boundaries):
        if not hasattr(b, "_max_sigma"):
            continue
        assert float(b._max_sigma) == float(g[f"pml{bi}_max_sigma"])
        for a, sig, dec in zip("xyz", (b._sigma_x, b._sigma_y, b._sigma_z), b._decay):
            if sig is not None:
                assert np.array_equal(sig, g[f"pml{bi}_sigma_{a}"])
                assert np.array_equal(dec, g[f"pml{bi}_decay_{a}"])
    ny, nz = s.shape[1], s.shape[2]
    for pname, pr in s._probes.items():
        i, j, k = pr.position
        assert (i * ny + j) * nz + k == int(g["probe_idx_" + pname])
    for si, src in enumerate(s._sources):
        if src.source_type == "point":
            i, j, k = src.position
            assert (i * ny + j) * nz + k == int(g[f"source_idx_{si}"])
    if s.microphones and "mic_flat_indices" in g:
        lib = _lib.load()
        mics = list(s.microphones.values())
        gp = np.array([q for m in mics for q in m._grid_position], dtype=np.float32)
        idx8 = np.zeros(8 * len(mics), dtype=np.int64); w8 = np.zeros(8 * len(mics), dtype=np.float32)
        assert lib.sb_mic_tables(_lib.ptr(gp), len(mics), ny, nz, _lib.ptr(idx8), _lib.ptr(w8)) == 0
        assert np.array_equal(idx8, g["mic_flat_indices"]) and np.array_equal(w8, g["mic_weights"])


def test_source_table_csr_orders_and_masks():
    case = CASES["partial_pml_plane"]
    s = build_b200_solver(case)
    cells, start, sids, flds, wts = s._build_source_table()
    g = case["geometry"]
    n_plane = int(g[8].sum())
    assert len(cells) == n_plane + 1 and start[-1] == len(sids)
    assert np.all(np.diff(cells) > 0) and np.all(wts == 1.0) and np.all(flds == 0)
    # a source inside a solid is dropped, as the reference's per-step geometry test does (solver.py:2420)
    s2 = build_b200_solver(dict(case, sources=[dict(kind="point", position=(15, 2, 12), frequency=1e3)]))
    assert len(s2._build_source_table()[0]) == 0


def test_waveform_table_equals_per_step_evaluation():
    s = build_b200_solver(CASES["uniform_pml"])
    t, times = 0.0, []
    for _ in range(300):
        times.append(t); t = t + s.dt
    W = s._waveform_table(np.array(times))
    src = s._sources[0]
    ref = np.array([src.waveform(np.array([q]), s.dt)[0] for q in times])
    assert W[:, 0].tobytes() == ref.tobytes()
    assert W[:, 0].tobytes() == np.array([O.gaussian_pulse(q, 20e3, None, 1.0) for q in times]).tobytes()


def test_position_rounding_and_errors():
    s = sb.FDTDSolver(shape=(17, 23, 13), resolution=2e-3)
    s.add_probe("m", (0.030, 0.040, 0.020))
    assert s._probes["m"].position == (15, 20, 10)
    s.add_probe("i", (3, 3, 3))
    assert s._probes["i"].position == (3, 3, 3)
    with pytest.raises(ValueError):
        s.add_probe("m", (1, 1, 1))
    with pytest.raises(ValueError):
        s.add_probe("far", (0.5, 0.01, 0.01))
    with pytest.raises(ValueError):
        s.add_microphone(position=(0.5, 0.01, 0.01), name="out")
    with pytest.raises(ValueError):
        s.set_geometry(np.ones((3, 3, 3), bool))
    m = s.add_microphone(position=(0.01, 0.01, 0.01), name="c", pattern="cardioid", direction=(0, 2, 0))
    assert m.is_directional() and m.direction == (0.0, 1.0, 0.0)
    with pytest.raises(ValueError):
        s.add_microphone(position=(0.01, 0.01, 0.01), name="bad", pattern="shotgun")


def test_pole_coefficients_match_oracle_restatement():
    dt = 1.599075089592258e-06
    for p in CASES["ade_two_materials_nonuniform"]["materials"][1]["poles"]:
        if p["type"] == "debye":
            pole = sb.Pole(sb.PoleType.DEBYE, p["delta_chi"], p["target"], tau=p["tau"])
        else:
            pole = sb.Pole(sb.PoleType.LORENTZ, p["delta_chi"], p["target"], omega_0=p["omega_0"], gamma=p["gamma"])
        assert pole.fdtd_coefficients(dt) == O.pole_coefficients(p, dt)


@pytest.mark.skipif(not R.have_reference_package(), reason="reference sources not on this box")
def test_grids_match_reference_live():
    ref = R.load_reference_package()
    for kw in (dict(shape=(41, 30, 52), base_resolution=1e-3, stretch_x=1.03, stretch_z=1.05),
               dict(shape=(40, 31, 20), base_resolution=2e-3, stretch_y=1.02, center_fine=False)):
        a, b = sb.NonuniformGrid.from_stretch(**kw), ref.NonuniformGrid.from_stretch(**kw)
        for attr in ("x_coords", "y_coords", "z_coords", "dx", "dy", "dz"):
            assert np.array_equal(getattr(a, attr), getattr(b, attr)), attr
        assert a.min_spacing == b.min_spacing and a.shape == b.shape
        sa, sb_ = a.get_spacing_arrays_for_stencil(), b.get_spacing_arrays_for_stencil()
        assert all(np.array_equal(sa[k], sb_[k]) for k in sb_)
    regs = [(0, 0.05, 2e-3), (0.05, 0.15, 1e-3), (0.15, 0.2, 2e-3)]
    a = sb.NonuniformGrid.from_regions(regs, [(0, 0.1, 1e-3)], [(0, 0.1, 1e-3)])
    b = ref.NonuniformGrid.from_regions(regs, [(0, 0.1, 1e-3)], [(0, 0.1, 1e-3)])
    assert np.array_equal(a.x_coords, b.x_coords) and a.shape == b.shape
    # the reference's own PML accepts our solver object (duck typing) and yields the same profiles
    s = sb.FDTDSolver(grid=sb.NonuniformGrid.from_stretch((40, 24, 30), 1e-3, stretch_x=1.04))
    ours, theirs = sb.PML(depth=6), ref.PML(depth=6)
    ours.initialize(s); theirs.initialize(s)
    for a_ in "xyz":
        assert np.array_equal(getattr(ours, "_sigma_" + a_), getattr(theirs, "_sigma_" + a_))


@pytest.mark.skipif(not R.have_reference_package(), reason="reference sources not on this box")
def test_script_solids_match_reference_live():
    """The few SDF shapes the reference's example scripts import (compat.py) against geometry/sdf.py: same distances,
    same voxels -- on a uniform and on a stretched grid."""
    ref = R.load_reference_package()
    from strata_fdtd_b200 import compat as C
    rng = np.random.default_rng(11)
    pts = rng.uniform(-0.02, 0.12, size=(4000, 3))

    def build(m):
        a = m.Box(center=(0.05, 0.04, 0.06), size=(0.04, 0.03, 0.05))
        b = m.Box(min_corner=(0.03, 0.03, 0.03), max_corner=(0.06, 0.09, 0.05))
        c = m.Sphere(center=(0.05, 0.05, 0.05), radius=0.022)
        return [a, b, c, m.Union(a, c), m.Intersection(a, b, c), m.Difference(a, c), m.Difference(m.Union(a, b), c, b)]
    grids = [(sb.UniformGrid((30, 28, 34), 3e-3), ref.UniformGrid((30, 28, 34), 3e-3)),
             (sb.NonuniformGrid.from_stretch((30, 28, 34), 2e-3, stretch_x=1.05), ref.NonuniformGrid.from_stretch((30, 28, 34), 2e-3, stretch_x=1.05))]
    for ours, theirs in zip(build(C), build(ref)):
        assert np.array_equal(ours.sdf(pts), theirs.sdf(pts)), type(theirs).__name__
        assert np.array_equal(ours.contains(pts), theirs.contains(pts))
        for g_ours, g_theirs in grids:
            v = theirs.voxelize(g_theirs)
            assert v.any() and np.array_equal(ours.voxelize(g_ours), v), type(theirs).__name__
        if hasattr(ours, "bounding_box"):
            assert all(np.array_equal(x, y) for x, y in zip(ours.bounding_box, theirs.bounding_box))
    with pytest.raises(ValueError, match="Box size must be positive"):
        C.Box(center=(0, 0, 0), size=(1, 0, 1))
    with pytest.raises(ValueError, match="Must provide either"):
        C.Box(center=(0, 0, 0))


@pytest.mark.skipif(not R.have_reference_package(), reason="reference sources not on this box")
def test_membrane_sources_match_reference_live():
    """Circular / rectangular membrane mirrors (membranes.py) against core/solver.py:210-755: masks, weights (bit for bit),
    coordinate helpers, cached Bessel zeros, the off-grid warning and the validation messages."""
    import warnings
    ref = R.load_reference_package()
    from strata_fdtd.core.solver import CircularMembraneSource as RC, RectangularMembraneSource as RR
    wave = sb.GaussianPulse(position=(0, 0, 0), frequency=200.0)
    grids = [(sb.UniformGrid((24, 30, 20), 2e-3), ref.UniformGrid((24, 30, 20), 2e-3)),
             (sb.NonuniformGrid.from_stretch((26, 22, 30), 2e-3, stretch_x=1.04, stretch_z=1.03),
              ref.NonuniformGrid.from_stretch((26, 22, 30), 2e-3, stretch_x=1.04, stretch_z=1.03))]
    for g_ours, g_ref in grids:
        ext = g_ours.physical_extent()
        c = (0.47 * ext[0], 0.52 * ext[1], 0.41 * ext[2])
        for axis in "xyz":
            for mode in ((0, 1), (1, 1), (2, 2), (3, 2), (0, 3)):
                kw = dict(center=c, normal_axis=axis, waveform=wave, mode=mode, radius=0.35 * min(ext), injection_type="velocity")
                a, b = sb.CircularMembraneSource(**kw), RC(**kw)
                assert a._alpha == b._alpha and a._BESSEL_ZEROS == b._BESSEL_ZEROS and a.source_type == b.source_type
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    assert np.array_equal(a.get_injection_mask(g_ours), b.get_injection_mask(g_ref))
                    wa, wb = a.get_injection_weights(g_ours), b.get_injection_weights(g_ref)
                assert wa.dtype == wb.dtype and np.array_equal(wa, wb) and wb.max() == 1.0, (axis, mode)
                for x, y in zip(a._grid_to_membrane_coords(g_ours), b._grid_to_membrane_coords(g_ref)):
                    assert np.array_equal(x, y)
            for mode in ((1, 1), (2, 1), (3, 4)):
                kw = dict(center=c, normal_axis=axis, waveform=wave, mode=mode, size=(0.5 * min(ext), 0.3 * min(ext)))
                a, b = sb.RectangularMembraneSource(**kw), RR(**kw)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    assert np.array_equal(a.get_injection_mask(g_ours), b.get_injection_mask(g_ref))
                    assert np.array_equal(a.get_injection_weights(g_ours), b.get_injection_weights(g_ref)), (axis, mode)
                for x, y in zip(a._grid_to_rectangular_coords(g_ours), b._grid_to_rectangular_coords(g_ref)):
                    assert np.array_equal(x, y)
    pts_r, pts_t = np.linspace(0, 0.03, 50), np.linspace(-3, 3, 50)
    kw = dict(center=(0.01, 0.01, 0.01), normal_axis="z", waveform=wave, mode=(1, 2), radius=0.02)
    assert np.array_equal(sb.CircularMembraneSource(**kw).mode_shape(pts_r, pts_t), RC(**kw).mode_shape(pts_r, pts_t))
    off = dict(center=(0.0201, 0.02, -0.0005), normal_axis="z", waveform=wave, radius=0.01)         # 0.75 cells below the first plane
    for cls in (sb.CircularMembraneSource, RC):
        with pytest.warns(UserWarning, match=r"is 0\.00\d+m \(0\.\d cells\) off-grid in z-direction"):
            cls(**off).get_injection_weights(grids[0][0])
    for bad, msg in ((dict(mode=(-1, 1)), "Azimuthal mode m must be non-negative"), (dict(mode=(0, 0)), "Radial mode n must be positive")):
        with pytest.raises(ValueError, match=msg):
            sb.CircularMembraneSource(center=c, normal_axis="x", waveform=wave, radius=0.01, **bad)
    with pytest.raises(ValueError, match="Mode index n must be >= 1"):
        sb.RectangularMembraneSource(center=c, normal_axis="x", waveform=wave, size=(0.01, 0.01), mode=(1, 0))


@pytest.mark.skipif(not R.have_reference_package(), reason="reference sources not on this box")
def test_audio_file_waveform_matches_reference_live(tmp_path):
    """waveforms.AudioFileWaveform against core/waveforms.py:33-231 on a stereo 16-bit file: samples, resampling to 1/dt,
    channel selection, trimming, looping and the zero tail -- identical arrays."""
    from scipy.io import wavfile
    R.load_reference_package()
    from strata_fdtd.core.waveforms import AudioFileWaveform as RefWave
    rate = 22050
    tt = np.arange(int(0.05 * rate)) / rate
    stereo = np.stack([np.sin(2 * np.pi * 440 * tt), 0.5 * np.sin(2 * np.pi * 1000 * tt + 0.3)], axis=1)
    path = tmp_path / "tone.wav"
    wavfile.write(path, rate, (stereo * 32767).astype(np.int16))
    dt = 1.6e-6
    t = np.arange(0, 40000) * dt
    for kw in (dict(), dict(channel=1, amplitude=0.25), dict(start_time=0.01, duration=0.02, loop=True), dict(channel=0, start_time=0.03)):
        a, b = sb.AudioFileWaveform(path, **kw), RefWave(path, **kw)
        assert a.native_sample_rate == b.native_sample_rate and a.num_samples == b.num_samples and a.duration_seconds == b.duration_seconds
        wa, wb = a.waveform(t, dt), b.waveform(t, dt)
        assert wa.dtype == wb.dtype and np.array_equal(wa, wb) and np.abs(wb).max() > 0, kw
    with pytest.raises(ValueError, match="Channel 2 requested but file only has 2 channels"):
        sb.AudioFileWaveform(path, channel=2)
    with pytest.raises(FileNotFoundError, match="Audio file not found"):
        sb.AudioFileWaveform(tmp_path / "missing.wav")


@pytest.mark.skipif(not R.have_reference_package(), reason="reference sources not on this box")
def test_microphone_host_sampling_matches_reference_live(tmp_path):
    """Microphone.record (the per-step Python path of core/solver.py:1004-1173) on random fields: omnidirectional, every named
    polar pattern and a callable pattern give the reference's samples exactly; to_wav writes the same bytes."""
    import types
    ref = R.load_reference_package()
    rng = np.random.default_rng(5)
    shape = (12, 10, 14)
    fields = [rng.standard_normal(shape).astype(np.float32) for _ in range(4)]
    host = types.SimpleNamespace(dx=2e-3, dt=3.1e-6, rho=1.2, c=343.0, shape=shape)
    for pattern in list(sb.POLAR_PATTERNS) + [lambda th: 0.3 + 0.7 * np.cos(th) ** 2]:
        kw = dict(position=(0.0113, 0.0087, 0.0141), name="m", pattern=pattern, direction=(0.3, -0.5, 0.8))
        a, b = sb.Microphone(**kw), ref.Microphone(**kw)
        a._initialize(host); b._initialize(host)
        for q in range(120):                                      # (to_wav at 48 kHz keeps 17 of them)
            scaled = [f * np.float32(np.sin(0.21 * q)) for f in fields]
            a.record(scaled[0], q * host.dt, *scaled[1:]); b.record(scaled[0], q * host.dt, *scaled[1:])
        assert np.array_equal(a.get_waveform(), b.get_waveform()) and np.array_equal(a.get_time_axis(), b.get_time_axis()), pattern
        assert repr(a) == repr(b) and len(a) == len(b) == 120
        for bits in (16, 32):
            a.to_wav(str(tmp_path / "a.wav"), sample_rate=48000, bit_depth=bits); b.to_wav(str(tmp_path / "b.wav"), sample_rate=48000, bit_depth=bits)
            assert (tmp_path / "a.wav").read_bytes() == (tmp_path / "b.wav").read_bytes()
    with pytest.raises(ValueError, match="Velocity fields \\(vx, vy, vz\\) are required"):
        m = sb.Microphone(position=(0.01, 0.01, 0.01), pattern="cardioid"); m._initialize(host); m.record(fields[0], 0.0)
    with pytest.raises(RuntimeError, match="Microphone not initialized"):
        sb.Microphone(position=(0.01, 0.01, 0.01)).record(fields[0], 0.0)


@pytest.mark.skipif(not R.have_reference_package(), reason="reference sources not on this box")
def test_shim_registers_backend_with_reference():
    ref = R.load_reference_package()
    sb.install_into_reference(ref)
    s = ref.FDTDSolver(shape=(12, 12, 12), resolution=1e-3, backend="b200")
    assert type(s).__module__.startswith("strata_fdtd_b200") and s.backend == "b200"
    assert s.grid.num_cells == 12 ** 3
    r = ref.FDTDSolver(shape=(12, 12, 12), resolution=1e-3, backend="native")
    assert r.using_native and r.grid.num_cells == 12 ** 3
    r.step()


def test_result_writer_schema_without_device(tmp_path):
    """The asynchronous writer (io.ResultWriter) against a stub solver: reference schema, chunked appends."""
    import json, types
    from strata_fdtd_b200 import io as sbio
    solver = types.SimpleNamespace(shape=(4, 5, 6), dx=1e-3, dt=1.6e-6, c=343.0, rho=1.2, step_count=7, time=7 * 1.6e-6,
                                   grid=sb.UniformGrid((4, 5, 6), 1e-3), geometry=np.ones((4, 5, 6), bool),
                                   _sources=[sb.GaussianPulse(position=(1, 2, 3), frequency=1e3)],
                                   _probes={"a": sb.Probe("a", (1, 1, 1)), "b": sb.Probe("b", (2, 2, 2))})
    out = tmp_path / "r.h5"
    w = sbio.ResultWriter(out, solver, script_content="print(1)")
    blk = np.arange(14, dtype=np.float32).reshape(7, 2)
    w.append_probe_block(["a", "b"], blk[:3]); w.append_probe_block(["a", "b"], blk[3:])
    w.write_snapshot(np.full((4, 5, 6), 2.0, np.float32))
    w.finalize(runtime=0.5, backend="b200", num_threads=0)
    if sbio.HAVE_H5PY:
        import h5py
        with h5py.File(out, "r") as f:
            assert np.array_equal(f["probes/b"][:], blk[:, 1]) and f["fields/pressure"].shape == (1, 4, 5, 6)
            assert f["metadata"].attrs["backend"] == "b200" and f["simulation"].attrs["num_steps"] == 7
    else:
        z = np.load(out)
        assert np.array_equal(z["probes/a"], blk[:, 0]) and np.array_equal(z["probes/b"], blk[:, 1])
        assert z["fields/pressure"].shape == (1, 4, 5, 6) and z["materials/geometry"].dtype == np.uint8
        attrs = json.loads(str(z["__attrs__"]))
        assert attrs["metadata@backend"] == "b200" and attrs["simulation@num_steps"] == 7
        assert attrs["grid@shape"] == [4, 5, 6] and attrs["sources/source_0@frequency"] == 1e3
    # the reference reader's interface (io/hdf5.py:243-375) over whichever container was written
    with sbio.HDF5ResultReader(out) as rd:
        meta = rd.get_metadata()
        assert meta["metadata"]["backend"] == "b200" and meta["simulation"]["num_steps"] == 7 and list(meta["grid"]["shape"]) == [4, 5, 6]
        assert meta["sources"][0]["frequency"] == 1e3 and list(meta["probes"]["b"]["position"]) == [2, 2, 2]
        assert rd.get_probe_names() == ["a", "b"] and np.array_equal(rd.load_probe("b"), blk[:, 1])
        assert rd.get_num_snapshots() == 1 and np.array_equal(rd.load_timestep(0), np.full((4, 5, 6), 2.0, np.float32))
        assert rd.load_geometry().dtype == bool and rd.load_geometry().all()
        with pytest.raises(KeyError, match="Probe 'zz' not found. Available: \\['a', 'b'\\]"):
            rd.load_probe("zz")


def test_strata_fdtd_alias_resolves_names():
    import subprocess, sys
    code = ("from strata_fdtd_b200.compat import install_as_strata_fdtd; install_as_strata_fdtd(force_alias=True); "
            "from strata_fdtd import FDTDSolver, PML, GaussianPulse, NonuniformGrid, Box, Difference, Sphere; "
            "from strata_fdtd.materials import Pole, PoleType, SimpleMaterial; "
            "from strata_fdtd.boundaries import RadiationImpedance; "
            "s = FDTDSolver(shape=(6,6,6), resolution=1e-3); print(s.backend, s.using_native)")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT))
    assert res.stdout.strip() == "b200 False", res.stderr


def test_slab_material_storage_keeps_the_ghost_planes():
    """Host logic of ADE on slabs, no device: a slab takes masks over the WHOLE grid and keeps the material ids of its
    owned planes plus the live ghost plane on each cut (the device list needs them for the velocity correction across
    the cut); a mask of the wrong shape is refused with the reference's message."""
    import strata_fdtd_b200 as sb
    shape = (12, 6, 5)
    mask = np.zeros(shape, dtype=bool)
    mask[3:9, 1:4, :] = True
    mat = sb.PoleMaterial("m", 1.2, 1.2 * 343.0 ** 2, [sb.Pole(sb.PoleType.DEBYE, 0.1, "density", tau=1e-4)])
    for lo, hi in ((0, 4), (4, 8), (8, 12)):
        s = sb.FDTDSolver(shape=shape, resolution=1e-3, slab=(lo, hi))
        s.register_material(mat, material_id=5)
        s.set_material_region(mask, material_id=5)
        glo, ghi = max(lo - 1, 0), min(hi + 1, shape[0])
        assert s._material_ext.shape == (ghi - glo,) + shape[1:]
        assert np.array_equal(s._material_ext == 5, mask[glo:ghi])
        assert np.array_equal(s._material_id == 5, mask[lo:hi])            # the owned planes are a view of the same storage
        with pytest.raises(ValueError, match="doesn't match solver shape"):
            s.set_material_region(mask[lo:hi], material_id=5)
        with pytest.raises(ValueError, match="not registered"):
            s.set_material_region(mask, material_id=9)


def test_chunk_length_is_the_callers_when_given():
    import strata_fdtd_b200 as sb
    assert sb.FDTDSolver(shape=(8, 8, 8), resolution=1e-3)._chunk_auto is True
    s = sb.FDTDSolver(shape=(8, 8, 8), resolution=1e-3, chunk_steps=37)
    assert s._chunk_auto is False and s._chunk_steps == 37


def test_no_fused_multiply_add_in_any_kernel():
    """The arithmetic contract: every fp32 / fp64 operation is separately rounded, as in the
    reference's -O3 build without -ffast-math.  The library is compiled --fmad=false; the SASS of all kernels (K0-K6,
    ADE, plane ops, energy) must therefore contain no FFMA / DFMA."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    _lib.load()
    sass = subprocess.run([cuobjdump, "-sass", str(_lib.SO_PATH)], capture_output=True, text=True, check=True).stdout
    kernels = re.findall(r"Function : (\S+)", sass)
    assert len(kernels) >= 60 and any("k5_resident" in k for k in kernels) and any("k6_pipeline" in k for k in kernels)
    fused = re.findall(r"\b(FFMA|DFMA)\b", sass)
    assert not fused, f"{len(fused)} fused multiply-adds in the device code"
