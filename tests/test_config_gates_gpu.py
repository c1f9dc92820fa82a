"""SURVEY.md 8(d) parity gates at configuration size, against fixtures produced by the UNMODIFIED reference in the build
container (oracle/make_golden_large.py): every probe trace in full and the SHA-256 of the four final fields.

* config 3: 512^3 + PML + ADE sphere (2 Debye + 1 Lorentz), 64 probes, 1000 steps -- the reference's own ade.cpp kernels in
  the order of core/solver.py:2135-2193 (fixed-native harness); every ADE device layout.
* config 4: nonuniform grid + the reference's own CSG ported enclosure (examples/sdf_csg/ported_enclosure.py:34-121
  voxelised by geometry/sdf.py:99-125) + PML, 8 probes, 1000 steps; 256 x 128 x 128 and 1024 x 512 x 512, on 1, 2 and 4 slabs.
"""
from pathlib import Path

import numpy as np
import pytest

from strata_fdtd_b200 import _lib
from strata_fdtd_b200.workloads import build_solver, c3_case, c4_case, load_reference_enclosure_mask
from test_multi_gpu import _group_from_case
from util import sha

GOLDEN = Path(__file__).parent / "golden"


def _check_against_fixture(get_field, traces, g, what):
    for name, tr in traces.items():
        want = g["probe_" + name]
        assert np.array_equal(tr, want), f"{what}: probe {name} differs (max|d| {np.abs(tr - want).max():.3e})"
    assert any(np.abs(tr).max() > 0 for tr in traces.values())
    for f in ("p", "vx", "vy", "vz"):
        a = get_field(f)
        assert float(np.abs(a).max()) == float(g["absmax_" + f]), f"{what}: |{f}|max differs"
        if f == "p":
            st = int(g["sample_stride"])
            assert np.array_equal(a[::st, ::st, ::st], g["sample_p"]), f"{what}: strided sample of p differs"
        assert sha(a) == str(g["sha_" + f]), f"{what}: final {f} != reference (SHA-256)"


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["auto", "fused", "compact", "dense"])
def test_c3_512_ade_sphere_1000_steps_equals_the_reference_harness(layout):
    g = np.load(GOLDEN / "c3_512_ade.npz")
    steps = int(g["steps"])
    case = c3_case(512, steps=steps)
    assert int((np.asarray(case["material_id"]) != 0).sum()) == int(g["material_cells"])
    s = build_solver(case)
    assert float(s.dt) == float(g["dt"])
    s.set_kernel_option(_lib.OPT_ADE_LAYOUT, {"auto": 0, "fused": 3, "compact": 1, "dense": 2}[layout])
    s.run(steps=steps)
    _check_against_fixture(s.get_field, {n: s.get_probe_data(n)[n] for n in s._probes}, g, f"c3 512^3 ({layout})")
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(256, 128, 128), (1024, 512, 512)])
def test_c4_reference_enclosure_equals_the_reference(shape):
    name = "c4_enclosure_{}x{}x{}.npz".format(*shape)
    if not (GOLDEN / name).exists():
        pytest.skip(f"{name} not generated")
    g = np.load(GOLDEN / name)
    steps = int(g["steps"])
    mask = load_reference_enclosure_mask(shape)
    assert sha(mask) == str(g["mask_sha"]), "the committed mask is not the one the reference was run on"
    case = c4_case(shape, steps=steps, enclosure="reference", materialise=False)
    one = build_solver(case)
    assert float(one.dt) == float(g["dt"])
    one.run(steps=steps)
    _check_against_fixture(one.get_field, {n: one.get_probe_data(n)[n] for n in one._probes}, g, f"c4 {shape} 1 slab")
    one.close()
    for n_slabs in (2, 4):
        grp = _group_from_case(case, n_slabs, {}, halo="p2p")
        grp.run(steps)
        _check_against_fixture(grp.get_field, grp.get_probe_data(), g, f"c4 {shape} {n_slabs} slabs")
        grp.close()


def test_sdf_objects_are_voxelised_per_slab_like_the_whole_grid():
    """set_geometry(<SDF object>) evaluates the object's sdf() over the planes a slab holds (a few planes at a time) and
    equals geometry/sdf.py:99-125's whole-grid voxelize() restricted to them."""
    import strata_fdtd_b200 as sb

    class Ball:                                   # anything with .sdf(points) -> signed distance (SDFPrimitive protocol)
        def sdf(self, pts):
            return np.linalg.norm(pts - np.array([0.012, 0.010, 0.011]), axis=1) - 0.006

        def voxelize(self, grid):
            X, Y, Z = np.meshgrid(grid.x_coords, grid.y_coords, grid.z_coords, indexing="ij")
            return (self.sdf(np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)) <= 0).reshape(grid.shape)

    shape = (24, 20, 22)
    whole = sb.FDTDSolver(shape=shape, resolution=1e-3)
    want = Ball().voxelize(whole.grid)
    whole.set_geometry(Ball())
    assert np.array_equal(whole.geometry, want) and 0 < want.sum() < want.size
    for lo, hi in ((0, 9), (9, 17), (17, 24)):
        s = sb.FDTDSolver(shape=shape, resolution=1e-3, slab=(lo, hi))
        s.set_geometry(Ball())
        assert np.array_equal(s.geometry, want[lo:hi])
        assert np.array_equal(s._geometry_ext, want[max(lo - 1, 0):min(hi + 1, shape[0])])
