"""Shared parity cases: one plain-dict description from which the oracle, the real
reference (oracle/ref_loader.build_reference_solver) and the b200 backend
(tests/util.build_b200_solver) are all constructed.

Everything is deterministic (fields start at zero; geometry / material masks come
from closed-form rules), so no RNG state has to travel with the fixtures.
"""
from __future__ import annotations

import numpy as np


def _block_geometry(shape, lo, hi):
    g = np.ones(shape, dtype=bool)
    g[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = False
    return g


def _sphere_mask(shape, centre, radius):
    i, j, k = np.ogrid[:shape[0], :shape[1], :shape[2]]
    return ((i - centre[0]) ** 2 + (j - centre[1]) ** 2 + (k - centre[2]) ** 2) < radius ** 2


def _stretched(n, base, ratio):
    """Cell-centre coordinates with geometric growth away from the centre (own construction)."""
    half = n // 2
    sizes_r = base * ratio ** np.arange(n - half, dtype=np.float64)
    sizes_l = base * ratio ** np.arange(half, dtype=np.float64)
    sizes = np.concatenate([sizes_l[::-1], sizes_r])
    edges = np.concatenate([[0.0], np.cumsum(sizes)])
    return 0.5 * (edges[:-1] + edges[1:])


from strata_fdtd_b200.workloads import (BENIGN_POLES, SECOND_POLES, c1_case, c2_case, c3_case, c4_case,  # noqa: E402,F401
                                        enclosure_air_mask)


def make_cases() -> dict:
    C = {}

    # --- uniform, sponge on all axes, point source, probes + omni mics
    C["uniform_pml"] = dict(
        shape=(24, 20, 28), resolution=1e-3, steps=240,
        pml=[dict(depth=5)],
        sources=[dict(kind="point", position=(6, 10, 14), frequency=20e3)],
        probes=[("a", (18, 10, 14)), ("b", (12, 4, 20)), ("at_source", (6, 10, 14))],
        mics=[("m0", (0.0123, 0.0101, 0.0137)), ("m1", (0.0031, 0.0152, 0.0209))],
    )

    # --- odd extents (nz not a multiple of 4), no sponge: rigid box, positions in metres
    C["odd_rigid_box"] = dict(
        shape=(17, 23, 13), resolution=2e-3, steps=200,
        sources=[dict(kind="point", position=(0.010, 0.020, 0.012), frequency=8e3, amplitude=2.5)],
        probes=[("far", (0.030, 0.040, 0.020)), ("idx", (3, 3, 3))],
    )

    # --- the bit-exactness configuration of SURVEY.md F6: solid block + sponge
    C["block_pml"] = dict(
        shape=(48, 40, 56), resolution=1e-3, steps=300,
        geometry=_block_geometry((48, 40, 56), (20, 14, 22), (30, 26, 36)),
        pml=[dict(depth=8)],
        sources=[dict(kind="point", position=(10, 20, 28), frequency=20e3)],
        probes=[("shadow", (40, 20, 28)), ("side", (24, 6, 28))],
    )

    # --- sponge on a subset of axes, two boundary objects, plane source, explicit max_sigma
    C["partial_pml_plane"] = dict(
        shape=(30, 22, 26), resolution=1e-3, steps=200,
        geometry=_block_geometry((30, 22, 26), (14, 0, 10), (18, 8, 16)),
        pml=[dict(depth=6, axes=("x", "z")), dict(depth=4, axes=("y",), order=2, max_sigma=4.0e5)],
        sources=[dict(kind="plane", axis=0, index=8, frequency=15e3, amplitude=0.5),
                 dict(kind="point", position=(22, 11, 13), frequency=12e3)],
        probes=[("p0", (25, 11, 13)), ("on_plane", (8, 5, 5))],
    )

    # --- nonuniform grid + geometry + sponge
    xs, ys, zs = _stretched(26, 1e-3, 1.04), _stretched(22, 1e-3, 1.0), _stretched(30, 1e-3, 1.03)
    C["nonuniform_block_pml"] = dict(
        nonuniform=dict(x_coords=xs, y_coords=ys, z_coords=zs), steps=240,
        geometry=_block_geometry((26, 22, 30), (11, 8, 12), (15, 14, 18)),
        pml=[dict(depth=5)],
        sources=[dict(kind="point", position=(6, 11, 15), frequency=18e3)],
        probes=[("q", (20, 11, 15)), ("r", (13, 4, 25))],
        mics=[("nm", (0.0102, 0.0098, 0.0131))],
    )

    # --- ADE: material sphere (2 Debye + 1 Lorentz), sponge, uniform
    sh = (32, 32, 32)
    mid = np.zeros(sh, dtype=np.uint8)
    mid[_sphere_mask(sh, (20, 16, 16), 6.5)] = 1
    C["ade_sphere"] = dict(
        shape=sh, resolution=1e-3, steps=300,
        pml=[dict(depth=6)],
        sources=[dict(kind="point", position=(8, 16, 16), frequency=20e3)],
        probes=[("behind", (28, 16, 16)), ("inside", (20, 16, 16)), ("front", (12, 16, 16))],
        materials=[dict(id=1, rho_inf=1.2, K_inf=1.2 * 343.0 ** 2, poles=BENIGN_POLES)],
        material_id=mid,
    )

    # --- ADE: two materials (one with density Lorentz), touching each other and a solid, on a nonuniform grid
    xs2, ys2, zs2 = _stretched(28, 1e-3, 1.02), _stretched(24, 1e-3, 1.03), _stretched(26, 1e-3, 1.0)
    sh2 = (28, 24, 26)
    mid2 = np.zeros(sh2, dtype=np.uint8)
    mid2[16:24, 4:20, 5:21] = 3
    mid2[12:16, 6:18, 8:18] = 7
    g2 = _block_geometry(sh2, (18, 10, 10), (21, 14, 16))        # solid inside material 3
    C["ade_two_materials_nonuniform"] = dict(
        nonuniform=dict(x_coords=xs2, y_coords=ys2, z_coords=zs2), steps=260,
        geometry=g2,
        pml=[dict(depth=4)],
        sources=[dict(kind="point", position=(6, 12, 13), frequency=16e3)],
        probes=[("m3", (17, 12, 13)), ("m7", (13, 12, 13)), ("air", (8, 6, 6))],
        materials=[dict(id=3, rho_inf=1.5, K_inf=1.5 * 320.0 ** 2, poles=BENIGN_POLES),
                   dict(id=7, rho_inf=1.1, K_inf=1.1 * 350.0 ** 2, poles=SECOND_POLES)],
        material_id=mid2,
    )
    # --- ADE: two materials stacked so that they fill most of their bounding box (-> the dense device layout), up to
    #     the outer faces of the grid, with an air pocket and a solid inside the first one
    sh3 = (30, 22, 26)
    mid3 = np.zeros(sh3, dtype=np.uint8)
    mid3[16:23, :, :] = 3
    mid3[23:30, :, :] = 7
    mid3[18:20, 8:12, 8:12] = 0                                   # air pocket
    g3 = _block_geometry(sh3, (20, 3, 14), (22, 7, 20))           # solid inside material 3
    C["ade_dense_layers"] = dict(
        shape=sh3, resolution=1e-3, steps=260,
        geometry=g3,
        pml=[dict(depth=4)],
        sources=[dict(kind="point", position=(6, 11, 13), frequency=16e3)],
        probes=[("m3", (19, 15, 13)), ("m7", (26, 11, 13)), ("pocket", (18, 9, 9)), ("air", (8, 6, 6)), ("face", (29, 21, 25))],
        materials=[dict(id=3, rho_inf=1.5, K_inf=1.5 * 320.0 ** 2, poles=BENIGN_POLES),
                   dict(id=7, rho_inf=1.1, K_inf=1.1 * 350.0 ** 2, poles=SECOND_POLES)],
        material_id=mid3,
    )
    # --- first-order Mur ABC on every face (edges/corners depend on the x,y,z application order)
    C["mur_all"] = dict(
        shape=(22, 26, 24), resolution=1e-3, steps=200,
        plane_bcs=[dict(kind="mur", axes=("x", "y", "z"))],
        sources=[dict(kind="point", position=(7, 13, 12), frequency=20e3)],
        probes=[("corner", (0, 0, 0)), ("edge", (21, 13, 0)), ("face", (0, 13, 12)), ("mid", (15, 13, 12))],
    )

    # --- open pipe: sponge on y/z first, then radiation impedance at both x ends (constant R and
    #     pipe-radius R) and a Mur plane pair on y; solid block inside
    C["pml_radiation_mur"] = dict(
        shape=(40, 18, 20), resolution=1e-3, steps=220,
        geometry=_block_geometry((40, 18, 20), (18, 0, 0), (22, 7, 20)),
        pml=[dict(depth=4, axes=("z",))],
        plane_bcs=[dict(kind="radiation", axis="x", side="high", reflection_coeff=0.7),
                   dict(kind="radiation", axis="x", side="low", pipe_radius=0.01),
                   dict(kind="mur", axes=("y",))],
        sources=[dict(kind="point", position=(8, 9, 10), frequency=15e3)],
        probes=[("hi_end", (39, 9, 10)), ("lo_end", (0, 9, 10)), ("ywall", (30, 0, 10)), ("mid", (30, 9, 10))],
    )
    # --- odd extents WITH geometry and sponge: rows end inside a float4 / mask word, solids touch the outer faces
    g_odd = _block_geometry((19, 21, 13), (7, 0, 5), (12, 9, 13))
    g_odd[0:3, 15:21, 0:4] = False
    g_odd[18, :, 6] = False
    C["odd_geometry_pml"] = dict(
        shape=(19, 21, 13), resolution=1e-3, steps=220,
        geometry=g_odd,
        pml=[dict(depth=3)],
        sources=[dict(kind="point", position=(4, 12, 8), frequency=20e3), dict(kind="point", position=(9, 4, 9), frequency=9e3)],
        probes=[("a", (15, 15, 3)), ("in_solid", (9, 4, 9)), ("edge", (18, 20, 12))],
        mics=[("m", (0.0141, 0.0122, 0.0031))],
    )

    # --- directional microphones (cardioid / figure-8 / custom gain(theta)) next to an omni one: the reference
    #     then records every microphone through its Python path (solver.py:2453-2461)
    C["directional_mics"] = dict(
        shape=(26, 24, 22), resolution=1e-3, steps=180,
        geometry=_block_geometry((26, 24, 22), (15, 0, 0), (18, 9, 22)),
        pml=[dict(depth=4)],
        sources=[dict(kind="point", position=(6, 12, 11), frequency=16e3)],
        probes=[("p", (20, 12, 11))],
        mics=[("omni", (0.0123, 0.0101, 0.0107)),
              ("card", (0.0123, 0.0101, 0.0107), dict(pattern="cardioid", direction=(1.0, 0.0, 0.0))),
              ("fig8", (0.0201, 0.0152, 0.0093), dict(pattern="figure8", direction=(0.0, 1.0, 1.0))),
              ("hyper_edge", (0.0003, 0.0004, 0.0208), dict(pattern="hypercardioid", direction=(-1.0, 0.5, 0.0))),
              ("custom", (0.0180, 0.0120, 0.0110), dict(pattern=_subcardioid_gain, direction=(0.0, 0.0, 1.0), up=(0.0, 1.0, 0.0)))],
    )
    return C


def _subcardioid_gain(theta):
    return 0.7 + 0.3 * np.cos(theta)


MEMBRANE_SPEC = dict(
    shape=(28, 24, 26), resolution=1e-3, steps=220,
    geometry_block=((12, 0, 0), (16, 6, 26)),
    pml=[dict(depth=4)],
    membranes=[
        dict(shape="circular", center=(0.020, 0.012, 0.008), radius=0.005, normal_axis="z", mode=(0, 1),
             injection_type="pressure", frequency=18e3, amplitude=1.0),
        dict(shape="rectangular", center=(0.006, 0.014, 0.016), size=(0.008, 0.006), normal_axis="x", mode=(1, 1),
             injection_type="velocity", frequency=14e3, amplitude=1e-3),
    ],
    probes=[("a", (20, 12, 16)), ("b", (8, 14, 16)), ("c", (3, 3, 3))],
)


def membrane_case(weights: list) -> dict:
    """The membrane case with the injection-weight arrays (taken from the reference via the fixture)."""
    sp = MEMBRANE_SPEC
    lo, hi = sp["geometry_block"]
    srcs = []
    for m, w in zip(sp["membranes"], weights):
        fld = "p" if m["injection_type"] == "pressure" else "v" + m["normal_axis"]
        srcs.append(dict(kind="weighted", weights=np.asarray(w, dtype=np.float64), field=fld, spec=m,
                         frequency=m["frequency"], amplitude=m["amplitude"]))
    return dict(shape=sp["shape"], resolution=sp["resolution"], steps=sp["steps"],
                geometry=_block_geometry(sp["shape"], lo, hi), pml=sp["pml"], sources=srcs, probes=sp["probes"])
