"""Seeded random configurations of the hot path against the CPU oracle, bit for bit.

The fixed cases of tests/cases.py pin particular features; this sweeps their combinations -- awkward extents (rows that end
inside a float4, single-cell axes, rows longer than one 128-cell strip), random solids touching the outer faces, sponge
layers on any subset of axes, uniform / nonuniform spacing, several sources and probes, a microphone, up to two dispersive
materials as random boxes (density / modulus, Debye / Lorentz poles) in every ADE device layout, and every step kernel
(streaming K1 in both plane mappings and both rows-per-thread, naive K0, shared-memory-resident K5, step-pipelined K6) --
so that an index or predicate slip in one variant cannot hide behind the shapes the hand-written cases happen to use.
"""
import numpy as np
import pytest

from cases import BENIGN_POLES, SECOND_POLES
from oracle import oracle as O
from strata_fdtd_b200 import _lib
from util import assert_same_as_oracle, build_b200_solver

pytestmark = pytest.mark.gpu


def _random_case(seed: int) -> tuple[dict, dict, int]:
    rng = np.random.default_rng(1000 + seed)
    big_k = seed % 5 == 0
    shape = (int(rng.integers(3, 34)), int(rng.integers(3, 30)), int(rng.integers(130, 170) if big_k else rng.integers(3, 45)))
    if seed % 7 == 3:
        shape = (shape[0], 1 + seed % 2, shape[2])                       # a (nearly) two-dimensional grid
    case: dict = dict(steps=int(rng.integers(25, 60)))
    if rng.random() < 0.3 and min(shape) >= 2:                          # (a nonuniform axis needs two points, grid.py)
        def stretched(n):
            sizes = 1e-3 * rng.uniform(1.0, 1.3, size=n)
            edges = np.concatenate([[0.0], np.cumsum(sizes)])
            return 0.5 * (edges[:-1] + edges[1:])
        case["nonuniform"] = dict(x_coords=stretched(shape[0]), y_coords=stretched(shape[1]), z_coords=stretched(shape[2]))
    else:
        case["shape"], case["resolution"] = shape, 1e-3
    if rng.random() < 0.6:
        g = np.ones(shape, dtype=bool)
        for _ in range(int(rng.integers(1, 4))):
            lo = [int(rng.integers(0, n)) for n in shape]
            hi = [min(n, l + int(rng.integers(1, max(2, n // 2 + 1)))) for l, n in zip(lo, shape)]
            g[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = False
        if g.any():
            case["geometry"] = g
    air = case.get("geometry", np.ones(shape, dtype=bool))
    pml = []
    for _ in range(int(rng.integers(0, 3))):
        axes = tuple(a for a, n in zip("xyz", shape) if rng.random() < 0.7 and n >= 6)
        if axes:
            depth = int(rng.integers(1, max(2, min(6, min(n for a, n in zip("xyz", shape) if a in axes) // 2))))
            pml.append(dict(depth=depth, axes=axes, order=int(rng.integers(2, 4))))
    case["pml"] = pml

    def cell():
        return tuple(int(rng.integers(0, n)) for n in shape)
    case["sources"] = [dict(kind="point", position=cell(), frequency=float(rng.uniform(8e3, 3e4)),
                            amplitude=float(rng.uniform(0.2, 2.0))) for _ in range(int(rng.integers(1, 4)))]
    if not any(air[s["position"]] for s in case["sources"]):
        free = np.argwhere(air)
        case["sources"][0]["position"] = tuple(int(q) for q in free[len(free) // 2])
    case["probes"] = [(f"p{q}", cell()) for q in range(int(rng.integers(1, 5)))] + [("at_source", case["sources"][0]["position"])]
    if case.get("nonuniform") is None and all(n >= 3 for n in shape) and rng.random() < 0.4:
        case["mics"] = [("m", tuple(float(rng.uniform(0.1, n - 1.2)) * 1e-3 for n in shape))]
    opts: dict = {}
    if rng.random() < 0.5:
        mid = np.zeros(shape, dtype=np.uint8)
        mats = []
        for mat_id, poles in ((3, BENIGN_POLES), (7, SECOND_POLES))[: int(rng.integers(1, 3))]:
            lo = [int(rng.integers(0, n)) for n in shape]
            hi = [min(n, l + int(rng.integers(1, n + 1))) for l, n in zip(lo, shape)]
            mid[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = mat_id
            keep = [p for p in poles if rng.random() < 0.8] or poles[:1]
            mats.append(dict(id=mat_id, rho_inf=float(rng.uniform(1.0, 1.6)), K_inf=float(rng.uniform(1.0, 1.6) * 343.0 ** 2), poles=keep))
        mats = [m for m in mats if (mid == m["id"]).any()]
        if mats:
            case["materials"], case["material_id"] = mats, mid
            opts[_lib.OPT_ADE_LAYOUT] = int(rng.integers(0, 4))
    kernels = [_lib.KERNEL_AUTO, _lib.KERNEL_MARCH, _lib.KERNEL_MARCH, _lib.KERNEL_NAIVE]
    if "materials" not in case and "mics" not in case:
        kernels += [_lib.KERNEL_RESIDENT, _lib.KERNEL_PIPELINE]
    opts[_lib.OPT_KERNEL] = kernels[int(rng.integers(0, len(kernels)))]
    if opts[_lib.OPT_KERNEL] == _lib.KERNEL_NAIVE and opts.get(_lib.OPT_ADE_LAYOUT) == 3:
        opts[_lib.OPT_ADE_LAYOUT] = 1                                       # the fused layout belongs to the marching kernel
    if opts[_lib.OPT_KERNEL] in (_lib.KERNEL_MARCH, _lib.KERNEL_AUTO, _lib.KERNEL_PIPELINE) and rng.random() < 0.7:
        opts[_lib.OPT_ROWS_PER_THREAD] = int(rng.integers(1, 3))
        opts[_lib.OPT_PLANE_MAP] = int(rng.integers(0, 3))
        opts[_lib.OPT_CHUNK_I] = int(rng.integers(1, 9))
        opts[_lib.OPT_WARPS_J] = int(rng.choice([1, 2, 4, 8]))
        opts[_lib.OPT_USE_GRAPH] = int(rng.integers(-1, 2))
    chunk = int(rng.choice([7, 16, 64]))
    # Mur / radiation planes (a stream of their own, so that the configurations above stay what they were): any subset of
    # faces in any order after the sponges -- edges and corners depend on that order
    rng2 = np.random.default_rng(7000 + seed)
    if rng2.random() < 0.4 and min(shape) >= 3:
        bcs = []
        for _ in range(int(rng2.integers(1, 4))):
            if rng2.random() < 0.5:
                axes = tuple(a for a in "xyz" if rng2.random() < 0.6) or ("y",)
                bcs.append(dict(kind="mur", axes=axes))
            else:
                bcs.append(dict(kind="radiation", axis="xyz"[int(rng2.integers(0, 3))], side=("low", "high")[int(rng2.integers(0, 2))],
                                **(dict(reflection_coeff=float(rng2.uniform(0.1, 0.9))) if rng2.random() < 0.5
                                   else dict(pipe_radius=float(rng2.uniform(0.004, 0.02))))))
        case["plane_bcs"] = bcs
    return case, opts, chunk


@pytest.mark.parametrize("seed", range(48))
def test_random_configuration_matches_oracle(seed):
    case, opts, chunk = _random_case(seed)
    s = build_b200_solver(case, chunk_steps=chunk)
    for k in (_lib.OPT_KERNEL, _lib.OPT_ADE_LAYOUT):                       # (the ADE layout is chosen when the materials are uploaded)
        if k in opts:
            s.set_kernel_option(k, opts[k])
    for k, v in opts.items():
        s.set_kernel_option(k, v)
    o = O.OracleSolver(case)
    steps = case["steps"]
    try:
        s.run(steps=steps)
    except _lib.B200BackendError as e:
        if "not applicable" in str(e) or "does not fit" in str(e) or "cannot be co-resident" in str(e):   # (incl. K5 box grids)
            pytest.skip(f"kernel variant refused this configuration: {e}")
        raise
    o.run_steps(steps)
    what = f"seed {seed}: shape {s.shape}, opts {opts}, chunk {chunk}, " \
           f"{'nonuniform ' if 'nonuniform' in case else ''}{'geometry ' if 'geometry' in case else ''}" \
           f"{len(case['pml'])} sponge(s), {len(case.get('materials', []))} material(s), planes {case.get('plane_bcs', [])}"
    assert_same_as_oracle(s, o, what)
    s.close()


@pytest.mark.parametrize("seed", range(100, 124))
def test_random_configuration_is_decomposition_invariant(seed):
    """The same random configurations cut into 2-4 slabs (device copies, the peer-store / flag protocol, or the
    cut-planes-first schedule of the overlapped NCCL mode) equal the single-domain run bit for bit: fields, probes and
    microphones whose corners may straddle a cut, with solids, sponges, nonuniform spacing and dispersive materials
    (list layouts; their ghost-plane cells are advanced redundantly) wherever the cuts happen to fall."""
    from strata_fdtd_b200.multi import slab_ranges
    from test_multi_gpu import _group_from_case
    case, opts, chunk = _random_case(seed)
    rng = np.random.default_rng(5000 + seed)
    nx = (case.get("shape") or (len(case["nonuniform"]["x_coords"]),))[0]
    n_slabs = int(rng.integers(2, min(4, nx) + 1))
    halo = ["copy", "p2p", "copy_cuts"][int(rng.integers(0, 3))]
    opts = {k: v for k, v in opts.items() if k in (_lib.OPT_ROWS_PER_THREAD, _lib.OPT_PLANE_MAP, _lib.OPT_CHUNK_I, _lib.OPT_WARPS_J)}
    if "materials" in case:
        opts[_lib.OPT_ADE_LAYOUT] = 1                                  # slabs run the compact list
    one = build_b200_solver(case, chunk_steps=chunk)
    for k, v in opts.items():
        one.set_kernel_option(k, v)
    grp = _group_from_case(case, n_slabs, opts, halo=halo)
    steps = case["steps"]
    one.run(steps=steps); grp.run(steps)
    what = f"seed {seed}: shape {one.shape}, {n_slabs} slabs {slab_ranges(nx, n_slabs)}, halo {halo}, opts {opts}"
    for f in ("p", "vx", "vy", "vz"):
        a, b = grp.get_field(f), one.get_field(f)
        assert np.array_equal(a, b), f"{what}: {f} differs (first at {np.argwhere(a != b)[:1]})"
    tr = grp.get_probe_data()
    for pname in one._probes:
        assert np.array_equal(tr[pname], one.get_probe_data(pname)[pname]), f"{what}: probe {pname}"
    for mname, mic in one.microphones.items():
        assert np.array_equal(grp.microphones[mname].get_waveform(), mic.get_waveform()), f"{what}: mic {mname}"
    grp.close(); one.close()
