"""CPU check of the shared-memory-resident step kernel's logic (strata_fdtd_b200/csrc/sb_resident.cuh).

The kernel's phase functions are __host__ __device__; tests/emu/k5_emu.cu runs them in lockstep for every
"thread" of every "CTA" on the host.  Here that emulation is fed the same padded arrays and tables the C ABI
uploads and must reproduce the oracle bit-for-bit -- box decomposition, halo indexing, redundant ghost faces,
deferred sponge, chunk boundaries.  (The device-only parts -- flags, barriers, co-residency -- are covered by
the -m gpu tests.)  Test infrastructure only; nothing in the package loads the emulation."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from cases import c1_case, make_cases
from oracle import oracle as O
from strata_fdtd_b200 import _lib
from strata_fdtd_b200.boundaries import sponge_tables
from util import build_b200_solver

HERE = Path(__file__).parent
SRC = HERE / "emu" / "k5_emu.cu"
SO = HERE / "emu" / "_build" / "libk5emu.so"
DEPS = [SRC, _lib._PKG / "csrc" / "sb_resident.cuh", _lib._PKG / "csrc" / "sb_kernels.cuh"]


@pytest.fixture(scope="module")
def emu():
    if not SO.exists() or any(d.stat().st_mtime > SO.stat().st_mtime for d in DEPS):
        SO.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run([_lib._nvcc(), "-O2", "--fmad=false", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", str(SO), str(SRC)], check=True)
    lib = C.CDLL(str(SO))
    lib.k5emu_run.restype = C.c_int
    return lib


def _mask_bytes(geom: np.ndarray, rigid: bool) -> np.ndarray:
    """bit0 air, bits1-3 +x/+y/+z face open: what k_build_mask derives on the device (boundaries.cpp:13-64)."""
    air = geom.astype(bool)
    m = air.astype(np.uint8)
    for axis, bit in ((0, 2), (1, 4), (2, 8)):
        nxt = np.ones_like(air)
        sl_lo = [slice(None)] * 3; sl_hi = [slice(None)] * 3
        sl_lo[axis] = slice(0, -1); sl_hi[axis] = slice(1, None)
        nxt[tuple(sl_lo)] = air[tuple(sl_hi)]
        m |= (np.where((air & nxt) | (not rigid), bit, 0)).astype(np.uint8)
    return m


def _pad(a, n, fill, lead=0):
    t = np.full(n, fill, dtype=np.float32)
    t[lead:lead + len(a)] = a
    return t


def run_emulated(emu, case, n_steps, chunk=37, nbi=0, nbj=0, n_sm=148, smem_limit=227 * 1024, split=1):
    s = build_b200_solver(case)
    nx, ny, nz = s.shape
    pitch = C.c_int32(0)
    _lib.check(_lib.load().sb_choose_pitch(nz, C.byref(pitch)))           # the library's row pitch
    pitch = pitch.value
    faces, cells, cp = s._coefficient_tables()
    cvx, cvy, cvz = _pad(faces[0], nx + 2, 0.0, 1), _pad(faces[1], ny + 4, 0.0), _pad(faces[2], pitch + 4, 0.0)
    nu = cells[0] is not None
    ic = [_pad(cells[0], nx + 2, 1.0, 1), _pad(cells[1], ny + 4, 1.0), _pad(cells[2], pitch + 4, 1.0)] if nu else [None] * 3
    dec = []
    for b in s._boundaries:
        t = sponge_tables(b, s)
        assert t is not None
        one = lambda a, n: np.ones(n, np.float32) if a is None else a
        dec.append((_pad(one(t[0], nx), nx + 2, 1.0, 1), _pad(one(t[1], ny), ny + 4, 1.0), _pad(one(t[2], nz), pitch + 4, 1.0)))
    mask = None
    if s._geometry is not None and not s._geometry.all():
        mask = np.zeros((nx + 2, ny, pitch), dtype=np.uint8)
        mask[1:nx + 1, :, :nz] = _mask_bytes(s._geometry, s._rigid)
    cell_idx, start, sids, flds, wts = s._build_source_table()
    assert len(sids) <= 32 and not np.any(flds), "emulated cases use point sources into p"
    ijks, w = [], []
    for u, c in enumerate(cell_idx):
        for e in range(start[u], start[u + 1]):
            ijks += [int(c // (ny * nz)), int((c // nz) % ny), int(c % nz), int(sids[e])]
            w.append(float(wts[e]))
    ijks = np.array(ijks + [0], dtype=np.int32); w = np.array(w + [0.0], dtype=np.float64)
    names = list(s._probes)
    pijk = np.array([q for n in names for q in s._probes[n].position] + [0], dtype=np.int32)
    n_src, n_rec = max(1, len(s._sources)), max(1, len(names))
    fields = [np.zeros((nx + 2, ny, pitch), dtype=np.float32) for _ in range(8)]
    fptr = (C.c_void_p * 8)(*[f.ctypes.data for f in fields])
    arr = lambda tabs: (C.c_void_p * max(1, len(tabs)))(*[t.ctypes.data for t in tabs])
    dx_, dy_, dz_ = arr([d[0] for d in dec]), arr([d[1] for d in dec]), arr([d[2] for d in dec])
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    traces, cur, done, t, chosen = [], 0, 0, 0.0, (C.c_int * 2)()
    while done < n_steps:
        m = min(chunk, n_steps - done)
        times = np.empty(m)
        for q in range(m):
            times[q] = t; t = t + s.dt
        W = np.ascontiguousarray(s._waveform_table(times))
        rec = np.zeros((m, n_rec), dtype=np.float32)
        rc = emu.k5emu_run(nx, ny, nz, pitch, fptr, cur, m, vp(mask), vp(cvx), vp(cvy), vp(cvz), vp(ic[0]), vp(ic[1]), vp(ic[2]),
                           len(dec), dx_, dy_, dz_, C.c_float(float(cp)), len(w) - 1, vp(ijks), vp(w), vp(W), n_src,
                           len(names), vp(pijk), vp(rec), n_rec, nbi, nbj, n_sm, C.c_longlong(smem_limit), split, chosen)
        assert rc != 2, "a box received a face with the wrong step tag"
        if rc != 0 and nbi > 0:
            pytest.skip("this forced partition needs more than 512 columns per box")
        assert rc == 0, "grid does not fit the emulated machine"
        traces.append(rec); cur = (cur + m) & 1; done += m
    rec = np.concatenate(traces)
    out = {f: fields[4 * cur + q][1:nx + 1, :, :nz] for q, f in enumerate(("p", "vx", "vy", "vz"))}
    pads = [fields[4 * cur + q][1:nx + 1, :, nz:] for q in range(4)]
    return out, {n: rec[:, q] for q, n in enumerate(names)}, tuple(chosen), pads


def _strip(case, **over):
    c = {k: v for k, v in case.items() if k not in ("mics",)}
    c.update(over)
    return c


CASES = make_cases()
EMU_CASES = {
    "odd_rigid_box": CASES["odd_rigid_box"],
    "block_pml": CASES["block_pml"],
    "uniform_pml": _strip(CASES["uniform_pml"]),
    "nonuniform_block_pml": _strip(CASES["nonuniform_block_pml"]),
    "odd_geometry_pml_two_sources": _strip(CASES["odd_geometry_pml"]),
    "two_sponges": _strip(CASES["partial_pml_plane"], sources=[CASES["partial_pml_plane"]["sources"][1],
                                                               dict(kind="point", position=(8, 5, 5), frequency=15e3, amplitude=0.5)]),
}
# (nbi, nbj): automatic for 148 SMs, a tiny machine, extreme aspect ratios, one box
PARTITIONS = [(0, 0), (3, 2), (1, 5), (7, 2), (1, 1)]


@pytest.mark.parametrize("part", PARTITIONS)
@pytest.mark.parametrize("name", sorted(EMU_CASES))
def test_emulated_resident_kernel_matches_oracle(emu, name, part):
    case = EMU_CASES[name]
    steps = min(case["steps"], 80)
    out, traces, chosen, pads = run_emulated(emu, case, steps, chunk=37, nbi=part[0], nbj=part[1], split=(part[0] + part[1]) % 2)
    o = O.OracleSolver(case)
    o.run_steps(steps)
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(out[f], getattr(o, f)), f"{name} {chosen}: {f} differs at {np.argwhere(out[f] != getattr(o, f))[:3]}"
    for n, tr in traces.items():
        assert np.array_equal(tr, o.probe_array(n)), f"{name} {chosen}: probe {n}"
    assert all(not p.any() for p in pads), "row padding must stay zero"
    assert np.abs(out["p"]).max() > 0


def test_partition_of_config_1_fits_148_sms(emu):
    """100^3 (BASELINE config 1) on 148 SMs x 227 KB; 200 steps against the oracle through the automatic partition."""
    case = c1_case(0)
    case["sources"][0]["frequency"] = 40e3                    # a pulse that crosses boxes within the test's steps
    out, traces, chosen, _ = run_emulated(emu, case, 60, chunk=60)
    assert chosen[0] * chosen[1] <= 148 and chosen[0] > 1 and chosen[1] > 1
    o = O.OracleSolver(case)
    o.run_steps(60)
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(out[f], getattr(o, f)), f
    assert np.array_equal(traces["probe"], o.probe_array("probe"))


def test_grid_too_large_for_shared_memory_is_refused(emu):
    case = dict(shape=(160, 160, 160), resolution=1e-3, steps=0)
    with pytest.raises(AssertionError, match="does not fit"):
        run_emulated(emu, case, 1)
