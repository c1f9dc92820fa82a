#!/usr/bin/env python
"""Config-size golden fixtures from the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY; build container only).

SURVEY.md 8(d) parity gates that tests/golden/*.npz of make_golden.py do not reach:

    python oracle/make_golden_large.py c4mask 256 128 128      # the reference's CSG enclosure, voxelised in x-chunks
    python oracle/make_golden_large.py c4mask 1024 512 512
    python oracle/make_golden_large.py c4 256 128 128 1000     # reference native backend on that geometry
    python oracle/make_golden_large.py c4 1024 512 512 1000    # full size once (268 M cells, ~20 min on 8 cores)
    python oracle/make_golden_large.py c3 512 1000             # fixed-native ADE harness at 512^3 (~40 min)

* c4mask: ``create_ported_enclosure()`` of /root/reference/examples/sdf_csg/ported_enclosure.py:34-121 (imported
  with its two stale module names aliased), scaled per axis so that its bounding box fills the central
  60 % of the domain and voxelised exactly as ``SDFPrimitive.voxelize`` does (geometry/sdf.py:99-125: cell-centre
  points, ``sdf <= 0``), but over chunks of x-planes so that the 1024 x 512 x 512 grid never needs the 6.4 GB
  point array.  Solver geometry = ~voxelize (True = air).  Stored bit-packed.
* c4 / c3: every probe trace in full, SHA-256 of the four final fields, a strided sample of p for diagnosis.
"""
from __future__ import annotations

import hashlib
import importlib.util
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import ref_loader as R  # noqa: E402
from strata_fdtd_b200 import workloads as W  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def reference_enclosure():
    """The CSG tree of the reference's example, untouched."""
    sf = R.load_reference_package()
    import strata_fdtd.core.grid as _grid
    import strata_fdtd.geometry.sdf as _sdf
    sys.modules.setdefault("strata_fdtd.grid", _grid)          # the example's stale import names
    sys.modules.setdefault("strata_fdtd.sdf", _sdf)
    path = R.REF_ROOT / "examples" / "sdf_csg" / "ported_enclosure.py"
    spec = importlib.util.spec_from_file_location("_ref_ported_enclosure", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        geometry, _grid_unused = mod.create_ported_enclosure()
    return sf, geometry


def c4_mask(shape):
    sf, enc = reference_enclosure()
    case = W.c4_case(shape, steps=0, enclosure="closed_form", materialise=False)     # only for the grid coordinates
    nu = case["nonuniform"]
    grid = sf.NonuniformGrid(x_coords=nu["x_coords"], y_coords=nu["y_coords"], z_coords=nu["z_coords"])
    extent = np.asarray(grid.physical_extent(), dtype=np.float64)
    bb_min, bb_max = (np.asarray(q, dtype=np.float64) for q in enc.bounding_box)
    size = bb_max - bb_min
    factor = 0.6 * extent / size                                  # bounding box -> central 60 % of every axis
    placed = enc.scale(tuple(factor)).translate(tuple(0.2 * extent - bb_min * factor))
    nx, ny, nz = shape
    air = np.empty(shape, dtype=bool)
    Y, Z = np.meshgrid(grid.y_coords, grid.z_coords, indexing="ij")
    yz = np.stack([Y.ravel(), Z.ravel()], axis=1)
    chunk = max(1, (1 << 22) // (ny * nz))
    for a in range(0, nx, chunk):
        b = min(nx, a + chunk)
        pts = np.empty(((b - a) * ny * nz, 3), dtype=np.float64)
        pts[:, 0] = np.repeat(grid.x_coords[a:b], ny * nz)
        pts[:, 1:] = np.tile(yz, (b - a, 1))
        air[a:b] = ~(placed.sdf(pts).reshape(b - a, ny, nz) <= 0)   # voxelize() == (sdf <= 0); geometry = ~voxelize
    if int(np.prod(shape)) <= 5_000_000:                            # small enough: prove chunking == voxelize()
        assert np.array_equal(air, ~placed.voxelize(grid)), "chunked voxelisation differs from SDFPrimitive.voxelize"
    out = W.reference_enclosure_mask_path(shape)
    np.savez_compressed(out, bits=np.packbits(air.ravel()), shape=np.array(shape, dtype=np.int64),
                        sha=np.array(digest(air)), scale=factor, solid_cells=np.int64((~air).sum()))
    print(f"c4mask {shape}: {int((~air).sum())} solid cells, sha {digest(air)[:16]}, {out.stat().st_size / 1024:.0f} KiB")


def _finish(s, out, steps, t0, path):
    for pname, probe in s._probes.items():
        out["probe_" + pname] = probe.get_data()
    for f in ("p", "vx", "vy", "vz"):
        arr = getattr(s, f)
        out["sha_" + f] = np.array(digest(arr))
        out["absmax_" + f] = np.float32(np.abs(arr).max())
    st = max(1, s.shape[0] // 32)
    out["sample_stride"] = np.int64(st)
    out["sample_p"] = s.p[::st, ::st, ::st].copy()
    out["dt"] = np.float64(s.dt); out["steps"] = np.int64(steps)
    np.savez_compressed(path, **out)
    print(f"{path.name}: {steps} steps in {time.time() - t0:.0f} s, |p|max {float(np.abs(s.p).max()):.3e}, "
          f"{path.stat().st_size / 1024:.0f} KiB")


def run_steps(s, steps, label):
    t0 = time.time()
    for n in range(steps):
        s.step()
        if n % 50 == 49:
            print(f"  {label}: step {n + 1}/{steps}, {time.time() - t0:.0f} s", flush=True)


def c4_run(shape, steps):
    case = W.c4_case(shape, steps=steps, enclosure="reference")
    t0 = time.time()
    s = R.build_reference_solver(case)
    assert s.using_native
    run_steps(s, steps, f"c4 {shape}")
    out = {"mask_sha": np.array(digest(np.asarray(case["geometry"], dtype=bool)))}
    _finish(s, out, steps, t0, ROOT / "tests" / "golden" / "c4_enclosure_{}x{}x{}.npz".format(*shape))


def c3_run(n, steps):
    case = W.c3_case(n, steps=steps)
    t0 = time.time()
    s = R.build_reference_solver(case)            # fixed-native ADE arrangement
    run_steps(s, steps, f"c3 {n}^3 ade")
    out = {"material_cells": np.int64((np.asarray(case["material_id"]) != 0).sum())}
    _finish(s, out, steps, t0, ROOT / "tests" / "golden" / f"c3_{n}_ade.npz")


def main():
    cmd = sys.argv[1]
    if cmd == "c4mask":
        c4_mask(tuple(int(q) for q in sys.argv[2:5]))
    elif cmd == "c4":
        c4_run(tuple(int(q) for q in sys.argv[2:5]), int(sys.argv[5]))
    elif cmd == "c3":
        c3_run(int(sys.argv[2]), int(sys.argv[3]))
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
