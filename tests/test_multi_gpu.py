"""Slab decomposition: an N-slab run must be bit-identical to the single-slab run."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from cases import make_cases
from strata_fdtd_b200 import _lib
from strata_fdtd_b200.multi import LocalSlabGroup, owner_of, slab_ranges
from util import build_b200_solver

ROOT = Path(__file__).resolve().parents[1]
CASES = make_cases()


def test_slab_ranges_cover_and_balance():
    for nx, w in ((2048, 8), (100, 3), (7, 7), (17, 4)):
        r = slab_ranges(nx, w)
        assert r[0][0] == 0 and r[-1][1] == nx and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1
        assert [owner_of(i, r) for i in (0, nx - 1)] == [0, w - 1]
    with pytest.raises(ValueError):
        slab_ranges(3, 4)


def _group_from_case(case, n_slabs, opts, halo="copy"):
    import strata_fdtd_b200 as sb
    kw = dict(c=case.get("c", 343.0), rho=case.get("rho", 1.2), courant=case.get("courant", 0.95))
    nu = case.get("nonuniform")
    if nu is None:
        g = LocalSlabGroup(n_slabs, shape=tuple(case["shape"]), resolution=case["resolution"], chunk_steps=16,
                           halo=halo, **kw)
    else:
        g = LocalSlabGroup(n_slabs, grid=sb.NonuniformGrid(nu["x_coords"], nu["y_coords"], nu["z_coords"]),
                           chunk_steps=16, halo=halo, **kw)
    for s in g.slabs:
        if case.get("geometry") is not None:
            g_ = case["geometry"]
            s.set_geometry(g_ if callable(g_) else np.asarray(g_, dtype=bool))
        for b in case.get("pml", []):
            axes = tuple(b.get("axes", ("x", "y", "z")))
            s.add_boundary(sb.PML(depth=b.get("depth", 10), axis="all" if axes == ("x", "y", "z") else axes,
                                  max_sigma=b.get("max_sigma"), order=b.get("order", 3)))
        for src in case.get("sources", []):
            kind = src.get("kind", "point")
            pos = src["position"] if kind == "point" else {"axis": src["axis"], "index": src["index"]}
            s.add_source(sb.GaussianPulse(position=pos, frequency=src["frequency"], bandwidth=src.get("bandwidth"),
                                          amplitude=src.get("amplitude", 1.0), source_type=kind))
        for name, pos in case.get("probes", []):
            s.add_probe(name, position=pos)
        for k, v in opts.items():
            s.set_kernel_option(k, v)
    return g


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["march", "naive"])
@pytest.mark.parametrize("n_slabs", [2, 3])
@pytest.mark.parametrize("name", ["block_pml", "partial_pml_plane", "nonuniform_block_pml", "odd_rigid_box"])
def test_slabs_on_one_device_equal_single_domain(name, n_slabs, kernel):
    case = dict(CASES[name]); case.pop("mics", None)
    opts = {_lib.OPT_KERNEL: _lib.KERNEL_MARCH if kernel == "march" else _lib.KERNEL_NAIVE, _lib.OPT_CHUNK_I: 4}
    steps = 90
    one = build_b200_solver(case)
    for k, v in opts.items():
        one.set_kernel_option(k, v)
    grp = _group_from_case(case, n_slabs, opts)
    one.run(steps=steps)
    grp.run(steps)
    for f in ("p", "vx", "vy", "vz"):
        a, b = grp.get_field(f), one.get_field(f)
        assert np.array_equal(a, b), f"{name}: {f} differs between {n_slabs} slabs and one domain " \
                                     f"(first at {np.argwhere(a != b)[:1]})"
    traces = grp.get_probe_data()
    for pname in one._probes:
        assert np.array_equal(traces[pname], one.get_probe_data(pname)[pname]), pname
    assert np.abs(one.get_field("p")).max() > 0
    grp.close(); one.close()


@pytest.mark.gpu
def test_host_pokes_cross_the_cut():
    """Initial conditions written on the host next to a cut reach the neighbour's ghosts (p and vx)."""
    case = dict(shape=(24, 16, 20), resolution=1e-3, steps=0, pml=[dict(depth=3)])
    one = build_b200_solver(case)
    grp = _group_from_case(case, 2, {})
    one.p[11, 8, 10] = 1.0; one.vx[11, 4, 4] = 0.5; one.p[12, 3, 3] = -2.0
    grp.slabs[0].p[11, 8, 10] = 1.0; grp.slabs[0].vx[11, 4, 4] = 0.5; grp.slabs[1].p[0, 3, 3] = -2.0
    one.run(steps=30); grp.run(30)
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(grp.get_field(f), one.get_field(f)), f


_TWO_RANK = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from cases import make_cases
from util import build_b200_solver
import strata_fdtd_b200 as sb
from strata_fdtd_b200.multi import DistributedFDTDSolver
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
case = make_cases()["block_pml"]
d = DistributedFDTDSolver(shape=case["shape"], resolution=case["resolution"], chunk_steps=16, halo={halo!r})
print("HALO", d.halo)
d.set_geometry(case["geometry"])
d.add_boundary(sb.PML(depth=8))
for s in case["sources"]:
    d.add_source(sb.GaussianPulse(position=s["position"], frequency=s["frequency"]))
for n, p in case["probes"]:
    d.add_probe(n, p)
d.run(steps=100)
fields = {{f: d.gather_field(f) for f in ("p", "vx", "vy", "vz")}}
traces = d.get_probe_data()
e = d.compute_energy()
if rank == 0:
    one = build_b200_solver(case, device=0)
    one.run(steps=100)
    for f in fields:
        assert np.array_equal(fields[f], one.get_field(f)), f
    for n in traces:
        assert np.array_equal(traces[n], one.get_probe_data(n)[n]), n
    assert abs(e - one.compute_energy()) <= 1e-9 * abs(e)
    print("TWO_RANK_OK")
dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.gpu
@pytest.mark.parametrize("halo", ["nccl", "p2p"])
def test_two_ranks_nccl_equal_single_gpu(tmp_path, halo):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "two_rank.py"
    script.write_text(_TWO_RANK.format(root=str(ROOT), halo=halo))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert "TWO_RANK_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
    assert f"HALO {halo}" in res.stdout


@pytest.mark.gpu
def test_c4_enclosure_nonuniform_scaled_down():
    """BASELINE config 4 at 1/8 scale (128x64x64): nonuniform grid + ported-enclosure rigid masks + PML,
    single GPU vs oracle, and 2 / 4 slabs (geometry voxelised per slab through the callable) vs single GPU."""
    from cases import c4_case
    from oracle import oracle as O
    from util import assert_same_as_oracle
    shape, steps = (128, 64, 64), 200
    case = c4_case(shape, steps=steps, stretch_x=1.01)
    assert 0.02 < 1.0 - case["geometry"].mean() < 0.2            # the shell is there
    one = build_b200_solver(case)
    o = O.OracleSolver(case)
    one.run(steps=steps); o.run_steps(steps)
    assert_same_as_oracle(one, o, "c4 scaled")
    assert np.abs(o.probe_array("port_mouth")).max() > 0
    lazy = c4_case(shape, steps=steps, stretch_x=1.01, materialise=False)
    for n_slabs in (2, 4):
        grp = _group_from_case(lazy, n_slabs, {})
        grp.run(steps)
        for f in ("p", "vx", "vy", "vz"):
            assert np.array_equal(grp.get_field(f), one.get_field(f)), (n_slabs, f)
        tr = grp.get_probe_data()
        for pname in one._probes:
            assert np.array_equal(tr[pname], one.get_probe_data(pname)[pname]), pname
        grp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_slabs", [2, 3])
@pytest.mark.parametrize("name", ["block_pml", "partial_pml_plane", "nonuniform_block_pml"])
def test_peer_store_halo_protocol_on_one_device(name, n_slabs):
    """The fused halo path (K1 peer stores + neighbour flags, sb_set_peers) with the peers on one device:
    whole chunks are enqueued per slab with no host work between steps, results equal the single domain."""
    case = dict(CASES[name]); case.pop("mics", None)
    steps = 70
    one = build_b200_solver(case)
    one.set_kernel_option(_lib.OPT_CHUNK_I, 4)
    grp = _group_from_case(case, n_slabs, {_lib.OPT_CHUNK_I: 4}, halo="p2p")
    one.run(steps=steps)
    grp.run(steps)
    for f in ("p", "vx", "vy", "vz"):
        assert np.array_equal(grp.get_field(f), one.get_field(f)), (name, f)
    tr = grp.get_probe_data()
    for pname in one._probes:
        assert np.array_equal(tr[pname], one.get_probe_data(pname)[pname]), pname
    grp.close(); one.close()
