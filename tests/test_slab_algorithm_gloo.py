"""World-size-2 gloo test (CPU) of the slab algorithm itself: ghost planes, the redundant vx[-1] update,
'exchange p once per step after injection'.  Each rank steps the ORACLE's C kernels on its slab extended
by the live ghost planes and trades p planes with torch.distributed send/recv exactly as
strata_fdtd_b200.multi does on GPUs; the gathered result must equal the single-domain oracle bit-for-bit."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cases import make_cases
    from oracle import oracle as O
    from strata_fdtd_b200.multi import slab_ranges
    case = make_cases()["block_pml"]
    steps = 80
    full = O.OracleSolver(case)                      # host numbers (dt, tables) come from the global problem
    nx, ny, nz = full.shape
    i0, i1 = slab_ranges(nx, world)[rank]
    lo, hi = int(rank > 0), int(rank < world - 1)
    sl = slice(i0 - lo, i1 + hi)
    # extended local state: owned planes + live ghosts
    f = {k: np.zeros((i1 - i0 + lo + hi, ny, nz), np.float32) for k in ("p", "vx", "vy", "vz")}
    geom = np.ascontiguousarray(full.geometry[sl]).view(np.uint8)
    dec = [(np.ascontiguousarray(sp["decay"][0][sl]), sp["decay"][1], sp["decay"][2]) for sp in full.sponges]
    L, F, I = O.lib(), O.F, O.I
    fp, up = O._fp, O._up
    dims = (I(f["p"].shape[0]), I(ny), I(nz))
    src = full.sources[0]; si, sj, sk = src["position"]
    probes = [(n, p) for n, p in full.probes if i0 <= p[0] < i1]
    traces = {n: [] for n, _ in probes}
    t = 0.0
    for _ in range(steps):
        L.orc_update_velocity(fp(f["p"]), fp(f["vx"]), fp(f["vy"]), fp(f["vz"]), *dims, F(full.cv), None, None, None)
        L.orc_apply_rigid(fp(f["vx"]), fp(f["vy"]), fp(f["vz"]), up(geom), *dims)
        L.orc_update_pressure(fp(f["p"]), fp(f["vx"]), fp(f["vy"]), fp(f["vz"]), up(geom), *dims, F(full.cp), None, None, None)
        for d in dec:
            L.orc_sponge_velocity(fp(f["vx"]), fp(f["vy"]), fp(f["vz"]), *dims, fp(d[0]), fp(d[1]), fp(d[2]))
        for d in dec:
            L.orc_sponge_pressure(fp(f["p"]), *dims, fp(d[0]), fp(d[1]), fp(d[2]))
        w = O.gaussian_pulse(t, src["frequency"], src.get("bandwidth"), src.get("amplitude", 1.0))
        if i0 <= si < i1 and full.geometry[si, sj, sk]:
            q = si - i0 + lo
            f["p"][q, sj, sk] = np.float32(np.float64(f["p"][q, sj, sk]) + w)
        # exchange p once per step, after injection (multi.DistributedFDTDSolver._exchange)
        ops, bufs = [], []
        if lo:
            send = torch.from_numpy(f["p"][lo].copy()); recv = torch.empty_like(send); bufs.append((recv, 0))
            ops += [dist.P2POp(dist.isend, send, rank - 1), dist.P2POp(dist.irecv, recv, rank - 1)]
        if hi:
            send = torch.from_numpy(f["p"][-1 - hi].copy()); recv = torch.empty_like(send); bufs.append((recv, -1))
            ops += [dist.P2POp(dist.isend, send, rank + 1), dist.P2POp(dist.irecv, recv, rank + 1)]
        for wk in dist.batch_isend_irecv(ops):
            wk.wait()
        for recv, where in bufs:
            f["p"][where] = recv.numpy()
        for n, (pi, pj, pk) in probes:
            traces[n].append(float(f["p"][pi - i0 + lo, pj, pk]))
        t = t + full.dt
    own = slice(lo, lo + i1 - i0)
    np.savez(Path(out_dir) / f"rank{rank}.npz", **{k: v[own] for k, v in f.items()},
             **{"probe_" + n: np.array(v, np.float32) for n, v in traces.items()})
    dist.barrier(); dist.destroy_process_group()


def test_two_slab_oracle_over_gloo_equals_single_domain(tmp_path):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    from cases import make_cases
    from oracle import oracle as O
    world, port = 2, 29533 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    case = make_cases()["block_pml"]
    full = O.OracleSolver(case)
    full.run_steps(80)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for k in ("p", "vx", "vy", "vz"):
        got = np.concatenate([p[k] for p in parts], axis=0)
        assert np.array_equal(got, getattr(full, k)), k
    for n, _ in full.probes:
        got = next(p["probe_" + n] for p in parts if "probe_" + n in p)
        assert np.array_equal(got, full.probe_array(n)), n
    assert np.abs(full.p).max() > 0
