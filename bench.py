#!/usr/bin/env python
"""Benchmark of the strata-fdtd time-stepping hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own C++/OpenMP kernels on host cores

A "step" is one FDTD time step of the full path (velocity + rigid faces + pressure + sponge + source
injection + probe recording) over the whole grid.  Metric: cell-updates per second, whole job.

Default workload = the weak-scaling series of BASELINE.json configs[4] (2048^3 + PML over 8 GPUs):
every GPU owns a 256 x 2048 x 2048 slab (1.07 G cells, 34 GB of fields), so N GPUs simulate
256N x 2048 x 2048.  Fields are far larger than L2 (126 MB), so no flush is needed between steps.
Other workloads (--workload c3_512 | c2_200 | c1_100 | c3_512_ade) are the remaining BASELINE configs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "cell_updates_per_s"
UNIT = "Gcell-updates/s"
ALGO_BYTES_PER_CELL = 32.0          # read + write of p, vx, vy, vz in fp32


def measured_peak_gbs():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- workloads
def workload_case(name: str, n_gpus: int) -> tuple[dict, str]:
    from strata_fdtd_b200.workloads import c1_case, c2_case, c3_case, c4_case, c5_case
    if name == "c5_weak":
        case = c5_case(n_gpus)
        nx = case["shape"][0]
        return case, f"c5_weak: {nx}x2048x2048 uniform 1 mm, PML(10), 1 point source, 8 probes ({n_gpus} x 256x2048x2048 slabs)"
    if name == "c3_512":
        c = c2_case(512, steps=0)
        c["probes"] = [(f"p{a}{b}", (384, 32 + 64 * a, 32 + 64 * b)) for a in range(8) for b in range(8)]
        return c, "c3_512: 512^3 uniform 1 mm, PML(10), 1 point source, 64 probes (no material)"
    if name == "c3_512_ade":
        return c3_case(512, steps=0), "c3_512_ade: 512^3, PML(10), ADE sphere r=51 (2 Debye + 1 Lorentz), 64 probes"
    if name == "c3_512_ade_slab":
        return (c3_case(512, steps=0, slab=True),
                "c3_512_ade_slab: 512^3, PML(10), ADE material in the upper quarter (33.5 M cells, 2 Debye + 1 Lorentz), 64 probes")
    if name == "c2_200":
        return c2_case(200, steps=0), "c2_200: 200^3 uniform 1 mm, PML(10), 1 point source, 1 probe"
    if name == "c1_100":
        return c1_case(0), "c1_100: 100^3 uniform 1 mm, PML(10), 1 kHz pulse, 1 probe"
    if name in ("c4_enclosure", "c4_enclosure_closed_form"):
        enclosure = "reference" if name == "c4_enclosure" else "closed_form"
        return (c4_case((1024, 512, 512), steps=0, materialise=False, enclosure=enclosure),
                "c4_enclosure: 1024x512x512 nonuniform (axis 0 stretched 1.002 from the centre), ported-enclosure "
                f"rigid masks ({enclosure} CSG), PML(10), 1 source, 8 probes ({n_gpus} slab(s), strong scaling)")
    raise SystemExit(f"unknown workload {name}")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.proc.wait()
        self.tmp.flush()
        rows = [r.split(",") for r in Path(self.tmp.name).read_text().strip().splitlines() if r.count(",") >= 7]
        os.unlink(self.tmp.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                if r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU baseline (reference kernels)
def cpu_reference_sample(steps: int = 60, warmup: int = 2, shape=(512, 512, 512), case: dict | None = None):
    """The reference's compiled C++/OpenMP kernels (oracle/_ref) on the host cores, same path (PML + source +
    probes): on ``case`` itself when given, else on a bounded sub-grid of the workload.  Falls back to the oracle
    port if _ref is not built."""
    from oracle import oracle as O
    from oracle import ref_loader as R
    if case is None:
        n = shape[0]
        case = dict(shape=shape, resolution=1e-3, steps=0, pml=[dict(depth=10)],
                    sources=[dict(kind="point", position=(n // 4, shape[1] // 2, shape[2] // 2), frequency=1000.0)],
                    probes=[("probe", (3 * n // 4, shape[1] // 2, shape[2] // 2))])
        what = "sub-grid of the workload (PML 10, point source, probe)"
    else:
        shape = tuple(case["shape"])
        what = "the workload itself (whole grid, all sources and probes)"
    cores = os.cpu_count() or 1
    if R.have_ref_kernels():
        k = R.load_ref_kernels()
        k.set_num_threads(cores)
        drv, kind = R.RefKernelSolver(case), "reference"
        cores = k.get_num_threads()
    else:
        O.set_threads(cores)
        drv, kind = O.OracleSolver(case), "port"
    for _ in range(warmup):
        drv.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        drv.step()
    dt = time.perf_counter() - t0
    cells = int(np.prod(shape, dtype=np.int64))
    return {"value": cells * steps / dt / 1e9, "unit": UNIT, "cores": int(cores), "kind": kind,
            "sample": f"{shape[0]}x{shape[1]}x{shape[2]} {what}, {steps} steps after {warmup} warm-up, {dt:.2f} s",
            "ms_per_step": dt / steps * 1e3}


def reference_can_run(case: dict) -> tuple[bool, str]:
    """The reference indexes cells with 32-bit ints (fdtd_types.hpp:26-42) and keeps everything in host RAM."""
    cells = int(np.prod(case["shape"], dtype=np.int64))
    if cells >= 2 ** 31:
        return False, f"{cells} cells exceed the reference's int32 cell index (fdtd_types.hpp:26-42)"
    if case.get("nonuniform") or case.get("materials") or case.get("geometry") is not None:
        return False, "the kernel-level reference driver covers the uniform, material-free path"
    need = cells * (4 * 4 + 1) * 1.15
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except ImportError:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    if avail < need:
        return False, f"needs {need / 2**30:.0f} GiB of host memory, {avail / 2**30:.0f} GiB available"
    return True, ""


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    case, label = workload_case(a.workload, a.gpus)
    ok, why_not = reference_can_run(case)
    steps, warmup = max(1, a.steps), max(1, min(a.warmup, 3))
    if ok:
        r = cpu_reference_sample(steps=steps, warmup=warmup, case=case)
        note = "reference C++/OpenMP kernels on the host cores, on the workload as worded"
    else:
        r = cpu_reference_sample(steps=steps, warmup=warmup)
        note = ("reference C++/OpenMP kernels timed on a bounded 512^3 sub-grid on the host cores (rate is per cell, so "
                f"size-comparable); not the whole workload because {why_not}")
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if a.workload == "c5_weak" else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": label, "note": note, "whole_workload": ok},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- N > 1: is the decomposed run right?
def decomposition_selfcheck(dist, world: int, rank: int, local_rank: int, halo: str) -> dict:
    """Decomposition invariance inside the bench's own process group, with the halo mode of the timed run: a small grid
    cut into `world` slabs (solid block through the first cut, sponge, a source right next to a cut, probes on both
    sides) must reproduce the single-GPU run of rank 0 bit for bit -- the four final fields and every trace.  The
    reference has no multi-GPU path; this is the contract that replaces it."""
    from strata_fdtd_b200.workloads import build_distributed_solver, build_solver
    nx, steps = 16 * world, 64
    g = np.ones((nx, 96, 200), dtype=bool)
    g[12:20, 30:60, 80:140] = False
    case = dict(shape=(nx, 96, 200), resolution=1e-3, geometry=g, pml=[dict(depth=6)],
                sources=[dict(kind="point", position=(15, 20, 100), frequency=20e3),          # last plane below the first cut
                         dict(kind="point", position=(nx - 16, 20, 40), frequency=15e3, amplitude=0.5)],
                probes=[("below_cut", (15, 26, 104)), ("above_cut", (16, 26, 104)), ("behind_block", (18, 66, 100)),
                        ("low", (1, 70, 30)), ("last_cut", (nx - 17, 24, 44))])
    d = build_distributed_solver(case, device=local_rank, chunk_steps=32, halo=halo)
    d.run(steps=steps)
    fields = {f: d.gather_field(f) for f in ("p", "vx", "vy", "vz")}
    traces = d.get_probe_data()
    verdict = [None]
    if rank == 0:
        one = build_solver(case, device=local_rank, distributed=False)
        one.run(steps=steps)
        bad = [f for f in fields if not np.array_equal(fields[f], one.get_field(f))]
        bad += [n for n in traces if not np.array_equal(traces[n], one.get_probe_data(n)[n])]
        # not vacuous: the wave has crossed the first cut (the far probes may still be silent after 64 steps)
        alive = float(np.abs(fields["p"]).max()) > 0 and all(np.abs(traces[n]).max() > 0 for n in ("below_cut", "above_cut"))
        verdict[0] = {"invariance": "bit-exact" if not bad and alive else "MISMATCH: " + ",".join(bad or ["dead fields"]),
                      "case": f"{nx}x96x200, {world} slabs, solid block through a cut, PML(6), 2 sources, 5 probes, "
                              f"{steps} steps vs the single-GPU run", "halo": d.halo}
        one.close()
    dist.broadcast_object_list(verdict, src=0)
    d.close()
    return verdict[0]


def halo_planes_equal(dist, drv) -> bool:
    """After the timed region: every ghost plane equals the plane its owner holds, bit for bit -- the neighbour's p plane
    on both sides of each cut and the redundantly computed ghost face vx[-1] against the lower neighbour's vx[nx-1]."""
    import torch
    s = drv.slab
    dev = s._dev
    rank, world = drv.rank, drv.world
    torch.cuda.synchronize()
    dist.barrier()
    ok, live = True, False
    with torch.cuda.stream(dev.stream):
        for field in ("p", "vx"):
            h = s.halo_planes(field)
            ops, checks = [], []
            if rank < world - 1:
                ops.append(dist.P2POp(dist.isend, h["send_hi"].contiguous(), rank + 1))
            if rank > 0:
                buf = torch.empty_like(h["recv_lo"])
                ops.append(dist.P2POp(dist.irecv, buf, rank - 1))
                checks.append((buf, h["recv_lo"]))
            if field == "p":
                if rank > 0:
                    ops.append(dist.P2POp(dist.isend, h["send_lo"].contiguous(), rank - 1))
                if rank < world - 1:
                    buf = torch.empty_like(h["recv_hi"])
                    ops.append(dist.P2POp(dist.irecv, buf, rank + 1))
                    checks.append((buf, h["recv_hi"]))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            torch.cuda.current_stream().synchronize()
            for got, ghost in checks:
                ok = ok and bool(torch.equal(got, ghost))
                live = live or bool(got.abs().max() > 0)
    # equal everywhere, and not vacuously: at least one compared plane of the job carries signal
    t = torch.tensor([1 if ok else 0, 0 if live else 1], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t[0].item()) and not bool(t[1].item())


# ----------------------------------------------------------------------------- our arm
def run_b200_arm(a):
    import torch
    from strata_fdtd_b200 import _lib
    from strata_fdtd_b200.workloads import build_distributed_solver, build_solver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        from strata_fdtd_b200.solver import nccl_options
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=nccl_options())
    parity_n = decomposition_selfcheck(dist, world, rank, local_rank, a.halo) if world > 1 else None
    case, label = workload_case(a.workload, world)
    cells_total = int(np.prod(case["shape"], dtype=np.int64))
    K, W = a.steps, max(4, a.warmup)        # >= 4 so that the warm-up runs the same (chunk) kernels as the timed region

    if world == 1:
        s = build_solver(case, device=local_rank, distributed=False)      # the solver's own chunking, as a user gets it
        slab, drv = s, None
    else:
        drv = build_distributed_solver(case, device=local_rank, chunk_steps=max(K, W), halo=a.halo)
        slab = drv.slab
    for opt, val in ((_lib.OPT_ROWS_PER_THREAD, a.rows), (_lib.OPT_WARPS_J, a.warps_j), (_lib.OPT_WARPS_K, a.warps_k),
                     (_lib.OPT_CHUNK_I, a.chunk_i)):
        if val is not None:
            slab.set_kernel_option(opt, val)
    if a.graph:
        slab.set_kernel_option(_lib.OPT_USE_GRAPH, 1)
    if a.ade_layout is not None:
        slab.set_kernel_option(_lib.OPT_ADE_LAYOUT, a.ade_layout)
    if a.ade_chunk is not None:
        slab.set_kernel_option(_lib.OPT_ADE_CHUNK_I, a.ade_chunk)
    if a.ade_warps is not None:
        slab.set_kernel_option(_lib.OPT_ADE_WARPS, a.ade_warps)
    if a.ade_occ is not None:
        slab.set_kernel_option(_lib.OPT_ADE_OCCUPANCY, a.ade_occ)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    run = (lambda n: s.run(steps=n)) if world == 1 else (lambda n: drv.run(steps=n))

    # ---- e2e: the public API call, host buffers in, host traces out, every chunk --------------------
    run(W)                                   # warm-up (also builds device state, uploads tables)
    run(K)                                   # untimed pass at the timed chunk length (instantiates its CUDA graph)
    barrier()
    t0 = time.perf_counter()
    run(K)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    n_src, n_rec = len(case["sources"]), len(case["probes"])

    # ---- device-resident: inputs already in HBM, CUDA events on the launching stream ----------------
    dev = slab._dev
    lib, h = dev.lib, dev.handle
    st0 = slab.device_stats()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if world == 1:
        src = torch.zeros(K * max(1, n_src), dtype=torch.float64, device=dev.device)
        rec = torch.zeros(K * max(1, n_rec), dtype=torch.float32, device=dev.device)
        times = np.cumsum(np.full(K, float(s.dt))) + s.time
        src.copy_(torch.from_numpy(np.ascontiguousarray(s._waveform_table(times)[:, :max(1, n_src)])).reshape(-1))
        torch.cuda.synchronize()
        # The e2e passes above kept the board at its power limit for a while (a step kernel at ~100 % of the HBM bandwidth
        # draws ~1 kW on real data); let the limiter's average recover so that this region starts like a fresh job, then
        # warm up.  The untimed pass of K steps is only needed where the library captures a K-step CUDA graph or runs a
        # chunk kernel (<= 32 M cells); on large grids W steps are the warm-up.
        time.sleep(2.0)
        _lib.check(lib.sb_step_n_async(h, W, src.data_ptr(), rec.data_ptr()))
        if cells_total <= (32 << 20):
            _lib.check(lib.sb_step_n_async(h, K, src.data_ptr(), rec.data_ptr()))   # untimed: instantiates the K-step graph
        barrier()
        st0 = slab.device_stats()                # launches are counted over the timed region only
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(dev.stream)
        _lib.check(lib.sb_step_n_async(h, K, src.data_ptr(), rec.data_ptr()))
        e1.record(dev.stream)
        barrier()
        dev_ms = e0.elapsed_time(e1)
    else:
        if drv.halo != "p2p":
            drv._overlap_possible()
        slab.begin_chunk(K)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(dev.stream)
        if drv.halo == "p2p":
            slab.enqueue_steps(K)                # halos travel inside the step kernels (NVLink peer stores)
        else:
            for _ in range(K):
                drv._step_with_exchange()    # cut planes first, their send/recv on a second stream beside the interior
        e1.record(dev.stream)
        barrier()
        dev_ms = e0.elapsed_time(e1)
        slab.end_chunk()
        t = torch.tensor([dev_ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dev_ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    if parity_n is not None:
        parity_n["halo_planes_equal"] = halo_planes_equal(dist, drv)
    st1 = slab.device_stats()
    launches = int(st1["kernels_launched"] - st0["kernels_launched"]) - (0 if world > 1 else 0)

    # ---- roofline of the dominant kernel: per-launch CUDA events around the fused step kernel ------
    slab.set_kernel_option(_lib.OPT_PROFILE, 1)
    # the chunk kernels (resident K5 / pipelined K6) cover all steps of a call in one launch: time a launch of the full
    # length, so that the load / store of the chunk ends weighs as it does in the timed region
    kp = K if slab.device_stats()["kernel_variant"] in (4, 5) else min(K, 20)
    if world == 1:
        _lib.check(lib.sb_step_n_async(h, kp, src.data_ptr(), rec.data_ptr()))
    else:
        slab.begin_chunk(kp)
        if drv.halo == "p2p":
            slab.enqueue_steps(kp)
        else:
            for _ in range(kp):
                drv._step_with_exchange()
        slab.end_chunk()
    import ctypes as C
    mean_ms, min_ms, n_l = C.c_double(), C.c_double(), C.c_int()
    _lib.check(lib.sb_profile_read(h, C.byref(mean_ms), C.byref(min_ms), C.byref(n_l)))
    slab.set_kernel_option(_lib.OPT_PROFILE, 0)
    barrier()
    shape4, tuned_ms = (C.c_int32 * 4)(), C.c_float()
    _lib.check(lib.sb_tuned(h, shape4, C.byref(tuned_ms)))
    peak, peak_src = measured_peak_gbs()
    cells_rank = int(np.prod(slab.shape, dtype=np.int64))
    # algorithmic bytes of one launch of the step kernel: 32 B per cell; with dispersive materials the step kernel is
    # two concurrent launches (K1 and K1-ADE, or K1 beside the ADE list kernels) bracketed together, and the bytes are
    # sb_query's model: 32 + the mask byte and 8 (Debye) / 16 (Lorentz) B per pole on the cells that carry the material
    has_ade = bool(case.get("materials"))
    algo_bytes_per_cell = float(slab.device_stats()["algorithmic_bytes_per_cell"]) if has_ade else ALGO_BYTES_PER_CELL
    achieved = algo_bytes_per_cell * cells_rank / (mean_ms.value * 1e-3) / 1e9

    if rank == 0:
        value = cells_total * K / (dev_ms * 1e-3) / 1e9
        cpu = cpu_reference_sample() if (world == 1 and not a.no_cpu_baseline) else None
        traffic_file = ROOT / "profiles" / "k1_dram_traffic.json"
        traffic = json.loads(traffic_file.read_text()).get(a.workload) if traffic_file.exists() else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_ms / K, "higher_is_better": True,
                "scaling": "weak" if a.workload == "c5_weak" else "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": label, "cells_per_gpu": cells_rank, "l2_policy": "fields (34 GB/GPU) >> L2, no flush needed"
                           if a.workload == "c5_weak" else "inputs larger than L2 for >=200^3; small grids are L2-resident by nature",
                           "kernel": {0: "auto", 1: "naive", 2: "march", 3: "tma", 4: "resident", 5: "pipeline"}[st1["kernel_variant"]],
                           "launch_shape": ({"source": "chunk kernel: shape fixed by the library (DESIGN.md 4a / 4b)"}
                                            if st1["kernel_variant"] in (4, 5) else
                                            {"rows_per_thread": shape4[0], "warps_j": shape4[1], "warps_k": shape4[2],
                                             "chunk_planes": shape4[3], "source": "library autotune"} if a.rows is None else
                                            {"rows_per_thread": a.rows, "warps_j": a.warps_j, "warps_k": a.warps_k,
                                             "chunk_planes": a.chunk_i, "source": "flag"}),
                           "parallelism": f"slab{world}" if world > 1 else "single",
                           "halo": (drv.halo if drv is not None else None)},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "kernel": "k5_resident (per step of one chunk launch; fields stay in shared memory, so "
                                       "'achieved' is an HBM-equivalent rate)" if st1["kernel_variant"] == 4 else
                                       "k6_pipeline (k1_tile; per step of one chunk launch)" if st1["kernel_variant"] == 5 else
                                       "k1_step_march + ADE kernels of the same step (concurrent launches, timed together)" if has_ade else
                                       "k1_step_march",
                             "kernel_ms_mean": mean_ms.value, "kernel_ms_min": min_ms.value, "launches_timed": n_l.value,
                             "algorithmic_bytes_per_launch": algo_bytes_per_cell * cells_rank,
                             # whole-step model of sb_query: 32 (+1 with a face mask) + sum over poles of 8 (Debye) or 16
                             # (Lorentz) bytes x the fraction of cells that carry the material
                             "step_bytes_per_cell_model": st1["algorithmic_bytes_per_cell"]},
                "e2e": {"value": cells_total * K / e2e_s / 1e9, "unit": UNIT,
                        "h2d_bytes_per_step": 8 * n_src, "d2h_bytes_per_step": 4 * n_rec,
                        "note": "FDTDSolver.run(): waveform table host->device and probe traces device->host every chunk; "
                                "fields stay resident between steps as in the reference"},
                "gpu_launches": launches, "clocks": clocks}
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if parity_n is not None:
            line["parity_n"] = parity_n
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if parity_n is not None and (parity_n["invariance"] != "bit-exact" or not parity_n["halo_planes_equal"]):
        raise SystemExit(3)                  # a fast wrong answer is not a result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5_weak")
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--warps-j", type=int, default=None)
    ap.add_argument("--warps-k", type=int, default=None)
    ap.add_argument("--chunk-i", type=int, default=None)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--ade-layout", type=int, default=None, help="0 auto, 1 compact list, 2 dense box, 3 fused (ADE workloads)")
    ap.add_argument("--ade-chunk", type=int, default=None, help="K1-ADE planes per tile")
    ap.add_argument("--ade-warps", type=int, default=None, help="K1-ADE warps per block")
    ap.add_argument("--ade-occ", type=int, default=None, help="K1-ADE blocks per SM aimed at (2 or 3)")
    ap.add_argument("--halo", default="auto", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        a.steps = 60 if a.steps is None else a.steps
        a.warmup = 2 if a.warmup is None else a.warmup
        run_reference_arm(a)
    else:
        a.steps = 50 if a.steps is None else a.steps
        a.warmup = 5 if a.warmup is None else a.warmup
        run_b200_arm(a)


if __name__ == "__main__":
    main()
