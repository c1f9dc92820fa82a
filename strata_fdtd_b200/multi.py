"""Multi-GPU execution: slabs of the grid along reference axis 0, one plane of halo per face.

The reference has no distributed path at all; this is the design of
SURVEY.md 8(e).  Rank r owns planes [i0, i1) of every field plus a ghost plane on each interior
face.  Per time step only **p** crosses a face, once, after the step:

  * p_new[i1-1]  -> upper neighbour's ghost plane -1      (it needs it for its redundant vx[-1] update)
  * p_new[i0]    -> lower neighbour's ghost plane nx      (it needs it to update its last vx plane)

The normal velocity on the cut is *not* exchanged: each slab keeps a ghost copy of vx[-1] and
updates it redundantly from the two p planes it already has -- the same fp32 operations in the
same order as the owner performs, so an N-slab run is bit-identical to the single-GPU run
(tests/test_multi_gpu.py checks exactly that).  Sources are injected by the owning slab before
the exchange, probes are recorded by the owning slab; there is no collective in the step.

Two drivers share the stepping code:

  * ``DistributedFDTDSolver`` -- one process per GPU, ``torch.distributed`` (NCCL send/recv over
    NVLink batched into one group per step; gloo works for CPU plumbing tests).
  * ``LocalSlabGroup`` -- N slabs inside one process on one device (peer copies degrade to
    device-to-device copies); used to test the halo logic on a single GPU.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .solver import _FIELDS, FDTDSolver
from .sources import combine_corner_samples


def slab_ranges(nx: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced split of [0, nx) into ``world`` slabs (the first nx % world get one more plane)."""
    if world < 1 or world > nx:
        raise ValueError(f"cannot cut {nx} planes into {world} slabs")
    base, extra = divmod(nx, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _drain_corner_samples(slab) -> tuple[dict, np.ndarray]:
    """(corner traces recorded since the last drain, their time axis) of one slab."""
    data = {k: np.concatenate(v) for k, v in slab._corner_data.items()}
    times = np.concatenate(slab._corner_times) if slab._corner_times else np.zeros(0)
    slab._corner_data.clear()
    slab._corner_times.clear()
    return data, times


def owner_of(i: int, ranges: list[tuple[int, int]]) -> int:
    for r, (lo, hi) in enumerate(ranges):
        if lo <= i < hi:
            return r
    raise ValueError(f"plane {i} outside the grid")


class SlabSolver(FDTDSolver):
    """FDTDSolver restricted to one slab, stepped one step at a time so halos can be exchanged in between."""
    _is_slab = True

    def begin_chunk(self, m: int) -> None:
        dev = self._sync_to_device()
        torch = dev.torch
        self._chunk_m = m
        self._chunk_q = 0
        times = np.empty(m, dtype=np.float64)
        t = self._time
        for q in range(m):
            times[q] = t
            t = t + self.dt
        self._chunk_times, self._chunk_t_end = times, t
        self._n_src = max(1, len(self._sources))
        self._n_rec = len(self._local_probes) + len(self._corner_keys)
        W = self._waveform_table(times) if self._sources else np.zeros((m, 1))
        with torch.cuda.stream(dev.stream):
            self._W_dev = torch.from_numpy(np.ascontiguousarray(W)).to(dev.device, non_blocking=False)
            self._rec_dev = torch.zeros((m, max(1, self._n_rec)), dtype=torch.float32, device=dev.device)

    def enqueue_step(self) -> None:
        dev, q = self._dev, self._chunk_q
        _lib.check(dev.lib.sb_step_n_async(dev.handle, 1, self._W_dev.data_ptr() + q * self._n_src * 8,
                                           self._rec_dev.data_ptr() + q * max(1, self._n_rec) * 4))
        self._chunk_q += 1

    def enqueue_steps(self, m: int) -> None:
        """All remaining steps of the chunk in one call (peer-to-peer halo mode: no host work between steps)."""
        dev, q = self._dev, self._chunk_q
        _lib.check(dev.lib.sb_step_n_async(dev.handle, m, self._W_dev.data_ptr() + q * self._n_src * 8,
                                           self._rec_dev.data_ptr() + q * max(1, self._n_rec) * 4))
        self._chunk_q += m

    def end_chunk(self) -> None:
        dev, m = self._dev, self._chunk_m
        with dev.torch.cuda.stream(dev.stream):
            rec = self._rec_dev.cpu().numpy()
        _lib.check(dev.lib.sb_synchronize(dev.handle))
        for q, pr in enumerate(self._local_probes):
            pr.data.extend(rec[:, q].tolist())
        self._store_corner_samples(rec, len(self._local_probes), self._chunk_times)
        self._host_stale = set(_FIELDS)
        self._step_count += m
        self._time = self._chunk_t_end

    def halo_planes(self, field: str = "p", next_set: bool = False) -> dict:
        """torch views of the planes that take part in the exchange, in the CURRENT set (or the one being written)."""
        dev = self._dev
        t = dev.sets[dev.current_set() ^ int(next_set)][_FIELDS.index(field)]
        nx = self.shape[0]
        return {"send_lo": t[1], "send_hi": t[nx], "recv_lo": t[0], "recv_hi": t[nx + 1]}

    def enqueue_cuts(self) -> bool:
        """First part of the next step: only the planes next to the cuts (sb_step_cuts_async).  True if launched."""
        import ctypes as C
        applied = C.c_int(0)
        _lib.check(self._dev.lib.sb_step_cuts_async(self._dev.handle, C.byref(applied)))
        return bool(applied.value)

    def sources_touch_cut_planes(self) -> bool:
        """Some source writes into the first / last owned plane next to a neighbour (those planes then change after K1)."""
        cells = self._build_source_table()[0]
        if len(cells) == 0:
            return False
        planes = np.floor_divide(cells, self.shape[1] * self.shape[2])
        cut = ([0, -1] if self._has_lower else []) + ([self.shape[0] - 1] if self._has_upper else [])
        return bool(np.isin(planes, cut).any())


def _chunks(n_steps: int, chunk: int):
    done = 0
    while done < n_steps:
        m = min(chunk, n_steps - done)
        yield m
        done += m


class DistributedFDTDSolver:
    """One slab per process; the public surface of FDTDSolver (core/solver.py:1442-3442) with global coordinates.

    Launch with ``python -m torch.distributed.run --nproc-per-node N ...`` (one rank per GPU).  ``FDTDSolver(...)``
    itself returns this class when it is constructed inside such a job (solver.FDTDSolver.__new__), so a script
    written for one GPU runs unchanged on N.  Every rank makes the same calls with global arguments; results
    (`get_probe_data`, fields, energy, microphones) are available on every rank, files are written by rank 0.

    ``halo``: "p2p" = exchange fused into the step kernel over NVLink peer memory (needs the NCCL backend and
    symmetric memory), "nccl" = one grouped send/recv pair per face and step on the step's stream (with the gloo
    backend the planes are staged through host memory: the CPU-plumbing route the tests use on a single GPU),
    "auto" = p2p where available.
    """

    def __init__(self, shape=None, resolution=None, grid=None, c=343.0, rho=1.2, courant=0.95,
                 backend="b200", warn_energy_drift=False, energy_drift_threshold=0.01, device=None, chunk_steps=None,
                 group=None, halo="auto"):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if halo not in ("auto", "p2p", "nccl"):
            raise ValueError("halo must be 'auto', 'p2p' or 'nccl'")
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nx = int((grid.shape if grid is not None else shape)[0])
        self.ranges = slab_ranges(nx, self.world)
        chunk_steps = 64 if chunk_steps is None else int(chunk_steps)
        self.slab = SlabSolver(shape=shape, resolution=resolution, grid=grid, c=c, rho=rho, courant=courant,
                               backend=backend, device=device, chunk_steps=chunk_steps, slab=self.ranges[self.rank],
                               warn_energy_drift=warn_energy_drift, energy_drift_threshold=energy_drift_threshold)
        self.chunk_steps = chunk_steps
        self._ghosts_fresh = False
        self._staged = dist.get_backend(group) != "nccl"      # gloo cannot move device memory: stage planes on the host
        self.halo = "nccl"
        self._symm = []                          # keeps symmetric allocations / handles alive
        self._energy_history: list = []
        self._snapshots: list = []
        self._velocity_snapshots: list = []
        self._snapshot_velocity = False
        self._snapshot_interval = None
        self.last_run_stats: dict = {}
        if halo in ("auto", "p2p") and self.world > 1 and not self._staged:
            try:
                self._setup_p2p()
                self.halo = "p2p"
            except Exception as e:               # symmetric memory unavailable -> host-driven NCCL exchange
                if halo == "p2p":
                    raise
                import warnings
                warnings.warn(f"peer-to-peer halo unavailable ({type(e).__name__}: {e}); using NCCL send/recv")
                self.slab._p_allocator = None
        elif halo == "p2p":
            raise RuntimeError("halo='p2p' needs the NCCL backend and more than one rank")

    def _setup_p2p(self):
        """Fused halo: K1 stores its cut planes of p straight into the neighbours' ghost planes over NVLink.

        The two p buffers and a small flag array live in torch symmetric memory, so every rank can map its
        neighbours' copies; the addresses go to the C ABI (sb_set_peers) and no collective remains in the step."""
        import ctypes as C
        import torch
        import torch.distributed._symmetric_memory as symm
        dist, s = self.dist, self.slab
        group = self.group if self.group is not None else dist.group.WORLD
        max_planes = max(hi - lo for lo, hi in self.ranges) + 2
        made = []

        def alloc(planes, ny, pitch, device):
            t = symm.empty((max_planes, ny, pitch), dtype=torch.float32, device=device)
            t.zero_()
            made.append(t)
            return t

        s._p_allocator = alloc
        dev = s._ensure_device()
        flags = symm.empty(8, dtype=torch.int32, device=dev.device)
        flags.zero_()
        torch.cuda.synchronize(dev.device)
        handles = [symm.rendezvous(t, group) for t in made]
        fh = symm.rendezvous(flags, group)
        self._symm = [made, flags, handles, fh]

        def peer_ptr(h, t, r):
            return h.get_buffer(r, tuple(t.shape), t.dtype, 0).data_ptr()

        lo, hi = self.rank - 1, self.rank + 1
        lo_sets = (C.c_void_p * 2)(*[peer_ptr(h, t, lo) if lo >= 0 else None for h, t in zip(handles, made)])
        hi_sets = (C.c_void_p * 2)(*[peer_ptr(h, t, hi) if hi < self.world else None for h, t in zip(handles, made)])
        lo_nx = (self.ranges[lo][1] - self.ranges[lo][0]) if lo >= 0 else 0
        lo_flag = peer_ptr(fh, flags, lo) + 4 if lo >= 0 else None          # lower neighbour's flags[1]
        hi_flag = peer_ptr(fh, flags, hi) if hi < self.world else None       # upper neighbour's flags[0]
        _lib.check(dev.lib.sb_set_peers(dev.handle, lo_sets, hi_sets, lo_nx, C.c_void_p(flags.data_ptr()),
                                        C.c_void_p(lo_flag) if lo_flag else None,
                                        C.c_void_p(hi_flag) if hi_flag else None))
        dist.barrier(group)

    # ---- delegated set-up (global coordinates; every rank makes the same calls) -----------------
    shape = property(lambda self: self.slab.global_shape)
    dt = property(lambda self: self.slab.dt)
    dx = property(lambda self: self.slab.dx)
    c = property(lambda self: self.slab.c)
    rho = property(lambda self: self.slab.rho)
    time = property(lambda self: self.slab.time)
    step_count = property(lambda self: self.slab.step_count)
    grid = property(lambda self: self.slab.grid)
    backend = "b200"
    using_native = False
    using_gpu = True
    using_b200 = True
    has_materials = property(lambda self: self.slab.has_materials)
    material_count = property(lambda self: self.slab.material_count)
    _sources = property(lambda self: self.slab._sources)
    _probes = property(lambda self: self.slab._probes)

    def set_geometry(self, geometry):
        """A global bool array, an SDF object (voxelised over this slab's planes only) or ``f(i_lo, i_hi)``."""
        self.slab.set_geometry(geometry)

    def add_boundary(self, b):
        self.slab.add_boundary(b)

    def add_source(self, s):
        self.slab.add_source(s)

    def add_probe(self, name, position):
        self.slab.add_probe(name, position)

    def add_microphone(self, position, name=None, pattern="omni", direction=None, up=None):
        return self.slab.add_microphone(position, name=name, pattern=pattern, direction=direction, up=up)

    microphones = property(lambda self: self.slab.microphones)

    def register_material(self, material, material_id=None):
        return self.slab.register_material(material, material_id=material_id)

    def set_material_region(self, mask, material_id):
        self.slab.set_material_region(mask, material_id)

    def set_material_box(self, material_id, x_range, y_range, z_range):
        self.slab.set_material_box(material_id, x_range, y_range, z_range)

    def get_material_at(self, position):
        """Material of a global cell (asked of its owner; every rank gets the answer)."""
        i = int(position[0])
        owner = owner_of(i, self.ranges)
        box = [self.slab.get_material_at((i - self.slab._i0,) + tuple(position[1:])) if owner == self.rank else None]
        self.dist.broadcast_object_list(box, src=self._peer(owner), group=self.group)
        return box[0]

    def set_kernel_option(self, opt, val):
        self.slab.set_kernel_option(opt, val)

    def enable_snapshots(self, interval: int, capture_velocity: bool = False) -> None:
        self._snapshot_interval = int(interval)
        self._snapshot_velocity = bool(capture_velocity)

    def get_velocity_snapshots(self):
        """(time, vx, vy, vz) at the cell centres of the whole grid -- kept on rank 0 only, like the pressure snapshots."""
        return self._velocity_snapshots

    def get_snapshots(self):
        """(time, global p) pairs -- kept on rank 0 only (the other ranks return an empty list)."""
        return self._snapshots

    def _finish_microphones(self):
        """Collective: every rank receives the corner samples of all slabs and completes every microphone."""
        mics = list(self.slab._microphones.values())
        if not mics:
            return
        local, times = _drain_corner_samples(self.slab)
        parts = [None] * self.world
        self.dist.all_gather_object(parts, (local, times), group=self.group)
        merged = {}
        for part, t in parts:
            merged.update(part)
            if len(t):
                times = t
        combine_corner_samples(mics, self.slab._mic_gathers or self.slab.microphone_gathers(mics), merged, times)

    # ---- halo exchange ---------------------------------------------------------------------------
    def _exchange(self, fields=("p",), include_vx_ghost=False, next_set=False, stream=None):
        """One grouped send/recv per step: p planes both ways (and, once after uploads, vx upward)."""
        dist, s = self.dist, self.slab
        dev = s._dev
        torch = dev.torch
        pairs = []                                # (send view or None, recv view or None, peer)
        h = s.halo_planes("p", next_set=next_set)
        if self.rank > 0:
            pairs.append((h["send_lo"], h["recv_lo"], self.rank - 1))
        if self.rank < self.world - 1:
            pairs.append((h["send_hi"], h["recv_hi"], self.rank + 1))
        if include_vx_ghost:
            v = s.halo_planes("vx")
            if self.rank < self.world - 1:
                pairs.append((v["send_hi"], None, self.rank + 1))
            if self.rank > 0:
                pairs.append((None, v["recv_lo"], self.rank - 1))
        if not pairs:
            return
        with torch.cuda.stream(stream if stream is not None else dev.stream):
            ops, back = [], []
            for send, recv, peer in pairs:
                if send is not None:
                    ops.append(dist.P2POp(dist.isend, send.cpu() if self._staged else send, self._peer(peer), self.group))
                if recv is not None:
                    buf = torch.empty(recv.shape, dtype=recv.dtype) if self._staged else recv
                    ops.append(dist.P2POp(dist.irecv, buf, self._peer(peer), self.group))
                    if self._staged:
                        back.append((recv, buf))
            for w in dist.batch_isend_irecv(ops):
                w.wait()                          # stream-level wait on CUDA; blocking on gloo
            for recv, buf in back:
                recv.copy_(buf)

    def _peer(self, r: int) -> int:
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def _all_any(self, flag: bool) -> bool:
        """Collective OR of a per-rank flag (one 4-byte all-reduce: this sits at the start of every run())."""
        torch = self.slab._ensure_device().torch
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cpu" if self._staged else self.slab._dev.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return bool(t.item())

    def _prepare_ghosts(self):
        """Collective.  After host-side edits of the fields (initial conditions) or a reset the ghost planes of p and
        the redundantly kept ghost face vx[-1] are refreshed from their owners."""
        s = self.slab
        stale = (not self._ghosts_fresh) or bool(s._host_dirty)
        s._sync_to_device()
        if not self._all_any(stale):
            return
        if self.halo == "p2p":
            _lib.check(s._dev.lib.sb_synchronize(s._dev.handle))
            self.dist.barrier(self.group)    # nobody is still pushing into the ghosts we are about to fill
        self._exchange(include_vx_ghost=True)
        if self.halo == "p2p":
            _lib.check(s._dev.lib.sb_synchronize(s._dev.handle))
            self.dist.barrier(self.group)
        self._ghosts_fresh = True

    # ---- stepping --------------------------------------------------------------------------------
    def _overlap_possible(self) -> bool:
        """Collective (cached until the set-up changes).  The send/recv of a step can run beside its interior update when
        K1 is the last writer of the cut planes on EVERY rank: no source on a cut plane here, and the library agrees
        (no ADE fix-ups, no Mur / radiation planes, slabs of at least 32 planes)."""
        s = self.slab
        key = (len(s._sources), len(s._boundaries), s.has_materials, s._options.get(_lib.OPT_KERNEL, 0))
        if getattr(self, "_overlap_key", None) != key:
            mine = (not self._staged) and self.world > 1 and not s.sources_touch_cut_planes() and \
                   not s.has_materials and s.shape[0] >= 32
            self._overlap, self._overlap_key = not self._all_any(not mine), key
        return self._overlap

    def _step_with_exchange(self) -> None:
        """NCCL mode, one step: cut planes first, then their send/recv on a second stream while the interior runs."""
        s = self.slab
        dev = s._dev
        torch = dev.torch
        if self._overlap and s.enqueue_cuts():
            if getattr(self, "_comm", None) is None:
                self._comm = torch.cuda.Stream(device=dev.device)
                self._ev_cut, self._ev_halo = torch.cuda.Event(), torch.cuda.Event()
            self._ev_cut.record(dev.stream)
            self._comm.wait_event(self._ev_cut)
            self._exchange(next_set=True, stream=self._comm)      # the cut planes of the set this step writes
            self._ev_halo.record(self._comm)
            s.enqueue_step()                                      # interior + sources + probes; flips the sets
            dev.stream.wait_event(self._ev_halo)                  # the next step reads the ghosts
        else:
            s.enqueue_step()
            self._exchange()

    def _run_chunk(self, m: int) -> None:
        s = self.slab
        if self.halo != "p2p":
            self._overlap_possible()             # (collective: outside the step loop)
        s.begin_chunk(m)
        if self.halo == "p2p":               # the kernels exchange halos themselves: enqueue the whole chunk
            s.enqueue_steps(m)
        else:
            for _ in range(m):
                self._step_with_exchange()
        s.end_chunk()

    def run(self, duration=None, progress=False, track_energy=False, energy_sample_interval=1, output_file=None,
            script_content=None, callback=None, snapshot_interval=None, steps=None, output=None):
        """``FDTDSolver.run`` (core/solver.py:2520-2606) over all slabs.  Collective: every rank calls it with the same
        arguments.  ``callback(step)`` runs on every rank; the result file and snapshots are produced by rank 0 from
        the traces / planes their owners send at the end of every chunk of steps."""
        import time as _t
        if steps is None:
            if duration is None:
                raise ValueError("run() needs duration= or steps=")
            steps = int(np.ceil(duration / self.dt))
        output_file = output_file or output
        s = self.slab
        t0 = _t.time()
        self._prepare_ghosts()
        if track_energy and not self._energy_history:
            self._energy_history.append((s._step_count, s._time, self.compute_energy()))
        writer = None
        if output_file and self.rank == 0:
            from .io import ResultWriter
            writer = ResultWriter(output_file, _WriterView(self), script_content)
        elif output_file:
            _WriterView(self).geometry                  # takes part in the collective gather of the geometry
        bar = None
        if progress and self.rank == 0:
            try:
                from tqdm import tqdm
                bar = tqdm(total=steps, desc=f"FDTD simulation (b200 x{self.world})")
            except ImportError:
                bar = None
        launches0 = s.kernel_launches()
        done = 0
        try:
            while done < steps:
                m = min(self.chunk_steps, steps - done)
                for q in range(m):                     # a chunk ends right after any step whose fields the host has to see
                    idx = s._step_count + q
                    if (self._snapshot_interval and idx % self._snapshot_interval == 0) or \
                            (track_energy and (idx + 1) % energy_sample_interval == 0) or \
                            (output_file and snapshot_interval is not None and (done + q) % snapshot_interval == 0):
                        m = q + 1
                        break
                n0 = {pr.name: len(pr.data) for pr in s._local_probes}
                self._run_chunk(m)
                last_idx = s._step_count - 1
                if self._snapshot_interval and last_idx % self._snapshot_interval == 0:
                    p = self.gather_field("p")
                    v = [self.gather_field(f) for f in ("vx", "vy", "vz")] if self._snapshot_velocity else None
                    if self.rank == 0:
                        self._snapshots.append((float(s._chunk_times[-1]), p))
                        if v is not None:
                            from .solver import centre_velocities
                            self._velocity_snapshots.append((float(s._chunk_times[-1]), *centre_velocities(*v)))
                if track_energy and s._step_count % energy_sample_interval == 0:
                    self._energy_history.append((s._step_count, s._time, self.compute_energy()))
                if output_file:
                    mine = {pr.name: np.asarray(pr.data[n0[pr.name]:], dtype=np.float32) for pr in s._local_probes}
                    parts = [None] * self.world if self.rank == 0 else None
                    self.dist.gather_object(mine, parts, dst=self._peer(0), group=self.group)
                    snap = snapshot_interval is not None and (done + m - 1) % snapshot_interval == 0
                    p = self.gather_field("p") if snap else None
                    if writer is not None:
                        merged = {}
                        for part in parts:
                            merged.update(part)
                        names = [n for n in s._probes if n in merged]
                        if names:
                            writer.append_probe_block(names, np.stack([merged[n] for n in names], axis=1))
                        if snap:
                            writer.write_snapshot(p)
                if bar is not None:
                    bar.update(m)
                if callback is not None:
                    for q in range(m):
                        callback(last_idx - (m - 1 - q))
                done += m
            self._finish_microphones()
            if s._warn_energy_drift and len(self._energy_history) >= 2:
                import warnings
                rep = self.energy_report()
                if abs(rep["energy_change_percent"]) > s._energy_drift_threshold * 100:
                    warnings.warn(f"Energy drift detected: {rep['energy_change_percent']:.2f}% change "
                                  f"(threshold: {s._energy_drift_threshold * 100:.1f}%). "
                                  f"Status: {rep['conservation_status']}", UserWarning, stacklevel=2)
        finally:
            runtime = _t.time() - t0
            if bar is not None:
                bar.close()
            cells = int(np.prod(self.shape, dtype=np.int64))
            self.last_run_stats = {"steps": steps, "runtime_s": runtime, "n_gpus": self.world, "halo": self.halo,
                                   "cell_updates_per_s": cells * steps / runtime if runtime > 0 else float("inf"),
                                   "kernel_launches": s.kernel_launches() - launches0}
            if writer is not None:
                writer.finalize(runtime=runtime, backend="b200", num_threads=0, num_gpus=self.world)
            if output_file:
                self.dist.barrier(self.group)        # the file exists on return, on every rank (scripts stat it right away)

    def step(self):
        self.run(steps=1)

    def get_state(self) -> dict:
        """This rank's shard of the solver state (FDTDSolver.get_state of its slab plus the slab's place in the grid): every
        rank saves its own, e.g. ``np.savez(f"ckpt_{rank}.npz", ...)``.  No communication."""
        st = self.slab.get_state()
        st["world"], st["rank"], st["i_range"] = self.world, self.rank, (int(self.slab._i0), int(self.slab._i1))
        return st

    def set_state(self, state: dict) -> None:
        """Collective.  Restores a shard written by ``get_state`` under the same decomposition; the ghost planes are
        refreshed from their owners before the next step.  Dispersive materials are refused here: their auxiliary fields
        on the ghost planes belong to the neighbour's shard, which this call does not see."""
        if tuple(state.get("i_range", ())) != (int(self.slab._i0), int(self.slab._i1)) or state.get("world") != self.world:
            raise ValueError(f"state of planes {state.get('i_range')} in a job of {state.get('world')} ranks does not fit this "
                             f"rank's planes {(int(self.slab._i0), int(self.slab._i1))} in a job of {self.world}")
        if state.get("ade"):
            raise NotImplementedError("resuming dispersive materials on a decomposed grid (use a single-GPU checkpoint)")
        if self.slab._dev is not None:
            _lib.check(self.slab._dev.lib.sb_synchronize(self.slab._dev.handle))
        self.dist.barrier(self.group)            # no neighbour is still storing into our ghosts
        self.slab.set_state(state)
        self._ghosts_fresh = False
        self.dist.barrier(self.group)

    def reset(self) -> None:
        """Back to t = 0 on every slab (core/solver.py:2781-2800).  Collective."""
        s = self.slab
        if s._dev is not None:
            _lib.check(s._dev.lib.sb_synchronize(s._dev.handle))
        self.dist.barrier(self.group)            # no neighbour is still storing into our ghosts / flags
        s.reset()
        self.dist.barrier(self.group)
        self._ghosts_fresh = False
        self._energy_history.clear()
        self._snapshots.clear()
        self._velocity_snapshots.clear()

    # ---- results ---------------------------------------------------------------------------------
    def get_probe_data(self, name=None) -> dict:
        """All probes on every rank (gathered from their owners)."""
        if name is not None and name not in self.slab._probes:
            raise KeyError(f"Probe '{name}' not found")
        local = {pr.name: pr.get_data() for pr in self.slab._local_probes}
        parts = [None] * self.world
        self.dist.all_gather_object(parts, local, group=self.group)
        merged = {}
        for part in parts:
            merged.update(part)
        out = {n: merged[n] for n in self.slab._probes if n in merged}
        return out if name is None else {name: out[name]}

    def gather_field(self, name: str):
        """The whole field on rank 0 (None elsewhere) -- for tests and small grids."""
        local = self.slab.get_field(name)
        parts = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(local, parts, dst=self._peer(0), group=self.group)
        return np.concatenate(parts, axis=0) if self.rank == 0 else None

    def get_field(self, name: str) -> np.ndarray:
        """The whole field on every rank (a copy; small grids)."""
        parts = [None] * self.world
        self.dist.all_gather_object(parts, self.slab.get_field(name), group=self.group)
        return np.concatenate(parts, axis=0)

    def set_field(self, name: str, value) -> None:
        """Overwrite a field from a global array (initial conditions); ghosts are refreshed by the next run()."""
        value = np.asarray(value)
        if value.shape != tuple(self.shape):
            raise ValueError(f"field shape {value.shape} doesn't match solver shape {self.shape}")
        self.slab._field_set(name, value[self.slab._i0:self.slab._i1])
        self._ghosts_fresh = False

    p = property(lambda s: s.get_field("p"), lambda s, v: s.set_field("p", v))
    vx = property(lambda s: s.get_field("vx"), lambda s, v: s.set_field("vx", v))
    vy = property(lambda s: s.get_field("vy"), lambda s, v: s.set_field("vy", v))
    vz = property(lambda s: s.get_field("vz"), lambda s, v: s.set_field("vz", v))

    def compute_energy(self) -> float:
        import torch
        e = torch.tensor([self.slab.compute_energy()], dtype=torch.float64,
                         device="cpu" if self._staged else self.slab._dev.device)
        self.dist.all_reduce(e, group=self.group)
        return float(e.item())

    def get_energy_history(self):
        return self._energy_history.copy()

    def energy_report(self) -> dict:
        return FDTDSolver.energy_report(self)

    def get_sample_rate(self) -> float:
        return 1.0 / self.dt

    def get_frequency_response(self, probe_name: str, n_fft=None):
        data = self.get_probe_data(probe_name)[probe_name]
        if n_fft is None:
            n_fft = int(2 ** np.ceil(np.log2(len(data))))
        return np.fft.rfftfreq(n_fft, self.dt), np.abs(np.fft.rfft(data, n=n_fft))

    def kernel_launches(self) -> int:
        return self.slab.kernel_launches()

    def device_stats(self) -> dict:
        return self.slab.device_stats()

    def close(self):
        self.slab.close()


class _WriterView:
    """What io.ResultWriter reads from a solver, with global extents (rank 0 writes the file)."""

    def __init__(self, d: DistributedFDTDSolver):
        self._d = d
        self._geometry = None

    def __getattr__(self, name):
        return getattr(self._d, name)

    @property
    def geometry(self):
        """Collective: the global air mask gathered on rank 0 (all air if nobody set a geometry)."""
        if self._geometry is None:
            d = self._d
            local = None if d.slab._geometry is None else np.packbits(d.slab._geometry)
            parts = [None] * d.world if d.rank == 0 else None
            d.dist.gather_object((local, d.slab.shape), parts, dst=d._peer(0), group=d.group)
            if d.rank == 0:
                planes = []
                for bits, shp in parts:
                    n = int(np.prod(shp, dtype=np.int64))
                    planes.append(np.ones(shp, dtype=bool) if bits is None else
                                  np.unpackbits(bits)[:n].reshape(shp).astype(bool))
                self._geometry = np.concatenate(planes, axis=0)
            else:
                self._geometry = False
        return self._geometry


class LocalSlabGroup:
    """N slabs of one problem inside one process / one device, stepped in lock step.

    Exercises exactly the ghost-plane logic of the distributed path (same kernels, same tables, same
    exchange order) where only one GPU is available; peer copies become device-to-device copies.
    """

    def __init__(self, n_slabs: int, shape=None, resolution=None, grid=None, device=None, chunk_steps=32,
                 halo="copy", **kw):
        nx = int((grid.shape if grid is not None else shape)[0])
        self.ranges = slab_ranges(nx, n_slabs)
        self.slabs = [SlabSolver(shape=shape, resolution=resolution, grid=grid, device=device,
                                 chunk_steps=chunk_steps, slab=r, **kw) for r in self.ranges]
        self.chunk_steps = int(chunk_steps)
        self._ghosts_fresh = False
        self.halo = halo
        self._flags = []

    def _setup_p2p(self):
        """Same peer-store / flag protocol as the distributed path, with the 'peers' on the same device."""
        import ctypes as C
        if self._flags:
            return
        devs = [s._ensure_device() for s in self.slabs]
        torch = devs[0].torch
        self._flags = [torch.zeros(8, dtype=torch.int32, device=d.device) for d in devs]
        torch.cuda.synchronize()
        n = len(self.slabs)
        for r, (s, d) in enumerate(zip(self.slabs, devs)):
            lo, hi = r - 1, r + 1
            lo_sets = (C.c_void_p * 2)(*[devs[lo].sets[q][0].data_ptr() if lo >= 0 else None for q in range(2)])
            hi_sets = (C.c_void_p * 2)(*[devs[hi].sets[q][0].data_ptr() if hi < n else None for q in range(2)])
            lo_nx = self.slabs[lo].shape[0] if lo >= 0 else 0
            lo_flag = C.c_void_p(self._flags[lo].data_ptr() + 4) if lo >= 0 else None
            hi_flag = C.c_void_p(self._flags[hi].data_ptr()) if hi < n else None
            _lib.check(d.lib.sb_set_peers(d.handle, lo_sets, hi_sets, lo_nx, C.c_void_p(self._flags[r].data_ptr()),
                                          lo_flag, hi_flag))

    def for_all(self, fn):
        for s in self.slabs:
            fn(s)

    def _sync(self):
        for s in self.slabs:
            _lib.check(s._dev.lib.sb_synchronize(s._dev.handle))

    def _exchange(self, include_vx_ghost=False, next_set=False):
        self._sync()
        for lo, hi in zip(self.slabs[:-1], self.slabs[1:]):
            a, b = lo.halo_planes("p", next_set=next_set), hi.halo_planes("p", next_set=next_set)
            b["recv_lo"].copy_(a["send_hi"])
            a["recv_hi"].copy_(b["send_lo"])
            if include_vx_ghost:
                hi.halo_planes("vx")["recv_lo"].copy_(lo.halo_planes("vx")["send_hi"])
        self.slabs[0]._dev.torch.cuda.synchronize()

    def run(self, steps: int):
        for s in self.slabs:
            s._sync_to_device()
        if not self._ghosts_fresh:
            self._exchange(include_vx_ghost=True)
            self._ghosts_fresh = True
        if self.halo == "p2p":
            self._setup_p2p()
        for m in _chunks(steps, self.chunk_steps):
            for s in self.slabs:
                s.begin_chunk(m)
            if self.halo == "p2p":
                # One device runs every slab here, so a slab's spinning cut blocks wait for kernels of ANOTHER stream
                # of the same GPU.  Submitting step by step, slab after slab, keeps every kernel a waiter depends on
                # ahead of it in submission order (whole chunks per slab can leave the neighbour's kernels queued
                # behind the spinner).  With one GPU per slab (DistributedFDTDSolver) no such dependency exists.
                for _ in range(m):
                    for s in self.slabs:
                        s.enqueue_step()
            elif self.halo == "copy_cuts":
                # the order of the overlapped NCCL mode (cut planes first, exchange, then the rest of the step),
                # serialised on one device: exercises sb_step_cuts_async where only one GPU is available
                ok = not any(s.sources_touch_cut_planes() for s in self.slabs)
                for _ in range(m):
                    split = ok and all([s.enqueue_cuts() for s in self.slabs])
                    if split:
                        self._exchange(next_set=True)
                    for s in self.slabs:
                        s.enqueue_step()
                    if not split:
                        self._exchange()
                self.cut_steps = getattr(self, "cut_steps", 0) + (m if split else 0)
            else:
                for _ in range(m):
                    for s in self.slabs:
                        s.enqueue_step()
                    self._exchange()
            for s in self.slabs:
                s.end_chunk()
        mics = list(self.slabs[0]._microphones.values())
        if mics:                                  # every slab holds the same microphone objects' twins; finish slab 0's
            merged, times = {}, np.zeros(0)
            for s in self.slabs:
                part, t = _drain_corner_samples(s)
                merged.update(part)
                times = t if len(t) else times
            combine_corner_samples(mics, self.slabs[0]._mic_gathers, merged, times)

    def add_microphone(self, position, name=None, **kw):
        return [s.add_microphone(position, name=name, **kw) for s in self.slabs][0]

    microphones = property(lambda self: self.slabs[0].microphones)

    def get_field(self, name: str) -> np.ndarray:
        return np.concatenate([s.get_field(name) for s in self.slabs], axis=0)

    def get_probe_data(self) -> dict:
        out = {}
        for s in self.slabs:
            for pr in s._local_probes:
                out[pr.name] = pr.get_data()
        return out

    def close(self):
        for s in self.slabs:
            s.close()
