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

    def halo_planes(self, field: str = "p") -> dict:
        """torch views of the planes that take part in the exchange, in the CURRENT set."""
        dev = self._dev
        t = dev.sets[dev.current_set()][_FIELDS.index(field)]
        nx = self.shape[0]
        return {"send_lo": t[1], "send_hi": t[nx], "recv_lo": t[0], "recv_hi": t[nx + 1]}


def _chunks(n_steps: int, chunk: int):
    done = 0
    while done < n_steps:
        m = min(chunk, n_steps - done)
        yield m
        done += m


class DistributedFDTDSolver:
    """One slab per process; the public surface of FDTDSolver with global coordinates.

    Launch with ``python -m torch.distributed.run --nproc-per-node N ...`` (one rank per GPU).  The
    process group must exist (``torch.distributed.init_process_group``) before construction.
    """

    def __init__(self, shape=None, resolution=None, grid=None, c=343.0, rho=1.2, courant=0.95,
                 backend="b200", device=None, chunk_steps=64, group=None, halo="auto"):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if halo not in ("auto", "p2p", "nccl"):
            raise ValueError("halo must be 'auto', 'p2p' or 'nccl'")
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nx = int((grid.shape if grid is not None else shape)[0])
        self.ranges = slab_ranges(nx, self.world)
        self.slab = SlabSolver(shape=shape, resolution=resolution, grid=grid, c=c, rho=rho, courant=courant,
                               backend=backend, device=device, chunk_steps=chunk_steps, slab=self.ranges[self.rank])
        self.chunk_steps = int(chunk_steps)
        self._ghosts_fresh = False
        self.halo = "nccl"
        self._symm = []                          # keeps symmetric allocations / handles alive
        if halo in ("auto", "p2p") and self.world > 1 and dist.get_backend(group) == "nccl":
            try:
                self._setup_p2p()
                self.halo = "p2p"
            except Exception as e:               # symmetric memory unavailable -> host-driven NCCL exchange
                if halo == "p2p":
                    raise
                import warnings
                warnings.warn(f"peer-to-peer halo unavailable ({type(e).__name__}: {e}); using NCCL send/recv")
                self.slab._p_allocator = None

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
    time = property(lambda self: self.slab.time)
    step_count = property(lambda self: self.slab.step_count)
    grid = property(lambda self: self.slab.grid)

    def set_geometry(self, geometry):
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

    def set_kernel_option(self, opt, val):
        self.slab.set_kernel_option(opt, val)

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
    def _exchange(self, fields=("p",), include_vx_ghost=False):
        """One grouped send/recv per step: p planes both ways (and, once after uploads, vx upward)."""
        dist, s = self.dist, self.slab
        dev = s._dev
        ops = []
        with dev.torch.cuda.stream(dev.stream):
            h = s.halo_planes("p")
            if self.rank > 0:
                ops.append(dist.P2POp(dist.isend, h["send_lo"], self._peer(self.rank - 1), self.group))
                ops.append(dist.P2POp(dist.irecv, h["recv_lo"], self._peer(self.rank - 1), self.group))
            if self.rank < self.world - 1:
                ops.append(dist.P2POp(dist.isend, h["send_hi"], self._peer(self.rank + 1), self.group))
                ops.append(dist.P2POp(dist.irecv, h["recv_hi"], self._peer(self.rank + 1), self.group))
            if include_vx_ghost:
                v = s.halo_planes("vx")
                if self.rank < self.world - 1:
                    ops.append(dist.P2POp(dist.isend, v["send_hi"], self._peer(self.rank + 1), self.group))
                if self.rank > 0:
                    ops.append(dist.P2POp(dist.irecv, v["recv_lo"], self._peer(self.rank - 1), self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()                      # stream-level wait on CUDA; blocking on gloo

    def _peer(self, r: int) -> int:
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    # ---- stepping --------------------------------------------------------------------------------
    def run(self, duration=None, steps=None, **_ignored):
        if steps is None:
            steps = int(np.ceil(duration / self.dt))
        s = self.slab
        s._sync_to_device()
        if not self._ghosts_fresh:               # ghosts of p and vx after host-side edits / first use
            if self.halo == "p2p":
                _lib.check(s._dev.lib.sb_synchronize(s._dev.handle))
                self.dist.barrier(self.group)    # nobody is still pushing into the ghosts we are about to fill
            self._exchange(include_vx_ghost=True)
            if self.halo == "p2p":
                _lib.check(s._dev.lib.sb_synchronize(s._dev.handle))
                self.dist.barrier(self.group)
            self._ghosts_fresh = True
        for m in _chunks(steps, self.chunk_steps):
            s.begin_chunk(m)
            if self.halo == "p2p":               # the kernels exchange halos themselves: enqueue the whole chunk
                s.enqueue_steps(m)
            else:
                for _ in range(m):
                    s.enqueue_step()
                    self._exchange()
            s.end_chunk()
        self._finish_microphones()

    def step(self):
        self.run(steps=1)

    # ---- results ---------------------------------------------------------------------------------
    def get_probe_data(self, name=None) -> dict:
        """All probes on every rank (gathered from their owners)."""
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

    def compute_energy(self) -> float:
        import torch
        e = torch.tensor([self.slab.compute_energy()], dtype=torch.float64,
                         device=self.slab._dev.device if self.dist.get_backend(self.group) == "nccl" else "cpu")
        self.dist.all_reduce(e, group=self.group)
        return float(e.item())

    def close(self):
        self.slab.close()


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

    def _exchange(self, include_vx_ghost=False):
        self._sync()
        for lo, hi in zip(self.slabs[:-1], self.slabs[1:]):
            a, b = lo.halo_planes("p"), hi.halo_planes("p")
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
