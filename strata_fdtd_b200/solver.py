"""``FDTDSolver`` with ``backend="b200"``: the reference's solver surface, driven on one B200.

Mirrors the public interface of /root/reference/src/strata_fdtd/core/solver.py:1442-3442
(constructor :1507-1518, set_geometry :1682, add_source :1782, add_probe :1843,
add_microphone :1887, add_boundary :1983, step :2003, run :2520, get_probe_data :2608,
compute_energy :2689, reset :2781, materials :2836-2942) so that scripts written for the
reference run unchanged.  What differs is where the work happens: the whole step --
velocity, rigid faces, pressure, sponge, ADE, source injection, probe/microphone recording --
runs on the device in chunks of steps through the C ABI of include/strata_b200.h; the host
only evaluates source waveforms (float64, as the reference does) and collects traces.

PyTorch is used for nothing but device buffers and the stream handle.  There is no CPU
fallback: without the CUDA library or a CUDA device, construction of the device state raises.
"""
from __future__ import annotations

import ctypes as C
import time as _time_mod
import warnings
from collections.abc import Callable

import numpy as np

from . import _lib
from .boundaries import plane_ops, sponge_tables
from .grid import NonuniformGrid, UniformGrid
from .sources import Microphone, Probe, combine_corner_samples

_FIELDS = ("p", "vx", "vy", "vz")


class _DeviceState:
    """Owns the torch buffers and the sb_solver handle of one slab."""

    def __init__(self, shape, device_index: int | None, slab=None, p_allocator=None):
        import torch
        if not torch.cuda.is_available():
            raise _lib.B200BackendError("backend='b200' needs a CUDA device; none is visible (no CPU fallback)")
        self.torch = torch
        self.lib = _lib.load()
        self.index = torch.cuda.current_device() if device_index is None else int(device_index)
        self.device = torch.device("cuda", self.index)
        nx, ny, nz = shape
        has_lower, has_upper, global_nx, i_offset = slab or (0, 0, nx, 0)
        self.desc = _lib.GridDesc(nx=nx, ny=ny, nz=nz, pitch=0, global_nx=global_nx, i_offset=i_offset,
                                  has_lower=has_lower, has_upper=has_upper)
        pitch = C.c_int32(0)
        _lib.check(self.lib.sb_choose_pitch(nz, C.byref(pitch)))
        self.desc.pitch = pitch.value
        self.pitch = pitch.value
        with torch.cuda.device(self.index):
            # a dedicated non-default stream: the legacy default stream cannot be captured into CUDA graphs
            self.stream = torch.cuda.Stream(device=self.device)
            # two ping-pong sets of {p, vx, vy, vz}; planes -1 and nx are the slab ghosts
            self.sets = [[torch.zeros((nx + 2, ny, self.pitch), dtype=torch.float32, device=self.device)
                          for _ in range(4)] for _ in range(2)]
            if p_allocator is not None:
                # peer-visible p buffers (symmetric memory) for the NVLink halo path; may be over-sized
                for q in range(2):
                    self.sets[q][0] = p_allocator(nx + 2, ny, self.pitch, self.device)
            torch.cuda.synchronize(self.device)
            handle = C.c_void_p()
            _lib.check(self.lib.sb_create(C.byref(self.desc), self.index, C.c_void_p(self.stream.cuda_stream),
                                          C.byref(handle)))
        self.handle = handle
        arr = lambda ts: (C.c_void_p * 4)(*[t.data_ptr() for t in ts])
        _lib.check(self.lib.sb_bind_fields(self.handle, arr(self.sets[0]), arr(self.sets[1])))

    def current_set(self) -> int:
        s = C.c_int(0)
        _lib.check(self.lib.sb_current_set(self.handle, C.byref(s)))
        return s.value

    def field_view(self, name: str):
        """torch view [nx, ny, nz] of the current device field (no copy)."""
        t = self.sets[self.current_set()][_FIELDS.index(name)]
        return t[1:self.desc.nx + 1, :, : self.desc.nz]

    def close(self):
        if self.handle:
            self.lib.sb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_options():
    """Process-group options for the halo exchange: NCCL's kernels on a high-priority stream, so that the few CTAs of a
    send/recv get SM slots beside the step kernel's thousands of blocks instead of after them (None if unavailable)."""
    try:
        import torch.distributed as dist
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        return opts
    except Exception:
        return None


def distributed_world(auto_init: bool = True) -> int:
    """Number of ranks this process steps a grid with: the size of the torch.distributed job it runs in, else 1.

    Under ``python -m torch.distributed.run`` (RANK / WORLD_SIZE / MASTER_ADDR in the environment) the process group is
    created on first use: NCCL, one rank per GPU, rank r on CUDA device ``STRATA_B200_DEVICES[r]`` (a comma list;
    default LOCAL_RANK).  ``STRATA_B200_DISTRIBUTED=0`` switches the automatic route off."""
    import os
    if os.environ.get("STRATA_B200_DISTRIBUTED", "1") == "0":
        return 1
    try:
        import torch.distributed as dist
    except ImportError:
        return 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and auto_init and "RANK" in os.environ and "MASTER_ADDR" in os.environ:
        import torch
        local = int(os.environ.get("LOCAL_RANK", "0"))
        devs = [int(q) for q in os.environ.get("STRATA_B200_DEVICES", "").split(",") if q.strip() != ""]
        index = devs[local] if devs else local
        if torch.cuda.is_available():
            torch.cuda.set_device(index)
            dist.init_process_group("nccl", device_id=torch.device("cuda", index), pg_options=nccl_options())
        else:
            dist.init_process_group("gloo")
        return dist.get_world_size()
    return 1


def centre_velocities(vx, vy, vz):
    """Face velocities averaged to the cell centres, as the reference's velocity snapshots hold them: 0.5 * (v[c - 1] + v[c])
    along the field's own axis, with a zero face below the first cell."""
    out = []
    for axis, v in enumerate((vx, vy, vz)):
        c = np.zeros_like(v)
        lo = [slice(None)] * 3; hi = [slice(None)] * 3; first = [slice(None)] * 3
        lo[axis] = slice(0, -1); hi[axis] = slice(1, None); first[axis] = 0
        c[tuple(hi)] = 0.5 * (v[tuple(lo)] + v[tuple(hi)])
        c[tuple(first)] = 0.5 * v[tuple(first)]
        out.append(c)
    return out


class FDTDSolver:
    """3-D acoustic pressure-velocity FDTD on a staggered grid, executed on a B200.

    Constructed inside a multi-rank job (see ``distributed_world``) the same call returns a
    ``multi.DistributedFDTDSolver``: the grid is cut into slabs along axis 0, one per GPU, behind the same methods
    (``distributed=False`` keeps a solver on this rank's GPU only, ``True`` insists on the slab solver)."""

    _is_slab = False            # multi.SlabSolver: one slab of a decomposed grid, never re-dispatched
    _coerce_backend = False     # compat alias: "native" / "python" requests are served by this backend

    def __new__(cls, *args, distributed: bool | None = None, **kw):
        if not cls._is_slab and kw.get("slab") is None and distributed is not False:
            if distributed_world() > 1:
                import os
                from .multi import DistributedFDTDSolver
                kw.pop("slab", None)
                devices = kw.pop("devices", None)
                if devices:                                   # rank r of the node steps its slab on devices[r]
                    kw["device"] = int(devices[int(os.environ.get("LOCAL_RANK", "0")) % len(devices)])
                if cls._coerce_backend:
                    kw["backend"] = "b200"
                return DistributedFDTDSolver(*args, **kw)
            if distributed:
                raise RuntimeError("distributed=True needs a torch.distributed job with more than one rank "
                                   "(python -m torch.distributed.run --nproc-per-node N ...)")
        return super().__new__(cls)

    def __init__(self, shape=None, resolution=None, grid=None, c: float = 343.0, rho: float = 1.2,
                 courant: float = 0.95, backend: str = "b200", warn_energy_drift: bool = False,
                 energy_drift_threshold: float = 0.01, device: int | None = None, chunk_steps: int | None = None,
                 slab: tuple[int, int] | None = None, distributed: bool | None = None, devices=None):
        if devices:
            if len(devices) > 1:
                raise RuntimeError(f"devices={list(devices)}: one process drives one GPU; start one rank per device "
                                   "(python -m torch.distributed.run --nproc-per-node N ...) and the same call cuts the "
                                   "grid into slabs")
            device = int(devices[0])
        if backend not in ("b200", "auto"):
            raise ValueError(f"strata_fdtd_b200 provides backend='b200' only (got {backend!r}); "
                             "use the reference package for 'native' / 'python'")
        if grid is not None:
            self._grid = grid
            self.global_shape = tuple(int(n) for n in grid.shape)
        elif shape is not None and resolution is not None:
            self._grid = UniformGrid(shape=shape, resolution=resolution)
            self.global_shape = tuple(int(n) for n in shape)
        else:
            raise ValueError("Must provide either 'grid' or both 'shape' and 'resolution'")
        # A slab [i0, i1) of the grid along axis 0 (multi-GPU decomposition); the default is the whole grid.
        # Everything the caller passes (positions, geometry) is global; storage and kernels are local.
        self._i0, self._i1 = (0, self.global_shape[0]) if slab is None else (int(slab[0]), int(slab[1]))
        if not 0 <= self._i0 < self._i1 <= self.global_shape[0]:
            raise ValueError(f"bad slab range {slab} for nx={self.global_shape[0]}")
        self._has_lower = int(self._i0 > 0)
        self._has_upper = int(self._i1 < self.global_shape[0])
        self.shape = (self._i1 - self._i0, self.global_shape[1], self.global_shape[2])
        self.dx = self._grid.min_spacing
        self.c, self.rho = c, rho
        self.backend = "b200"
        # dt and coefficients in float64, as core/solver.py:1576-1592
        self.dt = courant * (self._grid.min_spacing / (self.c * np.sqrt(3)))
        if self._grid.is_uniform:
            self._coeff_p = -self.rho * self.c**2 * self.dt / self.dx
            self._coeff_v = -self.dt / (self.rho * self.dx)
        else:
            self._coeff_p_base = -self.rho * self.c**2 * self.dt
            self._coeff_v_base = -self.dt / self.rho
            self._spacing_arrays = self._grid.get_spacing_arrays_for_stencil()

        self._host = {f: np.zeros(self.shape, dtype=np.float32) for f in _FIELDS}
        self._host_stale: set[str] = set()     # device holds newer data
        self._host_dirty: set[str] = set()     # host may hold newer data (array was handed out)
        self._geometry = None                  # None = all air; materialised on first access
        self._rigid = False                    # face zeroing only after set_geometry (solver.py:1779-1780)

        self._sources: list = []
        self._probes: dict[str, Probe] = {}
        self._microphones: dict[str, Microphone] = {}
        self._boundaries: list = []
        self._step_count = 0
        self._time = 0.0
        self._snapshots: list = []
        self._velocity_snapshots: list = []
        self._snapshot_interval: int | None = None
        self._snapshot_velocity = False
        self._warn_energy_drift = warn_energy_drift
        self._energy_drift_threshold = energy_drift_threshold
        self._energy_history: list = []
        self._track_energy = False
        self._energy_sample_interval = 1
        self._materials: dict = {}
        # material ids of the owned planes plus, on a slab, the live ghost planes next to them (as the geometry)
        self._material_ext = np.zeros((self.shape[0] + self._has_lower + self._has_upper,) + self.shape[1:], dtype=np.uint8)
        self._material_id = self._material_ext[self._has_lower: self._has_lower + self.shape[0]]
        self._local_probes: list = []
        self._mic_slots: list = []
        self._corner_keys: list = []
        self._mic_gathers: list = []
        self._corner_data: dict = {}
        self._corner_times: list = []
        self._geometry_ext = None

        self._device_index = device
        self._chunk_auto = chunk_steps is None           # the caller left the chunk length to the solver
        self._chunk_steps = 256 if chunk_steps is None else int(chunk_steps)
        self._dev: _DeviceState | None = None
        self._dirty = {"coeffs", "geometry", "sponges", "sources", "records", "ade"}
        self._options: dict[int, int] = {}
        self.last_run_stats: dict = {}

    # ------------------------------------------------------------------ properties
    time = property(lambda self: self._time)
    step_count = property(lambda self: self._step_count)
    using_native = property(lambda self: False)
    using_gpu = property(lambda self: True)
    using_b200 = property(lambda self: True)
    grid = property(lambda self: self._grid)
    microphones = property(lambda self: self._microphones)
    has_materials = property(lambda self: len(self._materials) > 0)
    material_count = property(lambda self: len(self._materials))

    def _field_get(self, name):
        if name in self._host_stale:
            self._download(name)
        self._host_dirty.add(name)             # caller may write through the returned array
        return self._host[name]

    def _field_set(self, name, value):
        self._host[name][...] = value
        self._host_stale.discard(name)
        self._host_dirty.add(name)

    p = property(lambda s: s._field_get("p"), lambda s, v: s._field_set("p", v))
    vx = property(lambda s: s._field_get("vx"), lambda s, v: s._field_set("vx", v))
    vy = property(lambda s: s._field_get("vy"), lambda s, v: s._field_set("vy", v))
    vz = property(lambda s: s._field_get("vz"), lambda s, v: s._field_set("vz", v))

    def get_field(self, name: str) -> np.ndarray:
        """Read-only copy of a field; unlike ``solver.p`` it does not force a re-upload."""
        if name in self._host_stale:
            self._download(name)
        return self._host[name].copy()

    @property
    def geometry(self):
        if self._geometry is None:
            self._geometry_ext = np.ones((self.shape[0] + self._has_lower + self._has_upper,) + self.shape[1:], dtype=bool)
            self._geometry = self._geometry_ext[self._has_lower: self._has_lower + self.shape[0]]
        self._dirty |= {"geometry", "sources"}  # may be edited in place by the caller
        return self._geometry

    @geometry.setter
    def geometry(self, value):
        self._store_geometry(value)

    def _store_geometry(self, mask) -> None:
        """Keep the owned planes (``_geometry``) and, on a slab, the live ghost planes next to them
        (``_geometry_ext``: the rigid-face test of the first/last owned plane looks across the cut).
        ``mask`` is a global array or, for grids too large to materialise, ``f(i_lo, i_hi) -> bool planes``."""
        lo, hi = self._i0 - self._has_lower, self._i1 + self._has_upper
        if callable(mask):
            ext = np.asarray(mask(lo, hi))
            want = (hi - lo,) + self.global_shape[1:]
            if ext.shape != want:
                raise ValueError(f"Geometry shape {ext.shape} doesn't match solver shape {want}")
        else:
            mask = np.asarray(mask)
            if mask.shape != self.global_shape:
                raise ValueError(f"Geometry shape {mask.shape} doesn't match solver shape {self.global_shape}")
            ext = mask[lo:hi]
        self._geometry_ext = np.ascontiguousarray(ext.astype(bool))
        self._geometry = self._geometry_ext[self._has_lower: self._has_lower + self.shape[0]]
        self._dirty |= {"geometry", "sources"}

    # ------------------------------------------------------------------ set-up API
    def set_geometry(self, geometry) -> None:
        """bool array (True = air), an SDF primitive, or a MaterializedGeometry (solver.py:1682-1780)."""
        if hasattr(geometry, "voxelize_with_materials"):
            mask, ids = geometry.voxelize_with_materials(self.grid)
            self._store_geometry(mask)
            for mat_id, material in geometry.get_material_table().items():
                if mat_id == 0:
                    continue
                if mat_id not in self._materials:
                    self.register_material(material, material_id=mat_id)
                region = ids == mat_id
                if np.any(region):
                    self.set_material_region(region, material_id=mat_id)
        elif hasattr(geometry, "sdf"):
            # SDFPrimitive.voxelize (geometry/sdf.py:99-125: cell-centre points, sdf <= 0) restricted to the planes this
            # slab holds and evaluated a few planes at a time -- never the whole-grid point array (6.4 GB at config 4)
            self._store_geometry(lambda lo, hi: self._voxelize_planes(geometry, lo, hi))
            self._material_ext.fill(0)
            self._dirty.add("ade")
        elif hasattr(geometry, "voxelize"):
            self._store_geometry(geometry.voxelize(self.grid))
            self._material_ext.fill(0)
            self._dirty.add("ade")
        else:
            self._store_geometry(geometry)
        self._rigid = True
        self._dirty |= {"geometry", "sources"}

    def _voxelize_planes(self, sdf_object, i_lo: int, i_hi: int, max_points: int = 1 << 22) -> np.ndarray:
        """``sdf_object.voxelize(grid)[i_lo:i_hi]`` without materialising the rest of the grid."""
        g = self._grid
        ny, nz = self.global_shape[1], self.global_shape[2]
        Y, Z = np.meshgrid(np.asarray(g.y_coords, dtype=np.float64), np.asarray(g.z_coords, dtype=np.float64), indexing="ij")
        yz = np.stack([Y.ravel(), Z.ravel()], axis=1)
        xs = np.asarray(g.x_coords, dtype=np.float64)
        out = np.empty((i_hi - i_lo, ny, nz), dtype=bool)
        step = max(1, max_points // (ny * nz))
        for a in range(i_lo, i_hi, step):
            b = min(i_hi, a + step)
            pts = np.empty(((b - a) * ny * nz, 3), dtype=np.float64)
            pts[:, 0] = np.repeat(xs[a:b], ny * nz)
            pts[:, 1:] = np.tile(yz, (b - a, 1))
            out[a - i_lo:b - i_lo] = (np.asarray(sdf_object.sdf(pts)) <= 0).reshape(b - a, ny, nz)
        return out

    def _to_index(self, pos, what: str):
        """metres -> int(round(pos/dx)) when any entry is a float below max(shape) (solver.py:1818-1839)."""
        gs = self.global_shape
        if isinstance(pos, tuple) and len(pos) == 3 and any(isinstance(q, float) and q < max(gs) for q in pos):
            idx = tuple(int(round(q / self.dx)) for q in pos)
            for a, (i, n) in enumerate(zip(idx, gs)):
                if not 0 <= i < n:
                    raise ValueError(f"{what} {'xyz'[a]} position {pos[a]:.4f}m (index {i}) is outside grid (0-{n-1})")
            return idx
        if isinstance(pos, (tuple, list)) and len(pos) == 3:
            # integer indices address the array as NumPy would (solver.py:2437 reads p[i, j, k]): negative values
            # wrap once, anything else is an IndexError -- raised here instead of at the first recorded step
            out = []
            for a, (i, n) in enumerate(zip(pos, gs)):
                i = int(i)
                if not -n <= i < n:
                    raise IndexError(f"{what} index {i} is out of bounds for axis {a} with size {n}")
                out.append(i + n if i < 0 else i)
            return tuple(out)
        return pos

    def add_source(self, source) -> None:
        kind = getattr(source, "source_type", "point")
        if kind == "membrane":
            ext = self._grid.physical_extent()
            for a, (q, hi) in enumerate(zip(source.center, ext)):
                if not 0 <= q <= hi:
                    raise ValueError(f"Membrane center {'xyz'[a]}={q:.4f}m is outside grid (0-{hi:.4f}m)")
        elif kind == "point":
            source.position = self._to_index(source.position, "Source")
        self._sources.append(source)
        self._dirty.add("sources")

    def add_probe(self, name: str, position) -> None:
        if name in self._probes:
            raise ValueError(f"Probe '{name}' already exists")
        self._probes[name] = Probe(name=name, position=self._to_index(position, "Probe"))
        self._dirty.add("records")

    def add_microphone(self, position, name=None, pattern="omni", direction=None, up=None):
        if hasattr(position, "_initialize"):
            mic = position
            if not hasattr(mic, "_gather_tables"):            # a reference Microphone object: rebuild as ours
                mic = Microphone(position=mic.position, name=mic.name, pattern=mic.pattern,
                                 direction=tuple(mic._direction), up=tuple(mic._up))
        else:
            mic = Microphone(position=position, name=name, pattern=pattern, direction=direction, up=up)
        mic.name = mic.name or f"mic_{len(self._microphones)}"
        if mic.name in self._microphones:
            raise ValueError(f"Microphone '{mic.name}' already exists")
        mic._initialize(self)
        self._microphones[mic.name] = mic
        self._dirty.add("records")
        return mic

    def add_boundary(self, boundary) -> None:
        boundary.initialize(self)
        is_plane = plane_ops(boundary, self) is not None
        if not is_plane and sponge_tables(boundary, self) is None and type(boundary).__name__ != "RigidBoundary":
            raise NotImplementedError(f"{type(boundary).__name__} is not on the b200 device path "
                                      "(PML, ABCFirstOrder, RadiationImpedance, RigidBoundary are)")
        if not is_plane and any(plane_ops(b, self) is not None for b in self._boundaries) \
                and sponge_tables(boundary, self) is not None:
            # the sponge is fused into the step kernel and therefore always runs before the plane updates
            raise NotImplementedError("add PML boundaries before ABCFirstOrder / RadiationImpedance")
        self._boundaries.append(boundary)
        self._dirty.add("sponges")

    def enable_snapshots(self, interval: int, capture_velocity: bool = False) -> None:
        self._snapshot_interval = interval
        self._snapshot_velocity = capture_velocity

    # ------------------------------------------------------------------ materials (solver.py:2836-2942)
    def register_material(self, material, material_id: int | None = None) -> int:
        if material_id is None:
            material_id = next((q for q in range(1, 256) if q not in self._materials), None)
            if material_id is None:
                raise ValueError("Maximum of 255 materials reached")
        if not 1 <= material_id <= 255:
            raise ValueError("material_id must be in range 1-255 (0 is reserved for air)")
        if material_id in self._materials:
            raise ValueError(f"Material ID {material_id} is already registered")
        self._materials[material_id] = material
        self._dirty.add("ade")
        return material_id

    def set_material_region(self, mask, material_id: int) -> None:
        """``mask``: a bool array over the whole grid or, for grids too large to materialise on every rank,
        ``f(i_lo, i_hi) -> bool planes`` (as ``set_geometry``): a slab asks for its own planes and the live ghost
        planes next to them, which both sides of a cut need."""
        if material_id != 0 and material_id not in self._materials:
            raise ValueError(f"Material ID {material_id} not registered. Use register_material() first.")
        lo, hi = self._i0 - self._has_lower, self._i1 + self._has_upper
        if callable(mask):
            part = np.asarray(mask(lo, hi), dtype=bool)
            want = (hi - lo,) + self.global_shape[1:]
            if part.shape != want:
                raise ValueError(f"Mask shape {part.shape} doesn't match solver shape {want}")
        else:
            mask = np.asarray(mask)
            if mask.shape != self.global_shape:
                raise ValueError(f"Mask shape {mask.shape} doesn't match solver shape {self.global_shape}")
            part = mask[lo:hi]
        self._material_ext[part] = material_id
        self._dirty.add("ade")

    def set_material_box(self, material_id: int, x_range, y_range, z_range) -> None:
        if material_id != 0 and material_id not in self._materials:
            raise ValueError(f"Material ID {material_id} not registered. Use register_material() first.")
        lo, hi = self._i0 - self._has_lower, self._i1 + self._has_upper
        xs = slice(*x_range).indices(self.global_shape[0])       # NumPy slice semantics of the reference (solver.py:2917)
        a, b = max(xs[0], lo) - lo, max(min(xs[1], hi) - lo, 0)
        if b > a:
            self._material_ext[a:b, y_range[0]:y_range[1], z_range[0]:z_range[1]] = material_id
        self._dirty.add("ade")

    def get_material_at(self, position):
        mat_id = self._material_id[position]
        return None if mat_id == 0 else self._materials.get(mat_id)

    # ------------------------------------------------------------------ device plumbing
    def set_kernel_option(self, option: int, value: int) -> None:
        """Tuning knobs of the step kernel (``_lib.OPT_*``); results never depend on them."""
        self._options[option] = int(value)
        if option in (_lib.OPT_ADE_LAYOUT, _lib.OPT_KERNEL):
            self._dirty.add("ade")                   # the layout is chosen when the material tables are built
        if self._dev is not None:
            _lib.check(self._dev.lib.sb_set_option(self._dev.handle, option, int(value)))

    def _ensure_device(self) -> _DeviceState:
        if self._dev is None:
            self._dev = _DeviceState(self.shape, self._device_index,
                                     slab=(self._has_lower, self._has_upper, self.global_shape[0], self._i0),
                                     p_allocator=getattr(self, "_p_allocator", None))
            for opt, val in self._options.items():
                _lib.check(self._dev.lib.sb_set_option(self._dev.handle, opt, val))
        return self._dev

    def _download(self, name: str) -> None:
        dev = self._ensure_device()
        _lib.check(dev.lib.sb_download_field(dev.handle, _FIELDS.index(name), _lib.ptr(self._host[name])))
        self._host_stale.discard(name)

    def _coefficient_tables(self):
        """fp32 per-face velocity coefficients and per-cell inverse spacings (see strata_b200.h)."""
        nx, ny, nz = self.global_shape
        if self._grid.is_uniform:
            cv = np.float32(self._coeff_v)
            faces = [np.full(n, cv, dtype=np.float32) for n in (nx, ny, nz)]
            cells, cp = [None] * 3, np.float32(self._coeff_p)
        else:
            sa = self._spacing_arrays
            cvb = np.float32(self._coeff_v_base)
            faces = []
            for a, n in zip("xyz", (nx, ny, nz)):
                t = np.zeros(n, dtype=np.float32)
                t[: n - 1] = cvb * sa[f"inv_d{a}_face"]      # one rounded fp32 multiply (fdtd_step.cpp:263)
                faces.append(t)
            cells = [np.ascontiguousarray(sa[f"inv_d{a}_cell"], dtype=np.float32) for a in "xyz"]
            cp = np.float32(self._coeff_p_base)
        faces[0] = self._x_slice(faces[0])
        if cells[0] is not None:
            cells[0] = self._x_slice(cells[0])
        return faces, cells, cp

    def _x_slice(self, table):
        """Axis-0 table restricted to this slab's planes, live lower ghost first (see strata_b200.h)."""
        return np.ascontiguousarray(table[self._i0 - self._has_lower: self._i1])

    def _build_source_table(self):
        """CSR over cells: (cell, [source id, field, weight]...) in source order (solver.py:2386-2433)."""
        nx, ny, nz = self.shape
        cells, sids, flds, wts = [], [], [], []
        g = self._geometry if self._geometry is not None else np.broadcast_to(np.True_, self.shape)
        for sid, src in enumerate(self._sources):
            kind = getattr(src, "source_type", "point")
            if kind == "point":
                i, j, k = src.position
                i -= self._i0                                  # local plane; other slabs own the rest
                if 0 <= i < nx and g[i, j, k]:
                    cells.append(np.array([(i * ny + j) * nz + k], dtype=np.int64))
                    wts.append(np.ones(1)); flds.append(np.zeros(1, dtype=np.int32))
                    sids.append(np.full(1, sid, dtype=np.int32))
            elif kind == "plane":
                sel = [slice(None)] * 3
                axis, index = src.position["axis"], src.position["index"]
                if axis == 0:
                    index -= self._i0
                    if not 0 <= index < nx:
                        continue
                sel[axis] = index
                m = np.zeros(self.shape, dtype=bool)
                m[tuple(sel)] = g[tuple(sel)]
                idx = np.flatnonzero(m)
                cells.append(idx); wts.append(np.ones(idx.size)); flds.append(np.zeros(idx.size, dtype=np.int32))
                sids.append(np.full(idx.size, sid, dtype=np.int32))
            elif kind == "membrane":
                if getattr(src, "_cached_weights", None) is None:
                    src._check_grid_alignment(self._grid)
                    src._cached_weights = src.get_injection_weights(self._grid)
                    src._cached_mask = src._cached_weights > 0
                fld = 0 if src.injection_type == "pressure" else 1 + "xyz".index(src.normal_axis)
                # an x-normal velocity membrane on the lower neighbour's last plane: that face is kept redundantly here as
                # the ghost vx[-1], so it receives the same injection (cells of local plane -1 have negative indices)
                ghost = 1 if (fld == 1 and self._has_lower) else 0
                sl = slice(self._i0 - ghost, self._i1)
                gg = True if self._geometry is None else self._geometry_ext[self._has_lower - ghost: self._has_lower + nx]
                m = src._cached_mask[sl] & gg
                idx = np.flatnonzero(m) - ghost * ny * nz
                cells.append(idx); wts.append(np.asarray(src._cached_weights, dtype=np.float64)[sl][m])
                flds.append(np.full(idx.size, fld, dtype=np.int32)); sids.append(np.full(idx.size, sid, dtype=np.int32))
            else:
                raise ValueError(f"unknown source_type {kind!r}")
        if not cells:
            z = np.zeros(0, dtype=np.int64)
            return z, np.zeros(1, dtype=np.int32), z.astype(np.int32), z.astype(np.int32), z.astype(np.float64)
        cells, sids, flds, wts = (np.concatenate(q) for q in (cells, sids, flds, wts))
        order = np.argsort(cells, kind="stable")             # keeps source order inside a cell
        cells, sids, flds, wts = cells[order], sids[order], flds[order], wts[order]
        uniq, first = np.unique(cells, return_index=True)
        start = np.append(first, cells.size).astype(np.int32)
        return (uniq.astype(np.int64), start, np.ascontiguousarray(sids, dtype=np.int32),
                np.ascontiguousarray(flds, dtype=np.int32), np.ascontiguousarray(wts, dtype=np.float64))

    def _pole_table(self):
        """Debye poles of all materials first, then Lorentz, each in registration order (solver.py:3024-3054)."""
        deb, lor = [], []
        n_ids = max(self._materials) + 1
        rho_inf = np.zeros(n_ids, dtype=np.float32)
        k_inf = np.zeros(n_ids, dtype=np.float32)
        for mat_id, mat in self._materials.items():
            rho_inf[mat_id] = float(mat.rho_inf)
            k_inf[mat_id] = float(mat.K_inf)
            for pole in mat.poles:
                co = pole.fdtd_coefficients(self.dt)
                tgt = 0 if pole.target == "density" else 1
                if pole.is_debye:
                    deb.append(_lib.Pole(mat_id, 0, tgt, 0, float(co[0]), float(co[1]), 0.0, 0.0))
                else:
                    lor.append(_lib.Pole(mat_id, 1, tgt, 0, float(co[0]), float(co[1]), float(co[2]), 0.0))
        poles = deb + lor
        return (_lib.Pole * max(1, len(poles)))(*poles), len(poles), rho_inf, k_inf

    def _sync_to_device(self) -> _DeviceState:
        dev = self._ensure_device()
        lib, h = dev.lib, dev.handle
        nx, ny, nz = self.shape
        if "coeffs" in self._dirty:
            faces, cells, cp = self._coefficient_tables()
            _lib.check(lib.sb_set_coefficients(h, *(_lib.ptr(t) for t in faces), *(_lib.ptr(t) for t in cells),
                                               float(cp)))
        if "geometry" in self._dirty:
            if self._geometry is None:
                _lib.check(lib.sb_set_geometry(h, None, 0))
            else:
                g = np.ascontiguousarray(self._geometry_ext, dtype=np.uint8)   # owned planes + live ghosts
                _lib.check(lib.sb_set_geometry(h, _lib.ptr(g), int(self._rigid)))
        if "sponges" in self._dirty:
            _lib.check(lib.sb_clear_sponges(h))
            _lib.check(lib.sb_clear_plane_ops(h))
            for b in self._boundaries:
                for op in plane_ops(b, self) or []:
                    axis, side = op[0], op[1]
                    if axis == 0 and ((side == 0 and self._has_lower) or (side == 1 and self._has_upper)):
                        continue                             # that face of the grid belongs to another slab
                    _lib.check(lib.sb_add_plane_op(h, *op))
            for b in self._boundaries:                       # application order = list order (solver.py:2044-2047)
                tabs = sponge_tables(b, self)
                if tabs is not None:
                    tabs = (None if tabs[0] is None else self._x_slice(tabs[0]), tabs[1], tabs[2])
                    _lib.check(lib.sb_add_sponge(h, *(_lib.ptr(t) for t in tabs)))
        if "sources" in self._dirty:
            cells, start, sids, flds, wts = self._build_source_table()
            _lib.check(lib.sb_set_sources(h, len(self._sources), len(cells), _lib.ptr(cells), _lib.ptr(start),
                                          _lib.ptr(sids), _lib.ptr(flds), _lib.ptr(wts)))
        if "records" in self._dirty:
            self._local_probes = [pr for pr in self._probes.values() if self._i0 <= pr.position[0] < self._i1]
            flat = np.array([((i - self._i0) * ny + j) * nz + k
                             for (i, j, k) in (pr.position for pr in self._local_probes)], dtype=np.int64)
            _lib.check(lib.sb_set_probes(h, len(flat), _lib.ptr(flat) if len(flat) else None))
            mics = list(self._microphones.values())
            self._mic_slots = []                              # per microphone: its record slots after the probes
            self._corner_keys = []                            # slab only: (mic, gather, corner) of each extra record slot
            self._mic_gathers = []
            if mics and (self._has_lower or self._has_upper):
                self._set_corner_samples(mics, lib, h)
            elif mics and any(m.is_directional() for m in mics):
                # one directional microphone switches the reference to its Python path for ALL microphones
                # (solver.py:2453-2461): p (+ vx, vy, vz) gathers with that path's weights, combined on the host
                fields, idx_all, w_all = [], [], []
                for m in mics:
                    tabs = m._gather_tables(self.shape)
                    self._mic_slots.append(list(range(len(fields), len(fields) + len(tabs))))
                    for f, idx8, w8 in tabs:
                        fields.append(f); idx_all.append(idx8); w_all.append(w8)
                fields = np.array(fields, dtype=np.int32)
                idx_all, w_all = np.concatenate(idx_all), np.concatenate(w_all)
                _lib.check(lib.sb_set_gathers(h, len(fields), _lib.ptr(fields), _lib.ptr(idx_all), _lib.ptr(w_all)))
            elif mics:
                # Omnidirectional microphones are recorded as their eight raw corner samples of p -- plain probes --
                # and summed on the host in the order of microphones.cpp:82-116 (bit-identical, see
                # combine_corner_samples).  With nothing but probes on the device, the chunk kernels K5 / K6 apply.
                self._mic_gathers = self.microphone_gathers(mics, lib)
                n_cells = int(np.prod(self.shape, dtype=np.int64))
                corner_idx = []
                for mi, tabs in enumerate(self._mic_gathers):
                    for c, gidx in enumerate(tabs[0][1]):
                        if not 0 <= int(gidx) < n_cells:
                            raise _lib.B200BackendError(f"microphone corner {8 * mi + c} out of range")
                        corner_idx.append(int(gidx))
                        self._corner_keys.append((mi, 0, c))
                self._mic_tables = (np.concatenate([t[0][1] for t in self._mic_gathers]),
                                    np.concatenate([t[0][2] for t in self._mic_gathers]))
                flat_all = np.concatenate([flat, np.array(corner_idx, dtype=np.int64)])
                _lib.check(lib.sb_set_probes(h, len(flat_all), _lib.ptr(flat_all)))
                _lib.check(lib.sb_set_mics(h, 0, None, None))
            else:
                _lib.check(lib.sb_set_mics(h, 0, None, None))
        if "ade" in self._dirty:
            if self._materials and any(len(m.poles) for m in self._materials.values()):
                poles, n_poles, rho_inf, k_inf = self._pole_table()
                mid = np.ascontiguousarray(self._material_ext, dtype=np.uint8)   # owned planes + live ghosts
                _lib.check(lib.sb_set_ade(h, poles, n_poles, _lib.ptr(mid), _lib.ptr(rho_inf), _lib.ptr(k_inf),
                                          len(rho_inf), float(self.dt), float(1.0 / self.dx)))
            else:
                _lib.check(lib.sb_set_ade(h, None, 0, None, None, None, 0, 0.0, 0.0))
        self._dirty.clear()
        for name in list(self._host_dirty):
            if name not in self._host_stale:
                _lib.check(lib.sb_upload_field(h, _FIELDS.index(name), _lib.ptr(self._host[name])))
        self._host_dirty.clear()
        return dev

    def _native_mic_tables(self, mics, lib):
        """Corner indices (global flat) and fp32 weights of the native path (microphones.cpp:16-80)."""
        ny, nz = self.shape[1], self.shape[2]
        gp = np.array([q for m in mics for q in m._grid_position], dtype=np.float32)   # fp32 store, solver.py:2502-2509
        idx8 = np.zeros(8 * len(mics), dtype=np.int64)
        w8 = np.zeros(8 * len(mics), dtype=np.float32)
        _lib.check(lib.sb_mic_tables(_lib.ptr(gp), len(mics), ny, nz, _lib.ptr(idx8), _lib.ptr(w8)))
        return idx8, w8

    def microphone_gathers(self, mics, lib=None):
        """Per microphone the list of (field, idx8 global flat, w8 fp32) the reference would gather (native tables when
        every microphone is omnidirectional, the Python path's otherwise -- solver.py:2453-2461)."""
        lib = lib or _lib.load()
        if any(m.is_directional() for m in mics):
            return [m._gather_tables(self.global_shape) for m in mics]
        idx8, w8 = self._native_mic_tables(mics, lib)
        return [[(0, idx8[8 * q: 8 * q + 8], w8[8 * q: 8 * q + 8])] for q in range(len(mics))]

    def _set_corner_samples(self, mics, lib, h) -> None:
        """Microphones on a slab: the eight corners of a trilinear gather can lie on two slabs, and the fp32 sum runs
        over them in corner order.  Every slab therefore records the raw corner values it owns (one-hot gathers) and
        the driver adds them up in the reference's order once all corners are known (multi.combine_corner_samples)."""
        ny, nz = self.shape[1], self.shape[2]
        n_cells = int(np.prod(self.global_shape, dtype=np.int64))
        self._mic_gathers = self.microphone_gathers(mics, lib)
        fields, idx_all = [], []
        for mi, tabs in enumerate(self._mic_gathers):
            for g, (f, idx8, _w) in enumerate(tabs):
                for c, gidx in enumerate(idx8):
                    gidx = int(gidx)
                    if not 0 <= gidx < n_cells:
                        raise _lib.B200BackendError(f"microphone corner {c} of '{mics[mi].name}' out of range")
                    i = gidx // (ny * nz)
                    if self._i0 <= i < self._i1:
                        fields.append(f)
                        idx_all.append(gidx - self._i0 * ny * nz)
                        self._corner_keys.append((mi, g, c))
        n = len(fields)
        if n == 0:
            _lib.check(lib.sb_set_mics(h, 0, None, None))
            return
        one_hot = np.zeros((n, 8), dtype=np.float32)
        one_hot[:, 0] = 1.0
        idx8 = np.repeat(np.array(idx_all, dtype=np.int64), 8)
        fields = np.array(fields, dtype=np.int32)
        _lib.check(lib.sb_set_gathers(h, n, _lib.ptr(fields), _lib.ptr(idx8), _lib.ptr(np.ascontiguousarray(one_hot))))

    # ------------------------------------------------------------------ time stepping
    def _waveform_table(self, times: np.ndarray) -> np.ndarray:
        """W[n, s] = sample of source s at step n, float64, evaluated exactly like solver.py:2391/2416.

        The vectorised evaluation is spot-checked against the reference's one-element-array form (five samples of a
        source's first table, one of every later table); any difference (a SIMD/scalar libm split) switches that source
        to per-step evaluation for good."""
        n = len(times)
        W = np.zeros((n, max(1, len(self._sources))), dtype=np.float64)
        state = self.__dict__.setdefault("_waveform_vector_ok", {})
        for s, src in enumerate(self._sources):
            fn = src.waveform.waveform if getattr(src, "source_type", "") == "membrane" else src.waveform
            ok = state.get(id(src))
            col = None
            if ok is not False:
                col = np.asarray(fn(times.copy(), self.dt), dtype=np.float64)
                probe_at = sorted({0, n - 1, n // 2, n // 3, (2 * n) // 3}) if ok is None else [(7 * n) // 11]
                ok = col.shape == (n,) and all(
                    np.float64(fn(np.array([times[q]]), self.dt)[0]).tobytes() == col[q].tobytes() for q in probe_at)
                state[id(src)] = ok
            if not ok:
                col = np.array([fn(np.array([t]), self.dt)[0] for t in times], dtype=np.float64)
            W[:, s] = col
        return W

    def _io_slot(self, m: int, n_src: int, n_rec: int):
        """Two page-locked host buffers for the waveform table and the record block of a chunk, so that the host can
        prepare chunk n+1 and unpack chunk n-1 while the device runs chunk n (sb_step_n_submit / sb_step_n_wait)."""
        dev = self._dev
        torch = dev.torch
        io = getattr(self, "_io", None)
        if io is None or io["cap"] < m or io["n_src"] != n_src or io["n_rec"] != n_rec or io["dev"] is not dev:
            cap = max(m, io["cap"] if io and io["dev"] is dev else 0)
            keep = [torch.empty((cap, n_src), dtype=torch.float64).pin_memory() for _ in range(2)] + \
                   [torch.empty((cap, n_rec), dtype=torch.float32).pin_memory() for _ in range(2)]
            io = dict(cap=cap, n_src=n_src, n_rec=n_rec, dev=dev, turn=0, keep=keep,
                      W=[t.numpy() for t in keep[:2]], rec=[t.numpy() for t in keep[2:]])
            self._io = io
        io["turn"] ^= 1
        return io, io["turn"]

    def _launch_chunk(self, m: int) -> dict:
        """Enqueue m steps: waveform table up, kernels, records down (all asynchronous on the solver's stream)."""
        dev = self._dev
        n_src = max(1, len(self._sources))
        n_rec = len(self._local_probes) + sum(len(sl) for sl in self._mic_slots) + len(self._corner_keys)
        # the same float64 accumulation as solver.py:2072 (np.add.accumulate adds strictly left to right)
        steps_t = np.full(m + 1, self.dt, dtype=np.float64)
        steps_t[0] = self._time
        acc = np.add.accumulate(steps_t)
        times, t_end = acc[:m], float(acc[m])
        io, k = self._io_slot(m, n_src, max(1, n_rec))
        W = io["W"][k][:m]
        if self._sources:
            W[...] = self._waveform_table(times)
        _lib.check(dev.lib.sb_step_n_submit(dev.handle, k, m, _lib.ptr(W), _lib.ptr(io["rec"][k]) if n_rec else None))
        self._host_stale = set(_FIELDS)
        first_idx = self._step_count
        self._step_count += m
        self._time = t_end
        return dict(m=m, times=times, rec=io["rec"][k], n_rec=n_rec, slot=k, first_idx=first_idx)

    def _finish_chunk(self, tk: dict, writer=None) -> None:
        """Wait for a launched chunk and file its samples: probes, microphones, corner samples, result writer."""
        _lib.check(self._dev.lib.sb_step_n_wait(self._dev.handle, tk["slot"]))
        m, times = tk["m"], tk["times"]
        probes = self._local_probes
        mics = list(self._microphones.values())
        rec = tk["rec"][:m] if tk["n_rec"] else np.empty((m, 0), dtype=np.float32)
        for q, pr in enumerate(probes):
            pr.data.extend(rec[:, q].tolist())
        for mic, slots in zip(mics, self._mic_slots):
            cols = [np.array(rec[:, len(probes) + q]) for q in slots]
            mic._data.extend(mic._combine(cols[0], cols[1:]).tolist())
            mic._times.extend(times.tolist())
        self._store_corner_samples(rec, len(probes), times)
        if self._corner_keys and not (self._has_lower or self._has_upper):     # one GPU: every corner is here
            data = {k: np.concatenate(v) for k, v in self._corner_data.items()}
            self._corner_data.clear(); self._corner_times.clear()
            combine_corner_samples(mics, self._mic_gathers, data, times)
        if writer is not None and probes:
            writer.append_probe_block([pr.name for pr in probes], rec[:, :len(probes)])

    def _advance(self, n_steps: int, callback=None, writer=None, snapshot_interval=None) -> None:
        dev = self._sync_to_device()
        done = 0
        # On small grids a step takes microseconds and the per-chunk host work (waveform table, copies, trace lists)
        # shows; nobody is watching the steps go by unless a callback / writer is attached, so use longer chunks there.
        chunk = self._chunk_steps
        if self._chunk_auto and callback is None and writer is None and int(np.prod(self.shape, dtype=np.int64)) <= (2 << 20):
            chunk = max(chunk, 1024)
        pending = None                     # the chunk the device is working on while the host prepares the next one
        try:
            while done < n_steps:
                m = min(chunk, n_steps - done)
                # a chunk ends right after any step whose fields the host has to see
                host_looks = False
                if self._snapshot_interval or self._track_energy or (writer is not None and snapshot_interval is not None):
                    for q in range(m):
                        idx = self._step_count + q
                        need = (self._snapshot_interval and idx % self._snapshot_interval == 0) or \
                               (self._track_energy and (idx + 1) % self._energy_sample_interval == 0) or \
                               (writer is not None and snapshot_interval is not None and (done + q) % snapshot_interval == 0)
                        if need:
                            m, host_looks = q + 1, True
                            break
                tk = self._launch_chunk(m)
                if pending is not None:
                    self._finish_chunk(pending, writer)
                    pending = None
                if host_looks or callback is not None:
                    # the host reads fields / reports progress at this point: complete the chunk before going on
                    self._finish_chunk(tk, writer)
                    last_idx, t_last = tk["first_idx"] + m - 1, float(tk["times"][-1])
                    if self._snapshot_interval and last_idx % self._snapshot_interval == 0:
                        self._snapshots.append((t_last, self.get_field("p")))
                        if self._snapshot_velocity:
                            self._velocity_snapshots.append((t_last, *self._centred_velocities()))
                    if self._track_energy and self._step_count % self._energy_sample_interval == 0:
                        self._energy_history.append((self._step_count, self._time, self.compute_energy()))
                    if writer is not None and snapshot_interval is not None and (done + m - 1) % snapshot_interval == 0:
                        writer.write_snapshot(self.get_field("p"))
                    if callback is not None:
                        for q in range(m):
                            callback(last_idx - (m - 1 - q))
                else:
                    pending = tk
                done += m
        finally:
            if pending is not None:
                self._finish_chunk(pending, writer)
            _lib.check(dev.lib.sb_synchronize(dev.handle))       # also surfaces a timed-out wait inside a chunk kernel

    def _store_corner_samples(self, rec, first_slot: int, times) -> None:
        """Slab only: keep the raw microphone corner values of a chunk until the driver combines them."""
        if not self._corner_keys:
            return
        for q, key in enumerate(self._corner_keys):
            self._corner_data.setdefault(key, []).append(np.array(rec[:, first_slot + q], dtype=np.float32))
        self._corner_times.append(np.array(times, dtype=np.float64))

    def step(self) -> None:
        """Advance one time step (solver.py:2003-2077)."""
        self._advance(1)

    def run(self, duration: float | None = None, progress: bool = False, track_energy: bool = False,
            energy_sample_interval: int = 1, output_file: str | None = None, script_content: str | None = None,
            callback: Callable[[int], None] | None = None, snapshot_interval: int | None = None,
            steps: int | None = None, output: str | None = None) -> None:
        """Run for ``duration`` seconds (or ``steps`` steps) -- solver.py:2520-2606."""
        if steps is None:
            if duration is None:
                raise ValueError("run() needs duration= or steps=")
            steps = int(np.ceil(duration / self.dt))
        output_file = output_file or output
        t0 = _time_mod.time()
        self._track_energy = track_energy
        self._energy_sample_interval = energy_sample_interval
        if track_energy and not self._energy_history:
            self._energy_history.append((self._step_count, self._time, self.compute_energy()))
        writer = None
        if output_file:
            from .io import ResultWriter
            writer = ResultWriter(output_file, self, script_content)
        bar = None
        cb = callback
        if progress:
            try:
                from tqdm import tqdm
                bar = tqdm(total=steps, desc="FDTD simulation (b200)")
                cb = (lambda s: (bar.update(1), callback(s) if callback else None))
            except ImportError:
                bar = None
        launches0 = self.kernel_launches()
        try:
            self._advance(steps, callback=cb, writer=writer, snapshot_interval=snapshot_interval)
            if self._warn_energy_drift and len(self._energy_history) >= 2:
                rep = self.energy_report()
                if abs(rep["energy_change_percent"]) > self._energy_drift_threshold * 100:
                    warnings.warn(f"Energy drift detected: {rep['energy_change_percent']:.2f}% change "
                                  f"(threshold: {self._energy_drift_threshold * 100:.1f}%). "
                                  f"Status: {rep['conservation_status']}", UserWarning, stacklevel=2)
        finally:
            runtime = _time_mod.time() - t0
            if bar is not None:
                bar.close()
            cells = int(np.prod(self.shape, dtype=np.int64))
            self.last_run_stats = {"steps": steps, "runtime_s": runtime,
                                   "cell_updates_per_s": cells * steps / runtime if runtime > 0 else float("inf"),
                                   "kernel_launches": self.kernel_launches() - launches0}
            if writer is not None:
                writer.finalize(runtime=runtime, backend="b200", num_threads=0)

    # ------------------------------------------------------------------ results / diagnostics
    def get_probe_data(self, name: str | None = None) -> dict:
        if name is not None:
            if name not in self._probes:
                raise KeyError(f"Probe '{name}' not found")
            return {name: self._probes[name].get_data()}
        return {n: pr.get_data() for n, pr in self._probes.items()}

    def get_snapshots(self):
        return self._snapshots

    def get_velocity_snapshots(self):
        return self._velocity_snapshots

    def _centred_velocities(self):
        return centre_velocities(*(self.get_field(name) for name in ("vx", "vy", "vz")))

    def _compute_divergence(self) -> np.ndarray:
        """Velocity divergence of the current state, evaluated on the host from the downloaded fields -- a diagnostic with the
        arithmetic of the reference's helper of the same name (core/solver.py:3214-3270: backward differences with a zero
        face below index 0; 1/dx, or the per-cell inverse spacings on a nonuniform grid).  The step kernels do not use it."""
        v = [self.get_field(f) for f in ("vx", "vy", "vz")]
        parts = []
        for axis in range(3):
            d = v[axis].copy()
            hi = [slice(None)] * 3; lo = [slice(None)] * 3
            hi[axis], lo[axis] = slice(1, None), slice(0, -1)
            d[tuple(hi)] -= v[axis][tuple(lo)]
            parts.append(d)
        if getattr(self._grid, "is_uniform", self._spacing_arrays is None):
            return ((parts[0] + parts[1]) + parts[2]) / self.dx
        sa = self._spacing_arrays
        inv = [np.asarray(sa[f"inv_d{a}_cell"], dtype=np.float32) for a in "xyz"]
        return (parts[0] * inv[0][:, None, None] + parts[1] * inv[1][None, :, None]) + parts[2] * inv[2][None, None, :]

    def compute_energy(self) -> float:
        """(1/2) sum(p^2/(rho c^2) + rho |v|^2) dV over air cells (solver.py:2689-2706), reduced on the device."""
        dev = self._sync_to_device()
        out = C.c_double(0.0)
        _lib.check(dev.lib.sb_energy(dev.handle, float(self.rho), float(self.c), float(self.dx**3), C.byref(out)))
        return out.value

    def get_energy_history(self):
        return self._energy_history.copy()

    def energy_report(self) -> dict:
        if not self._energy_history:
            raise ValueError("No energy history recorded. Call run() with track_energy=True first.")
        e = np.array([q[2] for q in self._energy_history])
        first, last = e[0], e[-1]
        if first == 0:
            change = 0.0 if last == 0 else float("inf")
        else:
            change = (last - first) / first * 100
        status = "stable" if abs(change) <= 1.0 else ("growing" if change > 0 else "decaying")
        return {"initial_energy": float(first), "final_energy": float(last), "max_energy": float(e.max()),
                "min_energy": float(e.min()), "energy_change_percent": change, "conservation_status": status,
                "n_samples": len(e)}

    def get_sample_rate(self) -> float:
        return 1.0 / self.dt

    def get_frequency_response(self, probe_name: str, n_fft: int | None = None):
        data = self._probes[probe_name].get_data()
        if n_fft is None:
            n_fft = int(2 ** np.ceil(np.log2(len(data))))
        return np.fft.rfftfreq(n_fft, self.dt), np.abs(np.fft.rfft(data, n=n_fft))

    def reset(self) -> None:
        """Back to t = 0 with zero fields (solver.py:2781-2800)."""
        for f in _FIELDS:
            self._host[f].fill(0)
        self._host_stale.clear()
        self._host_dirty.clear()
        if self._dev is not None:
            _lib.check(self._dev.lib.sb_reset(self._dev.handle))
        self._step_count = 0
        self._time = 0.0
        self._snapshots.clear()
        self._velocity_snapshots.clear()
        self._energy_history.clear()
        self._track_energy = False
        for pr in self._probes.values():
            pr.clear()
        for mic in self._microphones.values():
            mic.clear()
        self._corner_data.clear()
        self._corner_times.clear()
        for b in self._boundaries:
            b.reset()

    # ------------------------------------------------------------------ checkpoint / resume
    def get_state(self) -> dict:
        """Everything a run needs to continue elsewhere: the four fields, time, step count and -- with dispersive
        materials -- the auxiliary fields J (and J_prev of Lorentz poles) as dense arrays in pole-table order
        (Debye poles of all materials first, then Lorentz; the reference keeps them private, solver.py:3061-3083), and
        the previous-plane arrays of Mur / radiation boundaries in the order their faces were registered."""
        st = {"time": float(self._time), "step_count": int(self._step_count)}
        for f in _FIELDS:
            st[f] = self.get_field(f)
        if self._materials and any(len(m.poles) for m in self._materials.values()):
            dev = self._sync_to_device()
            poles, n_poles, _r, _k = self._pole_table()
            ade = []
            for q in range(n_poles):
                ent = {"material_id": int(poles[q].material_id), "is_lorentz": bool(poles[q].is_lorentz),
                       "target": "density" if poles[q].target == 0 else "modulus"}
                for which, key in ((0, "J"), (1, "J_prev")):
                    if which == 1 and not poles[q].is_lorentz:
                        continue
                    a = np.zeros(self.shape, dtype=np.float32)
                    _lib.check(dev.lib.sb_ade_state(dev.handle, q, which, _lib.ptr(a), 0))
                    ent[key] = a
                ade.append(ent)
            st["ade"] = ade
        if any(plane_ops(b, self) is not None for b in self._boundaries):
            dev = self._sync_to_device()
            planes, q = [], 0
            while True:                                             # the previous-plane state of every Mur / radiation face
                n = C.c_int64(0)
                if dev.lib.sb_plane_op_state(dev.handle, q, None, C.byref(n), 0) != 0:
                    break
                a = np.zeros(n.value, dtype=np.float32)
                _lib.check(dev.lib.sb_plane_op_state(dev.handle, q, _lib.ptr(a), None, 0))
                planes.append(a)
                q += 1
            st["planes"] = planes
        return st

    def set_state(self, state: dict) -> None:
        """Inverse of ``get_state`` on a solver with the same set-up (grid, materials in the same order)."""
        for f in _FIELDS:
            self._field_set(f, np.asarray(state[f], dtype=np.float32))
        self._time, self._step_count = float(state["time"]), int(state["step_count"])
        dev = self._sync_to_device()
        for q, ent in enumerate(state.get("ade", [])):
            for which, key in ((0, "J"), (1, "J_prev")):
                if key in ent:
                    a = np.ascontiguousarray(ent[key], dtype=np.float32)
                    if a.shape != self.shape:
                        raise ValueError(f"auxiliary field shape {a.shape} doesn't match solver shape {self.shape}")
                    _lib.check(dev.lib.sb_ade_state(dev.handle, q, which, _lib.ptr(a), 1))
        for q, a in enumerate(state.get("planes", [])):
            a = np.ascontiguousarray(a, dtype=np.float32)
            n = C.c_int64(0)
            _lib.check(dev.lib.sb_plane_op_state(dev.handle, q, None, C.byref(n), 0))
            if a.size != n.value:
                raise ValueError(f"plane state {q} has {a.size} entries, the solver's face has {n.value}")
            _lib.check(dev.lib.sb_plane_op_state(dev.handle, q, _lib.ptr(a), None, 1))

    def kernel_launches(self) -> int:
        if self._dev is None:
            return 0
        st = _lib.Stats()
        _lib.check(self._dev.lib.sb_query(self._dev.handle, C.byref(st)))
        return int(st.kernels_launched)

    def device_stats(self) -> dict:
        dev = self._ensure_device()
        st = _lib.Stats()
        _lib.check(dev.lib.sb_query(dev.handle, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def device_field(self, name: str):
        """Zero-copy torch view [nx, ny, nz] of a field on the device (valid until the next step)."""
        dev = self._sync_to_device()
        return dev.field_view(name)

    def close(self) -> None:
        if self._dev is not None:
            self._dev.close()
            self._dev = None
