"""Asynchronous result writer for the b200 backend.

Writes the reference's result schema (/root/reference/src/strata_fdtd/io/hdf5.py:42-233):
``/metadata``, ``/grid``, ``/simulation``, ``/sources/source_i`` attribute groups, resizable
fp32 ``/probes/<name>`` datasets, ``/fields/pressure[t, nx, ny, nz]`` snapshots (gzip-4,
one chunk per snapshot) and ``/materials/geometry``.  Unlike the reference, which rebuilds
every probe array on every step (hdf5.py:164-181, O(steps^2)), probe samples arrive here in
whole chunks straight from the device record buffer and are appended by a background
thread while the GPU keeps stepping.

h5py is optional in this image: without it the same tree is stored as a flat ``.npz``
(keys are the HDF5 paths, attributes under ``<group>@<attr>``) next to the requested name.
"""
from __future__ import annotations

import hashlib
import json
import queue
import threading
from datetime import datetime, timezone
from pathlib import Path

import numpy as np

try:
    import h5py
    HAVE_H5PY = hasattr(h5py, "File")     # a stub module may be registered under this name
except ImportError:      # pragma: no cover - depends on the image
    h5py = None
    HAVE_H5PY = False


class _NpzTree:
    """Minimal stand-in with the subset of the h5py interface the writer uses."""

    def __init__(self, path: Path):
        self.path = path
        self.data: dict[str, np.ndarray] = {}
        self.attrs: dict[str, object] = {}

    def set_attr(self, group: str, key: str, value):
        self.attrs[f"{group}@{key}"] = value

    def append(self, name: str, block: np.ndarray):
        self.data.setdefault(name, []).append(block.copy())      # joined once, at close: linear in the run length

    def close(self):
        payload = {k: (np.concatenate(v) if isinstance(v, list) else v) for k, v in self.data.items()}
        payload["__attrs__"] = np.array(json.dumps(self.attrs, default=str))
        with open(self.path, "wb") as fh:            # a file object keeps the exact name (no ".npz" appended)
            np.savez_compressed(fh, **payload)


class ResultWriter:
    def __init__(self, filename, solver, script_content: str | None = None,
                 compression: str = "gzip", compression_level: int = 4):
        self.solver = solver
        self.filename = Path(filename)
        self.use_h5 = HAVE_H5PY
        self.compression = compression
        self.compression_opts = compression_level if compression == "gzip" else None
        self._n_snap = 0
        self._pressure = None
        if self.use_h5:
            self.file = h5py.File(self.filename, "w")
        else:
            # same path the caller asked for (scripts stat / upload it afterwards); the content is an .npz archive
            self.file = _NpzTree(self.filename)
        self._metadata(script_content)
        self._q: queue.Queue = queue.Queue(maxsize=8)
        self._err: BaseException | None = None
        self._thread = threading.Thread(target=self._drain, name="strata-b200-writer", daemon=True)
        self._thread.start()

    # ---- metadata (written synchronously, once) ----------------------------------------
    def _attr(self, group: str, key: str, value):
        if self.use_h5:
            self.file.require_group(group).attrs[key] = value
        else:
            self.file.set_attr(group, key, value)

    def _metadata(self, script_content):
        s = self.solver
        if script_content:
            self._attr("metadata", "script_hash", hashlib.sha256(script_content.encode()).hexdigest())
            self._attr("metadata", "script_content", script_content)
        self._attr("metadata", "created_at", datetime.now(timezone.utc).isoformat())
        self._attr("metadata", "solver_version", "0.1.0")
        self._attr("grid", "shape", list(s.shape))
        uniform = bool(getattr(s.grid, "is_uniform", True))
        self._attr("grid", "is_uniform", uniform)
        if uniform:
            self._attr("grid", "resolution", s.dx)
        else:
            for a in "xyz":
                self._dataset(f"grid/{a}_coords", getattr(s.grid, f"{a}_coords"))
        self._attr("grid", "extent", [n * s.dx for n in s.shape])
        self._attr("simulation", "timestep", float(s.dt))
        self._attr("simulation", "cfl_number", float(s.c * s.dt / s.grid.min_spacing))
        self._attr("simulation", "c", s.c)
        self._attr("simulation", "rho", s.rho)
        if self.use_h5:
            self.file.require_group("sources"); self.file.require_group("fields"); self.file.require_group("probes")
        for i, src in enumerate(s._sources):
            g = f"sources/source_{i}"
            self._attr(g, "type", src.source_type)
            pos = getattr(src, "position", None)
            if isinstance(pos, dict):
                pos = [pos["axis"], pos["index"], -1]
            elif pos is None:
                pos = list(getattr(src, "center", (0, 0, 0)))
            self._attr(g, "position", list(pos))
            if hasattr(src, "frequency"):
                self._attr(g, "frequency", src.frequency)
            if hasattr(src, "bandwidth"):
                self._attr(g, "bandwidth", src.bandwidth)
        if self.use_h5:
            for name, probe in s._probes.items():
                d = self.file["probes"].create_dataset(name, shape=(0,), maxshape=(None,), dtype=np.float32,
                                                       chunks=True, compression=self.compression,
                                                       compression_opts=self.compression_opts)
                d.attrs["position"] = list(probe.position)
                d.attrs["units"] = "Pa"
        else:
            for name, probe in s._probes.items():
                self._attr(f"probes/{name}", "position", list(probe.position))
                self._attr(f"probes/{name}", "units", "Pa")
        self._dataset("materials/geometry", np.asarray(s.geometry).astype(np.uint8))

    def _dataset(self, name: str, data):
        if self.use_h5:
            self.file.create_dataset(name, data=data, compression=self.compression,
                                     compression_opts=self.compression_opts)
        else:
            self.file.data[name] = np.asarray(data)

    # ---- producer side (solver thread) ---------------------------------------------------
    def append_probe_block(self, names: list[str], block: np.ndarray) -> None:
        """block[n_steps, n_probes] of new samples, columns in ``names`` order."""
        self._put(("probes", names, np.array(block, dtype=np.float32, copy=True)))

    def write_snapshot(self, pressure: np.ndarray) -> None:
        self._put(("snapshot", np.array(pressure, dtype=np.float32, copy=True)))

    def _put(self, item):
        if self._err is not None:
            raise self._err
        self._q.put(item)

    # ---- consumer side (writer thread) ---------------------------------------------------
    def _drain(self):
        while True:
            item = self._q.get()
            try:
                if item is None:
                    return
                if self._err is None:
                    self._handle(item)
            except BaseException as e:          # surfaced on the next put / finalize
                self._err = e
            finally:
                self._q.task_done()

    def _handle(self, item):
        if item[0] == "probes":
            _, names, block = item
            for col, name in enumerate(names):
                if self.use_h5:
                    d = self.file["probes"][name]
                    n0 = d.shape[0]
                    d.resize((n0 + block.shape[0],))
                    d[n0:] = block[:, col]
                else:
                    self.file.append(f"probes/{name}", block[:, col])
        else:
            p = item[1]
            if self.use_h5:
                if self._pressure is None:
                    self._pressure = self.file["fields"].create_dataset(
                        "pressure", shape=(1,) + p.shape, maxshape=(None,) + p.shape, dtype=np.float32,
                        chunks=(1,) + p.shape, compression=self.compression, compression_opts=self.compression_opts)
                    self._pressure.attrs["units"] = "Pa"
                    self._pressure.attrs["snapshot_interval"] = 1
                if self._n_snap >= self._pressure.shape[0]:
                    self._pressure.resize((self._n_snap + 1,) + p.shape)
                self._pressure[self._n_snap] = p
            else:
                self.file.append("fields/pressure", p[None])
            self._n_snap += 1

    def finalize(self, runtime: float | None = None, **extra) -> None:
        self._q.put(None)
        self._thread.join()
        if self._err is not None:
            raise self._err
        self._attr("simulation", "num_steps", int(self.solver.step_count))
        self._attr("simulation", "total_time", float(self.solver.time))
        if runtime is not None:
            self._attr("metadata", "total_runtime_seconds", runtime)
        for k, v in extra.items():
            self._attr("metadata", k, v)
        if self.use_h5:
            self.file.flush()
        self.file.close()
