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


class HDF5ResultWriter(ResultWriter):
    """The reference writer's per-step interface (io/hdf5.py:24-240) on top of the chunked writer above, for scripts that
    drive the solver themselves: ``write_timestep(step, save_snapshot)`` after every ``solver.step()``, then ``finalize``.
    Only the samples recorded since the previous call are appended, so a long loop stays linear (the reference rebuilds each
    probe array on every step)."""

    def __init__(self, filename, solver, script_content: str | None = None, compression: str = "gzip", compression_level: int = 4):
        super().__init__(filename, solver, script_content, compression, compression_level)
        self._written = {name: 0 for name in solver._probes}
        self._open = True

    def write_timestep(self, step: int, save_snapshot: bool = False) -> None:
        for name, probe in self.solver._probes.items():
            data = probe.get_data()
            n0 = self._written.get(name, 0)
            if len(data) > n0:
                self.append_probe_block([name], np.asarray(data[n0:], dtype=np.float32)[:, None])
                self._written[name] = len(data)
        if save_snapshot:
            self.write_snapshot(np.asarray(self.solver.p))

    def finalize(self, runtime: float | None = None, **extra_metadata) -> None:
        if self._open:
            self._open = False
            super().finalize(runtime, **extra_metadata)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.finalize()


class HDF5ResultReader:
    """The reference reader's interface (io/hdf5.py:243-375) over either container this package writes: an HDF5 file (h5py
    present) or the flat ``.npz`` tree that stands in for it (h5py absent; HDF5 paths as keys, attributes under
    ``<group>@<attr>``) -- so analysis code written against the reference reads results from both."""

    def __init__(self, filename):
        self.filename = Path(filename)
        self._npz = None
        self.file = None
        with open(self.filename, "rb") as fh:
            magic = fh.read(8)
        if magic.startswith(b"PK"):                           # a zip archive: the .npz tree
            self._npz = np.load(self.filename, allow_pickle=False)
            self._attrs = json.loads(str(self._npz["__attrs__"]))
        elif HAVE_H5PY:
            self.file = h5py.File(self.filename, "r")
        else:
            raise OSError(f"{self.filename} is an HDF5 file and h5py is not installed")

    # ---- the two containers behind one small interface ---------------------------------------------------
    def _group_attrs(self, group: str) -> dict:
        if self.file is not None:
            return dict(self.file[group].attrs) if group in self.file else {}
        pre = group + "@"
        return {k[len(pre):]: v for k, v in self._attrs.items() if k.startswith(pre)}

    def _children(self, group: str) -> list[str]:
        if self.file is not None:
            return list(self.file[group].keys()) if group in self.file else []
        pre = group + "/"
        names = [k[len(pre):] for k in self._npz.files if k.startswith(pre)] + \
                [k[len(pre):].split("@")[0] for k in self._attrs if k.startswith(pre)]
        return sorted(set(n.split("/")[0] for n in names))        # (h5py lists group members in alphabetical order)

    def _has(self, path: str) -> bool:
        return path in self.file if self.file is not None else path in self._npz.files

    def _array(self, path: str):
        return self.file[path] if self.file is not None else self._npz[path]

    # ---- reference interface ---------------------------------------------------------------------------------
    def get_metadata(self) -> dict:
        out = {}
        for group in ("metadata", "grid", "simulation"):
            attrs = self._group_attrs(group)
            if attrs:
                out[group] = attrs
        if "grid" in out and not out["grid"].get("is_uniform", True):
            for a in "xyz":
                out["grid"][f"{a}_coords"] = np.asarray(self._array(f"grid/{a}_coords")[:])
        sources = self._children("sources")
        if sources or self._has("sources"):
            out["sources"] = [self._group_attrs(f"sources/{name}") for name in sources]
        probes = self.get_probe_names()
        if probes or self._has("probes"):
            out["probes"] = {name: self._group_attrs(f"probes/{name}") for name in probes}
        return out

    def load_timestep(self, step: int):
        if not self._has("fields/pressure"):
            raise ValueError("No pressure field data in file")
        return self._array("fields/pressure")[step]

    def load_probe(self, probe_name: str):
        if probe_name not in self.get_probe_names():
            raise KeyError(f"Probe '{probe_name}' not found. Available: {self.get_probe_names()}")
        path = f"probes/{probe_name}"
        return np.asarray(self._array(path)[:]) if self._has(path) else np.zeros(0, dtype=np.float32)

    def get_probe_names(self) -> list[str]:
        return self._children("probes")

    def get_num_snapshots(self) -> int:
        return int(self._array("fields/pressure").shape[0]) if self._has("fields/pressure") else 0

    def load_geometry(self):
        return np.asarray(self._array("materials/geometry")[:]).astype(bool) if self._has("materials/geometry") else None

    def close(self) -> None:
        if self.file is not None:
            self.file.close()
            self.file = None
        if self._npz is not None:
            self._npz.close()
            self._npz = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.close()
