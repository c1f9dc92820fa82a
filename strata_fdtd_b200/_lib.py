"""ctypes binding of libstrata_b200.so (the C ABI declared in include/strata_b200.h).

The library is built in-tree by ``build()`` (nvcc, sm_100a, --fmad=false).  There is no
CPU fallback: if the shared object cannot be built or loaded, or no CUDA device is
present when a solver is created, the backend raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
ROOT = _PKG.parent
SO_PATH = Path(os.environ.get("STRATA_B200_LIB_OVERRIDE") or _PKG / "lib" / "libstrata_b200.so")   # override: kernel experiments
SOURCES = [_PKG / "csrc" / "sb_api.cu", _PKG / "csrc" / "sb_kernels.cuh", _PKG / "csrc" / "sb_resident.cuh",
           _PKG / "csrc" / "sb_pipeline.cuh",
           ROOT / "include" / "strata_b200.h"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
              "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


class B200BackendError(RuntimeError):
    """Raised for every failure of the native backend (build, load, or a non-zero C-ABI return)."""


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise B200BackendError("nvcc not found: cannot build libstrata_b200.so (no CPU fallback exists)")


def needs_build() -> bool:
    if not SO_PATH.exists():
        return True
    t = SO_PATH.stat().st_mtime
    return any(s.stat().st_mtime > t for s in SOURCES if s.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA extension for sm_100a (cross-compiles without a GPU).

    Safe under torchrun, where every rank may get here at once: one process compiles (file lock), into a temporary
    name that is moved over the library atomically, so nobody ever dlopens a half-written file."""
    if not force and not needs_build():
        return SO_PATH
    import fcntl
    SO_PATH.parent.mkdir(parents=True, exist_ok=True)
    with open(SO_PATH.parent / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():          # another rank built it while we waited
                return SO_PATH
            tmp = SO_PATH.with_name(f".{SO_PATH.name}.{os.getpid()}.tmp")
            cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(tmp), str(SOURCES[0])]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                tmp.unlink(missing_ok=True)
                raise B200BackendError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
            os.replace(tmp, SO_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO_PATH


class GridDesc(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("pitch", C.c_int32),
                ("global_nx", C.c_int64), ("i_offset", C.c_int64),
                ("has_lower", C.c_int32), ("has_upper", C.c_int32)]


class Pole(C.Structure):
    _fields_ = [("material_id", C.c_int32), ("is_lorentz", C.c_int32), ("target", C.c_int32),
                ("reserved", C.c_int32), ("c0", C.c_float), ("c1", C.c_float), ("c2", C.c_float),
                ("reserved_f", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("cells", C.c_int64), ("steps_done", C.c_int64), ("kernels_launched", C.c_int64),
                ("algorithmic_bytes_per_cell", C.c_double), ("kernel_variant", C.c_int32),
                ("pitch", C.c_int32)]


KERNEL_AUTO, KERNEL_NAIVE, KERNEL_MARCH, KERNEL_RESIDENT, KERNEL_PIPELINE = 0, 1, 2, 4, 5
(OPT_KERNEL, OPT_ROWS_PER_THREAD, OPT_WARPS_J, OPT_WARPS_K, OPT_CHUNK_I, OPT_USE_GRAPH, OPT_PROFILE,
 OPT_FUSE_K3, OPT_RESIDENT_SPLIT, OPT_RESIDENT_MIN_STEPS, OPT_PLANE_MAP, OPT_ADE_LAYOUT, OPT_ADE_CHUNK_I,
 OPT_ADE_WARPS, OPT_ADE_OCCUPANCY) = range(15)

_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
_fp = C.POINTER(C.c_float)

# name -> (restype, argtypes); every symbol include/strata_b200.h declares
SIGNATURES = {
    "sb_last_error": (C.c_char_p, []),
    "sb_abi_version": (_i, []),
    "sb_choose_pitch": (_i, [C.c_int32, C.POINTER(C.c_int32)]),
    "sb_field_elems": (_i64, [C.POINTER(GridDesc)]),
    "sb_create": (_i, [C.POINTER(GridDesc), _i, _vp, C.POINTER(_vp)]),
    "sb_destroy": (_i, [_vp]),
    "sb_bind_fields": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "sb_current_set": (_i, [_vp, C.POINTER(_i)]),
    "sb_upload_field": (_i, [_vp, _i, _vp]),
    "sb_download_field": (_i, [_vp, _i, _vp]),
    "sb_set_coefficients": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float]),
    "sb_set_geometry": (_i, [_vp, _vp, _i]),
    "sb_sponge_decay": (_i, [_vp, _i, C.c_float, _vp]),
    "sb_clear_sponges": (_i, [_vp]),
    "sb_add_sponge": (_i, [_vp, _vp, _vp, _vp]),
    "sb_clear_plane_ops": (_i, [_vp]),
    "sb_add_plane_op": (_i, [_vp, _i, _i, _i, C.c_double, C.c_double, _i]),
    "sb_set_ade": (_i, [_vp, C.POINTER(Pole), _i, _vp, _vp, _vp, _i, C.c_float, C.c_float]),
    "sb_set_sources": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "sb_set_probes": (_i, [_vp, _i, _vp]),
    "sb_set_mics": (_i, [_vp, _i, _vp, _vp]),
    "sb_set_gathers": (_i, [_vp, _i, _vp, _vp, _vp]),
    "sb_mic_tables": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "sb_ade_state": (_i, [_vp, _i, _i, _vp, _i]),
    "sb_plane_op_state": (_i, [_vp, _i, _vp, _vp, _i]),
    "sb_step_n": (_i, [_vp, _i, _vp, _vp]),
    "sb_step_n_async": (_i, [_vp, _i, _vp, _vp]),
    "sb_step_n_submit": (_i, [_vp, _i, _i, _vp, _vp]),
    "sb_step_n_wait": (_i, [_vp, _i]),
    "sb_step_cuts_async": (_i, [_vp, C.POINTER(_i)]),
    "sb_halo_planes": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64)]),
    "sb_set_peers": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), _i, _vp, _vp, _vp]),
    "sb_energy": (_i, [_vp, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "sb_reset": (_i, [_vp]),
    "sb_set_option": (_i, [_vp, _i, _i]),
    "sb_query": (_i, [_vp, C.POINTER(Stats)]),
    "sb_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i)]),
    "sb_tuned": (_i, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_float)]),
    "sb_synchronize": (_i, [_vp]),
}

_LIB = None


def load():
    """Load (building first if needed) and type the shared library.  Raises, never falls back."""
    global _LIB
    if _LIB is None:
        build()
        try:
            lib = C.CDLL(str(SO_PATH))
        except OSError as e:
            raise B200BackendError(f"cannot load {SO_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        if lib.sb_abi_version() != 1:
            raise B200BackendError("libstrata_b200.so ABI version mismatch")
        _LIB = lib
    return _LIB


def check(rc: int):
    if rc != 0:
        raise B200BackendError(load().sb_last_error().decode())


def ptr(a):
    """void* of a numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
