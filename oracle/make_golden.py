#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference exists:

    ./oracle/build_ref.sh && python oracle/make_golden.py

For every case in tests/cases.py the reference's own ``FDTDSolver(backend="native")``
(with the fixed-native ADE arrangement for material cases) is stepped
``case["steps"]`` times and the following is stored:

* ``dt`` and the fp32 update coefficients, sponge sigma profiles and decay tables
* resolved probe / source flat indices (integer parity)
* every probe and microphone trace
* final p, vx, vy, vz: full arrays for small cases, SHA-256 digests always

The fixtures pin oracle/ (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_parity_gpu.py, GPU box, where /root/reference does not exist).
"""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from cases import c1_case, make_cases  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

FULL_FIELD_LIMIT = 40_000      # cells; above this only digests + a strided sample are stored


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(name: str, case: dict, out_dir: Path):
    s = R.build_reference_solver(case)
    assert s.using_native
    for _ in range(case["steps"]):
        s.step()
    out = {"dt": np.float64(s.dt), "steps": np.int64(case["steps"])}
    if s._grid.is_uniform:
        out["cv"] = np.float32(s._coeff_v); out["cp"] = np.float32(s._coeff_p)
    else:
        out["cv"] = np.float32(s._coeff_v_base); out["cp"] = np.float32(s._coeff_p_base)
        for key, arr in s._spacing_arrays.items():
            out["sp_" + key] = arr
    for bi, b in enumerate(s._boundaries):
        if not hasattr(b, "_max_sigma"):
            continue                       # Mur / RadiationImpedance: nothing tabulated
        out[f"pml{bi}_max_sigma"] = np.float64(b._max_sigma)
        for a, sig in zip("xyz", (b._sigma_x, b._sigma_y, b._sigma_z)):
            if sig is not None:
                out[f"pml{bi}_sigma_{a}"] = sig
        # PMLData does not expose its tables (kernels.cpp:279-284): read them back by damping ones
        k = R.load_ref_kernels()
        ones = [np.ones(s.shape, dtype=np.float32) for _ in range(3)]
        k.apply_pml_velocity(ones[0], ones[1], ones[2], b._pml_data)
        if b._sigma_x is not None: out[f"pml{bi}_decay_x"] = ones[0][:, 0, 0].copy()
        if b._sigma_y is not None: out[f"pml{bi}_decay_y"] = ones[1][0, :, 0].copy()
        if b._sigma_z is not None: out[f"pml{bi}_decay_z"] = ones[2][0, 0, :].copy()
    ny, nz = s.shape[1], s.shape[2]
    for pname, probe in s._probes.items():
        i, j, k = probe.position
        out["probe_idx_" + pname] = np.int64((i * ny + j) * nz + k)
        out["probe_" + pname] = probe.get_data()
    for si, src in enumerate(s._sources):
        if src.source_type == "point":
            i, j, k = src.position
            out[f"source_idx_{si}"] = np.int64((i * ny + j) * nz + k)
    for mname, mic in s.microphones.items():
        out["mic_" + mname] = np.array(mic._data, dtype=np.float32)
    if s._microphone_data is not None:
        # MicrophoneData hides its tables (kernels.cpp:361-366).  Recover them by one-hot probing:
        # the 8 corners are floor(pos/dx)+{0,1}; a unit pressure at corner c makes the batch kernel
        # return exactly w[c] (0 + w*1), and 0 if the kernel does not gather from that cell.
        k = R.load_ref_kernels()
        n_m = len(s._microphone_names_order)
        idx = np.zeros(8 * n_m, dtype=np.int64); w = np.zeros(8 * n_m, dtype=np.float32)
        outbuf = np.zeros(n_m, dtype=np.float32)
        for m, mname in enumerate(s._microphone_names_order):
            g = np.array(s.microphones[mname]._grid_position, dtype=np.float32)
            i0, j0, k0 = (int(v) for v in g)
            for c in range(8):
                i, j, kk = i0 + (c & 1), j0 + ((c >> 1) & 1), k0 + ((c >> 2) & 1)
                onehot = np.zeros(s.shape, dtype=np.float32); onehot[i, j, kk] = 1.0
                k.record_microphones_batch(onehot, s._microphone_data, outbuf)
                idx[8 * m + c] = (i * ny + j) * nz + kk
                w[8 * m + c] = outbuf[m]
            assert abs(float(w[8 * m:8 * m + 8].sum()) - 1.0) < 1e-5
        out["mic_flat_indices"] = idx
        out["mic_weights"] = w
    for f in ("p", "vx", "vy", "vz"):
        arr = getattr(s, f)
        out["sha_" + f] = np.array(digest(arr))
        out["absmax_" + f] = np.float32(np.abs(arr).max())
        if arr.size <= FULL_FIELD_LIMIT:
            out["final_" + f] = arr
        else:
            out["sample_" + f] = arr[::3, ::3, ::3].copy()
    np.savez_compressed(out_dir / f"{name}.npz", **out)
    print(f"{name}: steps={case['steps']} |p|max={float(np.abs(s.p).max()):.3e} "
          f"-> {(out_dir / (name + '.npz')).stat().st_size / 1024:.0f} KiB")


def main():
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    only = set(sys.argv[1:])                  # optional: regenerate just the named fixtures
    for name, case in make_cases().items():
        if not only or name in only:
            run_case(name, case, out_dir)
    if not only or "c1_100cubed_1000" in only:
        run_case("c1_100cubed_1000", c1_case(1000), out_dir)
    if not only or "membranes" in only:
        run_membrane_case(out_dir)


def run_membrane_case(out_dir: Path):
    """Membrane sources (solver.py:210-755, injection :2389-2412) through the reference's own classes.
    The injection-weight arrays they produce are stored too, so the case can be rebuilt without the reference."""
    from cases import MEMBRANE_SPEC, membrane_case
    import numpy as _np
    placeholder = [_np.zeros(MEMBRANE_SPEC["shape"]) for _ in MEMBRANE_SPEC["membranes"]]
    case = membrane_case(placeholder)
    s = R.build_reference_solver(case)
    for _ in range(case["steps"]):
        s.step()
    out = {"dt": np.float64(s.dt), "steps": np.int64(case["steps"])}
    for q, src in enumerate(s._sources):
        w = np.asarray(src._cached_weights, dtype=np.float64)
        nz_idx = np.flatnonzero(w)
        out[f"weights_idx_{q}"] = nz_idx.astype(np.int64)
        out[f"weights_val_{q}"] = w.ravel()[nz_idx]
    for pname, probe in s._probes.items():
        out["probe_" + pname] = probe.get_data()
    for f in ("p", "vx", "vy", "vz"):
        out["sha_" + f] = np.array(digest(getattr(s, f)))
        out["final_" + f] = getattr(s, f)
    np.savez_compressed(out_dir / "membranes.npz", **out)
    print(f"membranes: |p|max={float(np.abs(s.p).max()):.3e} nnz weights "
          f"{[int(len(out[f'weights_idx_{q}'])) for q in range(len(s._sources))]}")


if __name__ == "__main__":
    main()
