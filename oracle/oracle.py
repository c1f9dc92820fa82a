"""CPU oracle for the strata-fdtd time-stepping hot path (TEST INFRASTRUCTURE ONLY).

Nothing in the product package (``strata_fdtd_b200``) imports this module.  It is
used by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, always as the checker or the CPU
baseline, never as the thing measured or shipped.

Parity status: PINNED -- ``OracleSolver`` is checked bit-for-bit against the
reference's ``FDTDSolver(backend="native")`` (tests/test_oracle_vs_reference.py,
run wherever /root/reference exists) and against the committed fixtures in
``tests/golden`` (tests/test_oracle_golden.py, runs everywhere).

The numeric kernels are the plain-C restatement in ``oracle_step.c``; this file
restates the *host* logic of the reference (all citations are file:line under
/root/reference/src/strata_fdtd/):

* dt / update coefficients ............ core/solver.py:1576-1592
* nonuniform spacing arrays ........... core/grid.py:219-237, 498-532
* sponge ("PML") sigma profiles ....... boundaries/_boundaries.py:178-204, 228-315
* step ordering ....................... core/solver.py:2003-2077, 2079-2126
* ADE ordering (fixed-native order) ... core/solver.py:2135-2193 with _kernels/src/ade.cpp
* source injection / probes / mics .... core/solver.py:2386-2439, 2475-2518
* Gaussian pulse ...................... core/solver.py:188-207
* ADE pole coefficients ............... materials/base.py:157-188

A *case* is a plain dict (see tests/cases.py) so that the oracle, the real
reference and the b200 backend are all built from the same description.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> Path:
    """Compile oracle_step.c -> liboracle.so (gcc, -ffp-contract=off)."""
    so = _HERE / "liboracle.so"
    src = _HERE / "oracle_step.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        for name in dir(_LIB):
            pass
        for fn in ("orc_update_velocity", "orc_apply_rigid", "orc_update_pressure",
                   "orc_zero_solids", "orc_decay_table", "orc_sponge_velocity",
                   "orc_sponge_pressure", "orc_mic_precompute", "orc_mic_record",
                   "orc_ade_debye", "orc_ade_lorentz", "orc_ade_velocity_correction",
                   "orc_ade_pressure_correction", "orc_divergence"):
            getattr(_LIB, fn).restype = None
    return _LIB


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f32p)


def _up(a):
    return None if a is None else a.ctypes.data_as(_u8p)


F = ctypes.c_float
I = ctypes.c_int
SZ = ctypes.c_size_t


# --------------------------------------------------------------------------- host logic
def compute_spacing(coords: np.ndarray) -> np.ndarray:
    """core/grid.py:219-237 -- cell size = half the distance between the two neighbours."""
    n = len(coords)
    sp = np.zeros(n, dtype=np.float64)
    for i in range(1, n - 1):
        sp[i] = (coords[i + 1] - coords[i - 1]) / 2
    sp[0] = coords[1] - coords[0]
    sp[-1] = coords[-1] - coords[-2]
    return sp


def sigma_profile_uniform(n: int, depth: int, order: int, max_sigma) -> np.ndarray:
    """boundaries/_boundaries.py:228-255."""
    s = np.zeros(n, dtype=np.float32)
    d = depth
    for i in range(d):
        s[i] = max_sigma * (((d - i) / d) ** order)
    for i in range(n - d, n):
        s[i] = max_sigma * (((i - (n - d - 1)) / d) ** order)
    return s


def sigma_profile_nonuniform(coords, spacing, depth: int, order: int, max_sigma) -> np.ndarray:
    """boundaries/_boundaries.py:257-315."""
    n = len(coords)
    s = np.zeros(n, dtype=np.float32)
    d = depth
    if d >= n // 2:
        d = max(1, n // 4)
    left_t = float(np.sum(spacing[:d]))
    right_t = float(np.sum(spacing[-d:]))
    left_if = coords[d] - spacing[d] / 2 if d < n else coords[-1]
    right_if = coords[n - d - 1] + spacing[n - d - 1] / 2 if n - d > 0 else coords[0]
    for i in range(d):
        nd = min(1.0, (left_if - coords[i]) / left_t) if left_t > 0 else 0.0
        s[i] = max_sigma * (nd ** order)
    for i in range(n - d, n):
        nd = min(1.0, (coords[i] - right_if) / right_t) if right_t > 0 else 0.0
        s[i] = max_sigma * (nd ** order)
    return s


def gaussian_pulse(t: float, frequency: float, bandwidth: float | None, amplitude: float) -> float:
    """core/solver.py:188-207, evaluated exactly like solver.py:2416 (1-element array)."""
    if bandwidth is None:
        bandwidth = 2.0 * frequency
    tt = np.array([t])
    sigma = 1.0 / (np.pi * bandwidth)
    t0 = 4.0 * sigma
    env = np.exp(-((tt - t0) ** 2) / (2 * sigma**2))
    car = np.sin(2 * np.pi * frequency * (tt - t0))
    return (amplitude * env * car)[0]


def pole_coefficients(pole: dict, dt) -> tuple:
    """materials/base.py:157-188."""
    if pole["type"] == "debye":
        tau_dt = pole["tau"] / dt
        denom = 1 + tau_dt
        return (tau_dt / denom, pole["delta_chi"] / denom)
    w0, g = pole["omega_0"], pole["gamma"]
    dt2 = dt * dt
    c = 1 + g * dt / 2
    return ((2 - w0**2 * dt2) / c, -(1 - g * dt / 2) / c, pole["delta_chi"] * w0**2 * dt2 / c)


def resolve_position(pos, shape, dx):
    """core/solver.py:1818-1839 / 1864-1883 -- metres -> int(round(pos/dx)) if any float < max(shape)."""
    if isinstance(pos, tuple) and len(pos) == 3:
        if any(isinstance(q, float) and q < max(shape) for q in pos):
            gp = tuple(int(round(q / dx)) for q in pos)
            for idx, dim in zip(gp, shape):
                if not 0 <= idx < dim:
                    raise ValueError("position outside grid")
            return gp
    return pos


class OracleSolver:
    """Step-for-step restatement of FDTDSolver(backend='native') + the fixed-native ADE order."""

    def __init__(self, case: dict):
        self.case = case
        c = self.c = case.get("c", 343.0)
        rho = self.rho = case.get("rho", 1.2)
        courant = case.get("courant", 0.95)
        nu = case.get("nonuniform")
        if nu is None:
            self.shape = tuple(case["shape"])
            self.dx = case["resolution"]
            self.uniform = True
            self.coords = [np.arange(n) * self.dx + self.dx / 2 for n in self.shape]
            self.spacing = [np.full(n, self.dx, dtype=np.float64) for n in self.shape]
        else:
            self.coords = [np.asarray(nu[a], dtype=np.float64).ravel() for a in ("x_coords", "y_coords", "z_coords")]
            self.spacing = [compute_spacing(q) for q in self.coords]
            self.shape = tuple(len(q) for q in self.coords)
            self.dx = float(min(np.min(s) for s in self.spacing))
            self.uniform = False
        nx, ny, nz = self.shape
        # core/solver.py:1578-1580 (np.float64 arithmetic)
        self.dt = courant * (self.dx / (c * np.sqrt(3)))
        if self.uniform:
            self.cp = np.float32(-rho * c**2 * self.dt / self.dx)
            self.cv = np.float32(-self.dt / (rho * self.dx))
            self.inv_face = [None] * 3
            self.inv_cell = [None] * 3
        else:
            self.cp = np.float32(-rho * c**2 * self.dt)
            self.cv = np.float32(-self.dt / rho)
            self.inv_face = [(1.0 / np.diff(q)).astype(np.float32) for q in self.coords]
            self.inv_cell = [(1.0 / s).astype(np.float32) for s in self.spacing]
        z = lambda: np.zeros(self.shape, dtype=np.float32)
        self.p, self.vx, self.vy, self.vz = z(), z(), z(), z()
        g = case.get("geometry")
        self.rigid = g is not None              # solver.py:1779-1780: only after set_geometry
        self.geometry = np.ones(self.shape, dtype=bool) if g is None else np.ascontiguousarray(g, dtype=bool)
        self._geom_u8 = self.geometry.view(np.uint8)
        # sponges
        self.sponges = []
        for b in case.get("pml", []):
            self.sponges.append(self._init_sponge(b))
        self.plane_ops = self._init_plane_ops(case.get("plane_bcs", []))
        # sources / probes / mics
        self.sources = []
        for s in case.get("sources", []):
            s = dict(s)
            if s.get("kind", "point") == "point":
                s["position"] = resolve_position(s["position"], self.shape, self.dx)
            self.sources.append(s)
        self.probes = [(name, resolve_position(pos, self.shape, self.dx)) for name, pos in case.get("probes", [])]
        self.probe_data = {name: [] for name, _ in self.probes}
        self.mics = list(case.get("mics", []))
        self.mic_data = {m[0]: [] for m in self.mics}
        self._mic_tables = None
        # materials
        self.materials = case.get("materials", [])
        self.material_id = case.get("material_id")
        if self.materials:
            self.material_id = np.ascontiguousarray(self.material_id, dtype=np.uint8)
            self._init_ade()
        self.step_count = 0
        self.time = 0.0

    # ---- setup -------------------------------------------------------------
    def _init_sponge(self, b: dict) -> dict:
        depth, order = b.get("depth", 10), b.get("order", 3)
        axes = b.get("axes", ("x", "y", "z"))
        max_sigma = b.get("max_sigma")
        if max_sigma is None:      # _boundaries.py:182-187
            d = depth * self.dx
            max_sigma = -(order + 1) * self.c * np.log(1e-6) / (2 * d)
        sig, dec = [], []
        for a, name in enumerate("xyz"):
            if name not in axes:
                sig.append(None); dec.append(None); continue
            n = self.shape[a]
            if self.uniform:
                s = sigma_profile_uniform(n, depth, order, max_sigma)
            else:
                s = sigma_profile_nonuniform(self.coords[a], self.spacing[a], depth, order, max_sigma)
            d_ = np.empty(n, dtype=np.float32)
            lib().orc_decay_table(_fp(s), I(n), F(self.dt), _fp(d_))   # pml.cpp:13-45, dt cast to float (kernels.cpp)
            sig.append(s); dec.append(d_)
        return {"sigma": sig, "decay": dec, "max_sigma": max_sigma}

    def _init_ade(self):
        """Pole lists in the order of solver.py:3024-3054 (materials in registration order)."""
        self.debye, self.lorentz = [], []
        for m in self.materials:
            for pole in m["poles"]:
                co = pole_coefficients(pole, self.dt)
                ent = {"mat": m["id"], "target": pole["target"], "co": tuple(np.float32(x) for x in co),
                       "rho_inf": np.float32(m["rho_inf"]), "K_inf": np.float32(m["K_inf"]),
                       "J": np.zeros(self.shape, dtype=np.float32)}
                if pole["type"] == "debye":
                    self.debye.append(ent)
                else:
                    ent["Jp"] = np.zeros(self.shape, dtype=np.float32)
                    self.lorentz.append(ent)
        self._div = np.zeros(self.shape, dtype=np.float32)

    # ---- per-step phases ---------------------------------------------------
    def _ade_poles(self, target: str, src: np.ndarray):
        n = SZ(self.p.size)
        for e in self.debye:
            if e["target"] == target:
                lib().orc_ade_debye(_fp(e["J"]), _fp(src), _up(self.material_id), n, I(e["mat"]),
                                    F(e["co"][0]), F(e["co"][1]))
        for e in self.lorentz:
            if e["target"] == target:
                lib().orc_ade_lorentz(_fp(e["J"]), _fp(e["Jp"]), _fp(src), _up(self.material_id), n,
                                      I(e["mat"]), F(e["co"][0]), F(e["co"][1]), F(e["co"][2]))

    def step(self):
        L = lib()
        nx, ny, nz = self.shape
        dims = (I(nx), I(ny), I(nz))
        ade = bool(self.materials)
        if ade:                                                     # solver.py:2138-2139
            self._ade_poles("density", self.p)
        L.orc_update_velocity(_fp(self.p), _fp(self.vx), _fp(self.vy), _fp(self.vz), *dims, F(self.cv),
                              *(_fp(a) for a in self.inv_face))     # solver.py:2087 / 2111
        if ade:                                                     # solver.py:2148-2149, ade.cpp:228-401
            inv_dx = np.float32(1.0 / self.dx)                      # solver.py:3304 (uniform 1/dx_min always)
            for e in self.debye + self.lorentz:
                if e["target"] == "density":
                    L.orc_ade_velocity_correction(_fp(self.vx), _fp(self.vy), _fp(self.vz), _fp(e["J"]),
                                                  _up(self.material_id), *dims, I(e["mat"]),
                                                  F(e["rho_inf"]), F(self.dt), F(inv_dx))
        if self.rigid:                                              # solver.py:2094-2097
            L.orc_apply_rigid(_fp(self.vx), _fp(self.vy), _fp(self.vz), _up(self._geom_u8), *dims)
        if ade:                                                     # solver.py:2155-2156, 3150-3185
            L.orc_divergence(_fp(self._div), _fp(self.vx), _fp(self.vy), _fp(self.vz), *dims,
                             F(np.float32(1.0 / self.dx)), *(_fp(a) for a in self.inv_cell))
            self._ade_poles("modulus", self._div)
        L.orc_update_pressure(_fp(self.p), _fp(self.vx), _fp(self.vy), _fp(self.vz), _up(self._geom_u8),
                              *dims, F(self.cp), *(_fp(a) for a in self.inv_cell))   # solver.py:2100 / 2123
        if ade:                                                     # solver.py:2189-2193
            n = SZ(self.p.size)
            for e in self.debye + self.lorentz:
                if e["target"] == "modulus":
                    L.orc_ade_pressure_correction(_fp(self.p), _fp(e["J"]), _up(self.material_id), n,
                                                  I(e["mat"]), F(e["K_inf"]), F(self.dt))
            L.orc_zero_solids(_fp(self.p), _up(self._geom_u8), n)
        for sp in self.sponges:                                     # solver.py:2044-2045
            L.orc_sponge_velocity(_fp(self.vx), _fp(self.vy), _fp(self.vz), *dims, *(_fp(d) for d in sp["decay"]))
        for sp in self.sponges:                                     # solver.py:2046-2047
            L.orc_sponge_pressure(_fp(self.p), *dims, *(_fp(d) for d in sp["decay"]))
        for op in self.plane_ops:                                   # Mur / RadiationImpedance planes, same phase
            self._apply_plane_op(op)
        self._inject()
        for name, (i, j, k) in self.probes:                         # solver.py:2435-2439
            self.probe_data[name].append(float(self.p[i, j, k]))
        self._record_mics()
        self.step_count += 1
        self.time += self.dt                                        # solver.py:2071-2072

    def _init_plane_ops(self, specs):
        """boundaries/_boundaries.py:420-526 (Mur) and :529-793 (RadiationImpedance), in list order."""
        mur = (self.c * self.dt - self.dx) / (self.c * self.dt + self.dx)          # np.float64
        ops = []
        for b in specs:
            if b["kind"] == "mur":
                for a, name in enumerate("xyz"):
                    if name in b.get("axes", ("x", "y", "z")):
                        for side in (0, 1):
                            ops.append(dict(kind="mur", axis=a, side=side, mur=mur))
            else:
                if b.get("reflection_coeff") is not None:
                    R = b["reflection_coeff"]                                        # Python float
                else:
                    ka = (2 * np.pi * 1000 / self.c) * b["pipe_radius"]
                    z = (ka ** 2) / 4
                    R = min(abs((z - 1) / (z + 1)), 0.95)
                ops.append(dict(kind="radiation", axis="xyz".index(b["axis"]), side=0 if b["side"] == "low" else 1,
                                mur=mur, R=R))
        for op in ops:
            shp = tuple(n for q, n in enumerate(self.shape) if q != op["axis"])
            op["prev"] = np.zeros(shp, dtype=np.float32)
        return ops

    def _apply_plane_op(self, op):
        """The reference's NumPy expressions verbatim in type behaviour: fp32 difference, float64 coefficient
        (np.float64 scalar), Python-float R multiplies in fp32 (NEP 50 weak scalar), one rounding on store."""
        sl_b = [slice(None)] * 3; sl_i = [slice(None)] * 3
        sl_b[op["axis"]] = -1 if op["side"] else 0
        sl_i[op["axis"]] = -2 if op["side"] else 1
        p_b, p_i = self.p[tuple(sl_b)], self.p[tuple(sl_i)]
        abc = op["prev"] + op["mur"] * (p_i - p_b)
        if op["kind"] == "mur":
            p_b[...] = abc
        else:
            R = op["R"]
            p_b[...] = R * p_i + (1 - R) * abc
        op["prev"] = p_i.copy()

    def _inject(self):
        """solver.py:2386-2433: float64 add, float32 store."""
        for s in self.sources:
            w = gaussian_pulse(self.time, s["frequency"], s.get("bandwidth"), s.get("amplitude", 1.0))
            kind = s.get("kind", "point")
            if kind == "point":
                i, j, k = s["position"]
                if self.geometry[i, j, k]:
                    self.p[i, j, k] = np.float32(np.float64(self.p[i, j, k]) + w)
            elif kind == "plane":
                sl = [slice(None)] * 3
                sl[s["axis"]] = s["index"]
                view = self.p[tuple(sl)]
                m = self.geometry[tuple(sl)]
                view[m] = (view[m].astype(np.float64) + w).astype(np.float32)
            elif kind == "weighted":
                # membrane sources, solver.py:2389-2412: field[mask] += w * weights[mask], mask = weights>0 & air,
                # float64 product and sum, float32 store
                wt = np.asarray(s["weights"], dtype=np.float64)
                m = (wt > 0) & self.geometry
                fld = getattr(self, s.get("field", "p"))
                fld[m] = (fld[m].astype(np.float64) + w * wt[m]).astype(np.float32)
            else:
                raise ValueError(kind)

    # ---- directional microphones: the reference's Python path (core/solver.py:1004-1173) ----------
    _PATTERN_K = {"omni": 0.0, "subcardioid": 0.3, "cardioid": 0.5, "supercardioid": 0.63,
                  "hypercardioid": 0.75, "figure8": 1.0}

    @staticmethod
    def _tri(gx, gy, gz):
        """solver.py:962-1002 in float64."""
        i0, j0, k0 = int(gx), int(gy), int(gz)
        fx, fy, fz = gx - i0, gy - j0, gz - k0
        w = [(1 - fx) * (1 - fy) * (1 - fz), fx * (1 - fy) * (1 - fz), (1 - fx) * fy * (1 - fz), fx * fy * (1 - fz),
             (1 - fx) * (1 - fy) * fz, fx * (1 - fy) * fz, (1 - fx) * fy * fz, fx * fy * fz]
        idx = [(i0 + (c & 1), j0 + ((c >> 1) & 1), k0 + ((c >> 2) & 1)) for c in range(8)]
        return idx, w

    def _record_mics_python(self):
        nx, ny, nz = self.shape
        for m in self.mics:
            name, pos = m[0], m[1]
            opt = m[2] if len(m) > 2 else {}
            g = tuple(q / self.dx for q in pos)
            pressure = 0.0
            for (i, j, k), w in zip(*self._tri(*g)):
                pressure += w * self.p[i, j, k]                      # float64 weight * fp32 -> fp32 (NEP 50)
            pattern = opt.get("pattern", "omni")
            if pattern == "omni":
                self.mic_data[name].append(pressure)
                continue
            vel = []
            for axis, fld in enumerate((self.vx, self.vy, self.vz)):  # solver.py:1060-1100
                gg = list(g)
                gg[axis] -= 0.5
                gg = [max(0.0, min(q, n - 1.001)) for q, n in zip(gg, self.shape)]
                v = 0.0
                for (i, j, k), w in zip(*self._tri(*gg)):
                    v += w * fld[min(i, nx - 1), min(j, ny - 1), min(k, nz - 1)]
                vel.append(v)
            d = np.array(opt.get("direction", (1.0, 0.0, 0.0)), dtype=np.float64)
            d = d / np.linalg.norm(d)
            v_dot_d = vel[0] * d[0] + vel[1] * d[1] + vel[2] * d[2]
            if callable(pattern):
                v_mag = np.sqrt(vel[0] ** 2 + vel[1] ** 2 + vel[2] ** 2)
                theta = np.arccos(np.clip(v_dot_d / v_mag, -1.0, 1.0)) if v_mag > 1e-20 else 0.0
                out = pressure * pattern(theta)
            else:
                k_ = self._PATTERN_K[pattern]
                out = (1 - k_) * pressure + k_ * (self.rho * self.c) * v_dot_d
            self.mic_data[name].append(out)

    def _record_mics(self):
        if not self.mics:
            return
        if any(len(m) > 2 and m[2].get("pattern", "omni") != "omni" for m in self.mics):
            return self._record_mics_python()                      # solver.py:2453-2461: Python path for all
        nx, ny, nz = self.shape
        if self._mic_tables is None:                                # solver.py:2496-2518
            gp = np.zeros(3 * len(self.mics), dtype=np.float32)
            for m, (_, pos, *_opt) in enumerate(self.mics):
                for a in range(3):
                    gp[3 * m + a] = pos[a] / self.dx                # solver.py:941-943 then fp32 store :2507-2509
            idx = np.zeros(8 * len(self.mics), dtype=np.int64)
            w = np.zeros(8 * len(self.mics), dtype=np.float32)
            lib().orc_mic_precompute(_fp(gp), I(len(self.mics)), I(ny), I(nz), idx.ctypes.data_as(_i64p), _fp(w))
            self._mic_tables = (idx, w)
        idx, w = self._mic_tables
        out = np.zeros(len(self.mics), dtype=np.float32)
        lib().orc_mic_record(_fp(self.p), idx.ctypes.data_as(_i64p), _fp(w), I(len(self.mics)), _fp(out))
        for m, (name, *_rest) in enumerate(self.mics):
            self.mic_data[name].append(float(out[m]))

    def run_steps(self, n: int):
        for _ in range(n):
            self.step()

    # ---- results -----------------------------------------------------------
    def probe_array(self, name): return np.array(self.probe_data[name], dtype=np.float32)
    def mic_array(self, name): return np.array(self.mic_data[name], dtype=np.float32)


def set_threads(n: int | None = None) -> int:
    """OpenMP team size for the oracle kernels (does not change results)."""
    n = n or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n
