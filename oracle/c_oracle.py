"""ctypes front-end of the C oracle (oracle/mis_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of mis_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  PARITY UNPINNED by the reference (it ships no tests or
golden vectors and cannot run here); pins are the known-answer tests, the fp64
numpy restatement in np_oracle.py and brute-force neighbour search.

The class mirrors the reference's script-level control functions
(sim.py:279-308, 341-358) so parity tests read like the reference's own loop.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmis_oracle.so")
_LIB_PATH_F64 = os.path.join(_HERE, "libmis_oracle_f64.so")     # the same source built with -DORC_DOUBLE (real = double)

FAITHFUL = 0   # sim.py:218-235 as written (per-candidate svd3, every candidate visited)
CACHED = 1     # R_j, S_j once per particle; bit-identical results


def _params_struct(real):
    class _P(C.Structure):
        _fields_ = [
            ("h", real), ("damping", real), ("dt", real),
            ("k_col", real), ("col_range", real),
            ("grid_x", C.c_int), ("grid_y", C.c_int), ("grid_z", C.c_int),
            ("symmetric_pair", C.c_int), ("identity_rot", C.c_int), ("self_density", C.c_int),
            ("euler", C.c_int), ("no_contact", C.c_int),
            ("stiff_a", real), ("stiff_b", real),
        ]
    return _P


OrcParams = _params_struct(C.c_float)
OrcParams64 = _params_struct(C.c_double)


def build(force: bool = False) -> str:
    """Compile libmis_oracle.so (real = float) and libmis_oracle_f64.so (real = double, -DORC_DOUBLE) next to their source
    (gcc, OpenMP, no FMA contraction)."""
    src = os.path.join(_HERE, "mis_oracle.c")
    for path, extra in ((_LIB_PATH, []), (_LIB_PATH_F64, ["-DORC_DOUBLE"])):
        if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
            cmd = [cc, "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC"] + extra + \
                  ["-shared", "-o", path + ".tmp", src, "-lm"]
            subprocess.run(cmd, check=True, cwd=_HERE)
            os.replace(path + ".tmp", path)
    return _LIB_PATH


_libs = {}


def lib(precision: str = "f32"):
    """The oracle library in the given working precision ("f32" = sim.py's wp.float32, "f64" = sim_taichi.py's ti.f64)."""
    if precision not in _libs:
        build()
        f64 = precision == "f64"
        L = C.CDLL(_LIB_PATH_F64 if f64 else _LIB_PATH)
        real = C.c_double if f64 else C.c_float
        fp = C.POINTER(real)
        ip = C.POINTER(C.c_int)
        vp = C.c_void_p
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.c_int, fp, C.POINTER(OrcParams64 if f64 else OrcParams), C.c_int]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_threads.argtypes = [vp, C.c_int]
        L.orc_set_order.argtypes = [vp, C.c_int]
        for name in ("orc_set_youngs_modulus", "orc_set_poisson_ratio", "orc_set_mass",
                     "orc_set_external_forces", "orc_set_free_points", "orc_set_ratio"):
            getattr(L, name).argtypes = [vp, fp]
        L.orc_set_design.argtypes = [vp, fp, real]
        L.orc_startup.argtypes = [vp, fp, C.c_int]
        L.orc_set_state.argtypes = [vp, fp, fp, C.c_int]
        L.orc_step.argtypes = [vp, C.c_int, C.c_int]
        L.orc_eval.argtypes = [vp, fp, C.c_int, fp, fp, fp, fp, fp]
        L.orc_get_state.argtypes = [vp, fp, fp]
        L.orc_get_forces.argtypes = [vp, fp]
        L.orc_get_fields.argtypes = [vp, fp, fp]
        L.orc_get_volume.argtypes = [vp, fp, fp]
        L.orc_get_lame.argtypes = [vp, fp, fp, fp]
        L.orc_get_grid.argtypes = [vp, ip, ip, ip]
        L.orc_num_cells.argtypes = [vp]
        L.orc_num_cells.restype = C.c_int
        L.orc_get_cell_ranges.argtypes = [vp, ip, ip]
        L.orc_neighbor_lists.argtypes = [vp, ip, C.POINTER(C.c_longlong), ip, C.c_longlong]
        L.orc_neighbor_lists.restype = C.c_longlong
        L.orc_candidate_count.argtypes = [vp]
        L.orc_candidate_count.restype = C.c_longlong
        L.orc_W.argtypes = [real] * 4
        L.orc_W.restype = real
        L.orc_nabla_W.argtypes = [real] * 4 + [fp]
        L.orc_polar.argtypes = [fp, fp]
        L.orc_sigma.argtypes = [fp, real, real, real, fp]
        L.orc_max_threads.restype = C.c_int
        L._real = real
        L._np = np.float64 if f64 else np.float32
        _libs[precision] = L
    return _libs[precision]


def _f32(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
    return a


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double if a.dtype == np.float64 else C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def grid_dims(x0: np.ndarray, h: float):
    """sim.py:123-125: int(2 * extent / float(h) / 3) per axis, float(h) being the fp32 value."""
    hf = float(np.float32(h))
    p = np.asarray(x0, dtype=np.float64)
    return tuple(max(1, int(2 * (p[:, a].max() - p[:, a].min()) / hf / 3)) for a in range(3))


class Oracle:
    """CPU oracle of the sim.py step.  Method names follow sim.py:279-308, 341-358."""

    def __init__(self, x0, h=0.007, dt=5e-5, damping=1e-6, k_col=3e5, col_range=1e-4,
                 threads=0, variant="warp", grid=None, precision="f32", **flags):
        self.L = lib(precision)
        self.dt_np = self.L._np
        self.x0 = self._a(x0).reshape(-1, 3)
        self.n = self.x0.shape[0]
        gx, gy, gz = grid if grid is not None else grid_dims(self.x0, h)
        p = (OrcParams64 if precision == "f64" else OrcParams)()
        p.h, p.dt, p.damping, p.k_col, p.col_range = h, dt, damping, k_col, col_range
        p.grid_x, p.grid_y, p.grid_z = gx, gy, gz
        p.stiff_a, p.stiff_b = 200.0, 199.0
        if variant == "taichi":          # SURVEY 2.2 deltas (precision="f64" for the prototype's ti.f64)
            p.symmetric_pair = p.identity_rot = p.self_density = p.euler = p.no_contact = 1
            p.stiff_a, p.stiff_b = 1.0, 1.0
        for k, v in flags.items():
            setattr(p, k, v)
        self.params = p
        self.grid = (gx, gy, gz)
        self.h = h
        self.tanh_k = 5.0 if variant == "taichi" else 3.0
        self.o = self.L.orc_create(self.n, _fp(self.x0), C.byref(p), int(threads))
        self.mode = CACHED

    def _a(self, a, shape=None):
        a = np.ascontiguousarray(np.asarray(a, dtype=self.dt_np))
        if shape is not None:
            a = np.ascontiguousarray(np.broadcast_to(a, shape))
        return a

    def __del__(self):
        try:
            if getattr(self, "o", None):
                self.L.orc_destroy(self.o)
                self.o = None
        except Exception:
            pass

    # --- control functions (sim.py:279-308) ---
    def set_all_external_force(self, f):
        self.L.orc_set_external_forces(self.o, _fp(self._a(f, (self.n, 3))))

    def set_external_forces(self, f):
        self.L.orc_set_external_forces(self.o, _fp(self._a(f, (self.n, 3))))

    def set_free_points(self, d):
        self.L.orc_set_free_points(self.o, _fp(self._a(d, (self.n, 3))))

    def set_youngs_modulus(self, E):
        self.L.orc_set_youngs_modulus(self.o, _fp(self._a(E, (self.n,))))

    def set_poisson_ratio(self, nu):
        self.L.orc_set_poisson_ratio(self.o, _fp(self._a(nu, (self.n,))))

    def set_mass(self, m):
        self.L.orc_set_mass(self.o, _fp(self._a(m, (self.n,))))

    def set_design(self, x):
        self.L.orc_set_design(self.o, _fp(self._a(x, (self.n,))), self.L._real(self.tanh_k))

    def set_ratio(self, r):
        self.L.orc_set_ratio(self.o, _fp(self._a(r, (self.n,))))

    def set_threads(self, t):
        self.L.orc_set_threads(self.o, int(t))

    def set_order(self, order):
        self.L.orc_set_order(self.o, int(order))

    # --- rollout (sim.py:341-358) ---
    def startup(self, v0=(0.0, -0.4, 0.0), mode=None):
        self.L.orc_startup(self.o, _fp(self._a(v0)), self.mode if mode is None else mode)

    def set_state(self, x, v, mode=None):
        self.L.orc_set_state(self.o, _fp(self._a(x)), _fp(self._a(v)), self.mode if mode is None else mode)

    def step(self, n_steps=1, mode=None):
        self.L.orc_step(self.o, int(n_steps), self.mode if mode is None else mode)

    def eval(self, pos, mode=None):
        n = self.n
        out = {k: np.empty((n, 3, 3), self.dt_np) for k in ("A", "R", "F", "S")}
        out["f"] = np.empty((n, 3), self.dt_np)
        self.L.orc_eval(self.o, _fp(self._a(pos)), self.mode if mode is None else mode,
                        _fp(out["A"]), _fp(out["R"]), _fp(out["F"]), _fp(out["S"]), _fp(out["f"]))
        return out

    # --- state export ---
    def position(self):
        x = np.empty((self.n, 3), self.dt_np)
        self.L.orc_get_state(self.o, _fp(x), None)
        return x

    def velocity(self):
        v = np.empty((self.n, 3), self.dt_np)
        self.L.orc_get_state(self.o, None, _fp(v))
        return v

    def elastic_forces(self):
        f = np.empty((self.n, 3), self.dt_np)
        self.L.orc_get_forces(self.o, _fp(f))
        return f

    def fields(self):
        A = np.empty((self.n, 3, 3), self.dt_np)
        F = np.empty((self.n, 3, 3), self.dt_np)
        self.L.orc_get_fields(self.o, _fp(A), _fp(F))
        return A, F

    def volume(self):
        rho = np.empty(self.n, self.dt_np)
        vol = np.empty(self.n, self.dt_np)
        self.L.orc_get_volume(self.o, _fp(rho), _fp(vol))
        return rho, vol

    def lame(self):
        mu = np.empty(self.n, self.dt_np)
        lam = np.empty(self.n, self.dt_np)
        ratio = np.empty(self.n, self.dt_np)
        self.L.orc_get_lame(self.o, _fp(mu), _fp(lam), _fp(ratio))
        return mu, lam, ratio

    # --- neighbour structure ---
    def grid_arrays(self):
        cell = np.empty(self.n, np.int32)
        coords = np.empty((self.n, 3), np.int32)
        ids = np.empty(self.n, np.int32)
        self.L.orc_get_grid(self.o, _ip(cell), _ip(coords), _ip(ids))
        return cell, coords, ids

    def cell_ranges(self):
        nc = self.L.orc_num_cells(self.o)
        s = np.empty(nc, np.int32)
        e = np.empty(nc, np.int32)
        self.L.orc_get_cell_ranges(self.o, _ip(s), _ip(e))
        return s, e

    def neighbor_lists(self):
        counts = np.empty(self.n, np.int32)
        offsets = np.empty(self.n + 1, np.int64)
        total = self.L.orc_neighbor_lists(self.o, _ip(counts), offsets.ctypes.data_as(C.POINTER(C.c_longlong)), None, 0)
        flat = np.empty(total, np.int32)
        self.L.orc_neighbor_lists(self.o, _ip(counts), offsets.ctypes.data_as(C.POINTER(C.c_longlong)), _ip(flat), total)
        return counts, offsets, flat

    def candidate_count(self):
        return int(self.L.orc_candidate_count(self.o))


def W(xij, h):
    x = np.asarray(xij, np.float32)
    return float(lib().orc_W(float(x[0]), float(x[1]), float(x[2]), float(h)))


def nabla_W(xij, h):
    x = np.asarray(xij, np.float32)
    out = np.empty(3, np.float32)
    lib().orc_nabla_W(float(x[0]), float(x[1]), float(x[2]), float(h), _fp(out))
    return out


def polar(A):
    A = _f32(A).reshape(3, 3)
    R = np.empty((3, 3), np.float32)
    lib().orc_polar(_fp(A), _fp(R))
    return R


def sigma(F, mu, lam, ratio):
    F = _f32(F).reshape(3, 3)
    S = np.empty((3, 3), np.float32)
    lib().orc_sigma(_fp(F), float(mu), float(lam), float(ratio), _fp(S))
    return S


def max_threads():
    return int(lib().orc_max_threads())
