"""ctypes binding of the C-ABI library (include/mis.h) and its in-tree build.

The library is hand-written CUDA for sm_100a.  There is no fallback: if it cannot be
loaded, importing this module's `lib()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.environ.get("MIS_LIB") or os.path.join(_PKG, "libmis_b200.so")   # MIS_LIB: tuning builds only
SOURCES = ["mis_api.cu"]
HEADERS = ["mis_math.cuh", "mis_sort.cuh", "mis_neighbors.cuh", "mis_cluster.cuh", "mis_tile.cuh", "mis_ref.cuh", "mis_ref_host.cuh", "mis_sdf.cuh", "mis_sdf_host.cuh"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class MisParams(C.Structure):
    """Mirror of `struct MisParams` in include/mis.h."""
    _fields_ = [
        ("h", C.c_float), ("damping", C.c_float), ("dt", C.c_float), ("k_col", C.c_float), ("col_range", C.c_float),
        ("stiff_a", C.c_float), ("stiff_b", C.c_float), ("tanh_k", C.c_float),
        ("grid_x", C.c_int), ("grid_y", C.c_int), ("grid_z", C.c_int),
        ("symmetric_pair", C.c_int), ("identity_rot", C.c_int), ("self_density", C.c_int),
        ("euler", C.c_int), ("no_contact", C.c_int),
        ("lanes_per_particle", C.c_int), ("keep_fields", C.c_int), ("graph_steps", C.c_int),
        ("two_pass_deform", C.c_int), ("cluster_size", C.c_int), ("fp64", C.c_int),
    ]


class MisNeighborInfo(C.Structure):
    _fields_ = [
        ("total_pairs", C.c_longlong), ("max_neighbors", C.c_int), ("n", C.c_int),
        ("cell_min", C.c_int * 3), ("cell_dim", C.c_int * 3), ("cell_width", C.c_float),
        ("cluster_size", C.c_int), ("union_entries", C.c_longlong),
    ]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.access(cand, os.X_OK):
            return cand
    raise RuntimeError("nvcc not found: cannot build libmis_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(_PKG, "..", "include", "mis.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libmis_b200.so, in-tree."""
    if not (force or needs_build()):
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:      # torchrun ranks of one box must not run nvcc into the same file at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():           # another rank built it while this one waited
            return LIB_PATH
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", LIB_PATH + ".tmp"] + [os.path.join(_CSRC, f) for f in SOURCES]
        env = dict(os.environ)
        env.pop("CC", None)      # the image exports a gcc wrapper in CC; let nvcc pick the system gcc
        env.pop("CXX", None)
        r = subprocess.run(cmd, cwd=_CSRC, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        os.replace(LIB_PATH + ".tmp", LIB_PATH)       # a reader never sees a half-written library
        if verbose:
            print(r.stderr)
    return LIB_PATH


# name -> (restype, argtypes); every symbol include/mis.h declares
_vp, _fp, _ip = C.c_void_p, C.c_void_p, C.c_void_p   # device/host pointers travel as integers
SYMBOLS = {
    "mis_last_error": (C.c_char_p, []),
    "mis_version": (C.c_char_p, []),
    "mis_create": (C.c_int, [C.c_int, _fp, C.POINTER(MisParams), _vp, C.POINTER(C.c_void_p)]),
    "mis_destroy": (C.c_int, [_vp]),
    "mis_build_neighbors": (C.c_int, [_vp, _vp]),
    "mis_get_neighbor_info": (C.c_int, [_vp, C.POINTER(MisNeighborInfo)]),
    "mis_export_cells": (C.c_int, [_vp, _ip, _ip, _ip, _vp]),
    "mis_export_cell_ranges": (C.c_int, [_vp, _ip, _ip, _vp]),
    "mis_export_neighbors": (C.c_int, [_vp, _vp, _ip, _vp]),
    "mis_set_mass": (C.c_int, [_vp, _fp, _vp]),
    "mis_set_material": (C.c_int, [_vp, _fp, _fp, _vp]),
    "mis_set_design": (C.c_int, [_vp, _fp, _vp]),
    "mis_set_ext_force": (C.c_int, [_vp, _fp, _vp]),
    "mis_set_ext_force_host": (C.c_int, [_vp, _fp, _vp]),
    "mis_set_dirichlet": (C.c_int, [_vp, _fp, _vp]),
    "mis_startup": (C.c_int, [_vp, C.POINTER(C.c_float), _vp]),
    "mis_set_state": (C.c_int, [_vp, _fp, _fp, _vp]),
    "mis_step": (C.c_int, [_vp, C.c_int, _vp]),
    "mis_get_state": (C.c_int, [_vp, _fp, _fp, _vp]),
    "mis_get_state_host": (C.c_int, [_vp, _fp, _fp, _vp]),
    "mis_get_state_host_async": (C.c_int, [_vp, _fp, _fp, _vp]),
    "mis_wait_state_host": (C.c_int, [_vp, C.c_int]),
    "mis_get_fields": (C.c_int, [_vp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _vp]),
    "mis_eval_forces": (C.c_int, [_vp, _fp, _fp, _vp]),
    "mis_accumulate_loss": (C.c_int, [_vp, _fp, _fp, _vp, _vp]),
    "mis_set_gather_mode": (C.c_int, [_vp, C.c_int, _vp]),
    "mis_get_gather_info": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "mis_set_f64": (C.c_int, [_vp, C.c_int, _fp, _vp]),
    "mis_get_f64": (C.c_int, [_vp, C.c_int, _fp, _vp]),
    "mis_set_constants_f64": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "mis_startup_f64": (C.c_int, [_vp, C.POINTER(C.c_double), _vp]),
    "mis_rollout_grad": (C.c_int, [_vp, C.c_int, C.c_int, _fp, _fp, C.c_int, C.POINTER(C.c_double), _fp, _fp, _vp]),
    "mis_launch_count": (C.c_longlong, [_vp]),
    "mis_profile_step": (C.c_int, [_vp, C.c_int, _vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mis_gather_next_positions": (C.c_int, [_vp, _ip, C.c_int, _fp, _vp]),
    "mis_scatter_next_positions": (C.c_int, [_vp, _ip, C.c_int, _fp, _vp]),
    "mis_set_volumes": (C.c_int, [_vp, _ip, C.c_int, _fp, _vp]),
    "mis_export_slots": (C.c_int, [_vp, _ip, C.c_int, _ip, _vp]),
    "mis_halo_ipc_handles": (C.c_int, [_vp, C.c_char_p]),
    "mis_halo_local_ptrs": (C.c_int, [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "mis_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "mis_ipc_close": (C.c_int, [_vp]),
    "mis_halo_connect": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.c_int, _ip, _ip, _ip, C.c_int, _ip, _ip, _vp]),
    "mis_halo_disconnect": (C.c_int, [_vp]),
    "mis_halo_set_wait": (C.c_int, [_vp, C.c_int]),
    "mis_halo_status": (C.c_int, [_vp, _vp, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "mis_sdf_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_void_p), _vp, C.POINTER(C.c_void_p)]),
    "mis_sdf_destroy": (C.c_int, [_vp]),
    "mis_sdf_query": (C.c_int, [_vp, _fp, C.c_int, C.POINTER(C.c_float), _fp, _fp, C.c_float, _vp]),
    "mis_sdf_set_gemm_path": (C.c_int, [_vp, C.c_int]),
    "mis_sdf_launch_count": (C.c_longlong, [_vp, C.POINTER(C.c_longlong)]),
    "mis_sdf_profile_gemm": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.POINTER(C.c_double)]),
    "mis_set_sdf_contact": (C.c_int, [_vp, _vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, _vp]),
    "mis_get_contact_force": (C.c_int, [_vp, _fp, _vp]),
    "mis_get_contact_count": (C.c_int, [_vp, _vp, C.POINTER(C.c_int)]),
}

_lib = None


def lib():
    """Load libmis_b200.so (building it first if sources are newer).  Fails loudly."""
    global _lib
    if _lib is None:
        if needs_build():
            build()
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)       # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class MisError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().mis_last_error()
        raise MisError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
