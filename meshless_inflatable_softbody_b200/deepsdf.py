"""DeepSDF: the reference's `DeepSDFWithCode` (deepsdf.py:9-41) evaluated by the tcgen05 GEMM chain.

Interface mirrors the reference: built from the same state dict (`model_{k}.pth`, sim.py:57-60; keys
`network.{i}.bias`, `network.{i}.parametrizations.weight.original0/1`), called on `[n,3]` model-space
coordinates, returns `[n,1]` fp32 (`sdf(points_torch)`, sim.py:100).  No CPU path: construction raises
without a CUDA device or the built library.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import native

# asset placement of the reference (sim.py:46,52): p_world = p_model @ R + lift
ASSET_R = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]])
ASSET_LIFT = np.array([0.0, 0.07, 0.0])


def world_to_model_xform(R=ASSET_R, lift=ASSET_LIFT) -> np.ndarray:
    """12 floats (A row-major, t) with p_model = A (p_world - t): the inverse of p_world = p_model @ R + lift."""
    A = np.asarray(R, np.float64)          # (p - t) @ R^T as a column-vector product is R (p - t)
    return np.concatenate([A.reshape(-1), np.asarray(lift, np.float64)]).astype(np.float32)


class DeepSDF:
    def __init__(self, state_dict, device: str = "cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("meshless_inflatable_softbody_b200.DeepSDF needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        self.L = native.lib()
        idx = sorted({int(k.split(".")[1]) for k in state_dict if k.startswith("network.")})
        if not idx:
            raise ValueError("state dict has no 'network.*' entries (deepsdf.py:12)")
        def dev(t):
            t = t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t))
            return t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self._g = [dev(state_dict[f"network.{i}.parametrizations.weight.original0"]) for i in idx]
        self._v = [dev(state_dict[f"network.{i}.parametrizations.weight.original1"]) for i in idx]
        self._b = [dev(state_dict[f"network.{i}.bias"]) for i in idx]
        n_layers = len(idx)
        dims = [int(self._v[0].shape[1])] + [int(v.shape[0]) for v in self._v]
        self.dims = dims
        self.hidden = dims[1]
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream(device=self.device)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        arr = lambda ts: (C.c_void_p * n_layers)(*[t.data_ptr() for t in ts])
        self._h = C.c_void_p()
        native.check(self.L.mis_sdf_create(n_layers, (C.c_int * (n_layers + 1))(*dims), arr(self._g), arr(self._v), arr(self._b),
                                           C.c_void_p(self.stream.cuda_stream), C.byref(self._h)), "mis_sdf_create")

    @classmethod
    def from_module(cls, module, device: str = "cuda:0") -> "DeepSDF":
        """From a torch module with the reference's structure (e.g. deepsdf.DeepSDFWithCode)."""
        return cls(module.state_dict(), device=device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.stream.synchronize()
            self.L.mis_sdf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _pts(self, coords) -> torch.Tensor:
        t = coords if isinstance(coords, torch.Tensor) else torch.as_tensor(np.asarray(coords, dtype=np.float32))
        t = t.to(device=self.device, dtype=torch.float32).reshape(-1, 3).contiguous()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        return t

    def query(self, points, xform: Optional[Sequence[float]] = None, grad: bool = False, fd_eps: float = 1e-3):
        """sdf [n] (and d sdf / d p [n,3] in the frame of `points`) at `points`; xform = world_to_model_xform() for world-space input."""
        p = self._pts(points)
        n = int(p.shape[0])
        sdf = torch.empty(n, device=self.device, dtype=torch.float32)
        g = torch.empty((n, 3), device=self.device, dtype=torch.float32) if grad else None
        xf = None if xform is None else (C.c_float * 12)(*[float(v) for v in xform])
        native.check(self.L.mis_sdf_query(self._h, p.data_ptr(), n, xf, sdf.data_ptr(), g.data_ptr() if grad else None,
                                          float(fd_eps), C.c_void_p(self.stream.cuda_stream)), "mis_sdf_query")
        self.stream.synchronize()
        return (sdf, g) if grad else sdf

    def forward(self, coords) -> torch.Tensor:
        """DeepSDFWithCode.forward (deepsdf.py:40-41): [n,3] -> [n,1]."""
        return self.query(coords).reshape(-1, 1)

    __call__ = forward

    def design_field(self, points_model, out_num: int) -> torch.Tensor:
        """sim.py:100-101: x = sdf(points).squeeze(); x[:out_num] = clip(x[:out_num], 1, None)."""
        x = self.query(points_model).clone()
        x[:out_num] = torch.clamp(x[:out_num], min=1.0)
        return x

    def set_gemm_path(self, path: int) -> None:
        """0 = automatic, 1 = split-K cluster kernel, one launch per layer, 2 = persistent big-tile kernel (bulk),
        3 = split-K cluster kernel, all hidden layers in one cooperative launch (few rows: per-step contact)."""
        native.check(self.L.mis_sdf_set_gemm_path(self._h, int(path)), "mis_sdf_set_gemm_path")

    def profile_gemm(self, m: int, reps: int = 10) -> float:
        """Device milliseconds per hidden-layer GEMM launch on m rows (CUDA events)."""
        ms = C.c_double(0)
        native.check(self.L.mis_sdf_profile_gemm(self._h, int(m), int(reps), C.c_void_p(self.stream.cuda_stream), C.byref(ms)),
                     "mis_sdf_profile_gemm")
        return ms.value / reps

    @property
    def launch_counts(self):
        g = C.c_longlong(0)
        tot = int(self.L.mis_sdf_launch_count(self._h, C.byref(g)))
        return tot, int(g.value)
