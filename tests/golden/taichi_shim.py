"""A numpy stand-in for the `taichi` module, just large enough to EXECUTE the reference's Taichi prototype.

TEST INFRASTRUCTURE (fixture generation only); see warp_shim.py for the rationale.  Taichi is not installed here and
`sim_taichi.py` does `ti.init(arch=ti.gpu)` and loads point clouds at import, so it cannot be imported; its kernels
(sim_taichi.py:78-234, utils.py:25-43) and setters (240-294) are plain Python under `@ti.kernel` / `@ti.func` and are
executed from the reference file by line range (warp_shim.lift) with `ti` bound to this module.

Semantics kept: fields are dense arrays indexed `[frame, i]`; `ti.ndrange(n, n)` is the full i-j product (the reference
has no neighbour search here: O(N^2)); top-level `for` loops of a kernel run serially (Taichi parallelises them, `+=`
on a field being atomic -- same sums, other order); default_fp = f64 (options.py:3).  `ti.svd` (third-party) is LAPACK
with proper-rotation factors; the prototype overwrites R with the identity (sim_taichi.py:129) so it never reaches the
forward values.
"""
from __future__ import annotations

import itertools
import math as _math
import types

import numpy as np

f64 = np.float64
f32 = np.float32
i32 = np.int32
gpu = "gpu"


class _T(np.ndarray):
    """Field or value: numpy array with the handful of Taichi methods the prototype calls."""

    def outer_product(self, other):
        return (np.asarray(self)[:, None] * np.asarray(other)[None, :]).view(_T)

    def norm(self):
        v = np.asarray(self)
        return np.sqrt((v * v).sum())

    def from_numpy(self, a):
        self[...] = np.asarray(a, dtype=self.dtype)

    def to_numpy(self):
        return np.array(self)

    def transpose(self):          # 3x3 value (fields are never transposed by the prototype)
        return np.array(np.asarray(self).T).view(_T)


def field(dtype=f64, shape=(), needs_grad=False):
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    return np.zeros(shape, dtype).view(_T)


class Vector:
    def __new__(cls, data):
        return np.array(data, dtype=f64).view(_T)

    @staticmethod
    def field(n, dtype=f64, shape=(), needs_grad=False):
        shape = (shape,) if np.isscalar(shape) else tuple(shape)
        return np.zeros(shape + (n,), dtype).view(_T)

    @staticmethod
    def zero(dtype, n):
        return np.zeros(n, dtype).view(_T)


class Matrix:
    def __new__(cls, data):
        return np.array(data, dtype=f64).view(_T)

    @staticmethod
    def field(n, m, dtype=f64, shape=(), needs_grad=False):
        shape = (shape,) if np.isscalar(shape) else tuple(shape)
        return np.zeros(shape + (n, m), dtype).view(_T)

    @staticmethod
    def zero(dtype, n, m):
        return np.zeros((n, m), dtype).view(_T)

    @staticmethod
    def identity(dtype, n):
        return np.eye(n, dtype=dtype).view(_T)


def kernel(f):
    return f


def func(f):
    return f


ad = types.SimpleNamespace(grad_replaced=lambda f: f, grad_for=lambda fwd: (lambda f: f))
math = types.SimpleNamespace(pi=_math.pi, vec3=object)


def ndrange(*dims):
    return itertools.product(*[range(int(d)) for d in dims])


def static(x):
    return x


def tanh(x):
    return np.tanh(x)


def atomic_add(target, v):
    target += v


def svd(A):
    a = np.asarray(A)
    u, s, vt = np.linalg.svd(a)
    v = vt.T.copy()
    if np.linalg.det(u) < 0:
        u[:, 2] = -u[:, 2]
        s[2] = -s[2]
    if np.linalg.det(v) < 0:
        v[:, 2] = -v[:, 2]
        s[2] = -s[2]
    return u.view(_T), np.diag(s).view(_T), v.view(_T)
