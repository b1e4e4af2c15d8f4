"""A numpy stand-in for the `warp` module, just large enough to EXECUTE the reference's own kernel source.

TEST INFRASTRUCTURE (fixture generation only).  NVIDIA Warp is not installed here and the reference hard-codes
device="cuda", so `/root/reference/sim.py` cannot be imported.  Its kernels, however, are plain Python under
`@wp.func` / `@wp.kernel` (sim.py:107-110, 133-273) and so are its setters and its rollout loop (sim.py:279-322,
341-372).  `lift()` takes those top-level statements out of the reference file BY LINE RANGE with `ast` (the module top
level -- argparse, asset loading, DeepSDF checkpoint -- is never executed, nothing is copied into this repo) and runs
them in a namespace where `wp` is this module.  The fixtures written by make_simpy_golden.py are therefore outputs of
the reference's own arithmetic, statement by statement; what is NOT the reference's is stated here:

  * wp.svd3        third-party (Warp native code, not in the reference tree).  Here: LAPACK SVD in the working
                   precision with both factors made proper rotations (a reflection goes into the sign of the smallest
                   singular value), which is the documented contract of Warp's svd3.  R = U V^T is then the unique
                   proper polar rotation whenever A is non-singular, whichever SVD algorithm produced it.
  * wp.HashGrid    third-party.  Two interchangeable candidate generators: "grid" (truncating cell coordinates,
                   +2^20 offset, mod dim, 27-cell walk x fastest, points in ascending id inside a cell -- Warp's
                   native/hashgrid.h as far as it is known here) and "brute" (every index).  The reference's kernels
                   filter by their own compact support (W, nabla_W vanish for q >= 2), so both give the same sums up
                   to fp32 summation order; make_simpy_golden.py asserts that.
  * mat @ mat, mat @ vec, wp.outer, wp.length   evaluated with separate fp32 multiplies and adds in index order
                   (Warp's generated CUDA may contract them into FMAs: that difference is inside the summation-order
                   noise floor the parity tests measure).

Precision: every value is a numpy scalar/array of the working type (`wp.float32` -> np.float32), so an expression such
as `real(1.) / (real(wp.pi) * h * h * h)` rounds after every operation exactly as typed.
"""
from __future__ import annotations

import ast
import math

import numpy as np

pi = math.pi
float32 = np.float32
float64 = np.float64
uint64 = np.uint64
int32 = np.int32


# ------------------------------------------------------------------------------------------------ value types
class _Val(np.ndarray):
    """vec3 / mat33 value (or a writable view of one array element)."""

    def __matmul__(self, other):
        return _matmul(self, other)

    def __rmatmul__(self, other):
        return _matmul(other, self)


def _matmul(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.ndim == 2 and b.ndim == 1:          # mat @ vec: sum over k in order 0,1,2
        p = a * b[None, :]
        return (p[:, 0] + p[:, 1] + p[:, 2]).view(_Val)
    if a.ndim == 2 and b.ndim == 2:          # mat @ mat
        p = a[:, :, None] * b[None, :, :]
        return (p[:, 0, :] + p[:, 1, :] + p[:, 2, :]).view(_Val)
    raise TypeError("unsupported matmul operands")


def _make_vec_type(dtype):
    def ctor(*args):
        if len(args) == 0:
            return np.zeros(3, dtype).view(_Val)
        if len(args) == 1:
            return np.array(np.asarray(args[0]), dtype=dtype).reshape(3).view(_Val)
        return np.array(args, dtype=dtype).reshape(3).view(_Val)
    ctor._wp_shape = (3,)
    ctor._wp_dtype = dtype
    return ctor


def _make_mat_type(dtype):
    def ctor(*args):
        if len(args) == 0:
            return np.zeros((3, 3), dtype).view(_Val)
        if len(args) == 1:
            return np.array(np.asarray(args[0]), dtype=dtype).reshape(3, 3).view(_Val)
        return np.array(args, dtype=dtype).reshape(3, 3).view(_Val)
    ctor._wp_shape = (3, 3)
    ctor._wp_dtype = dtype
    return ctor


vec3 = vec3f = _make_vec_type(np.float32)
vec3d = _make_vec_type(np.float64)
mat33 = mat33f = _make_mat_type(np.float32)
mat33d = _make_mat_type(np.float64)


# ------------------------------------------------------------------------------------------------ arrays
class array:  # noqa: N801  (name mirrors wp.array)
    """wp.array: also used as a type annotation (`wp.array(dtype=real)`), hence every argument is optional."""

    def __init__(self, data=None, dtype=None, shape=None, device=None, requires_grad=False):
        inner = getattr(dtype, "_wp_shape", ())
        base = getattr(dtype, "_wp_dtype", dtype)
        self.dtype = dtype
        self._inner = inner
        if data is not None:
            a = np.array(np.asarray(data), dtype=base)
            if inner and a.shape == inner:
                a = a.reshape((1,) + inner)
            elif not inner and a.ndim == 0:
                a = a.reshape(1)
            self.a = a
        elif shape is not None:
            shape = (shape,) if np.isscalar(shape) else tuple(shape)
            self.a = np.zeros(shape + inner, base)
        else:
            self.a = None          # annotation use
        self._grad = None

    @property
    def shape(self):
        return self.a.shape[:self.a.ndim - len(self._inner)]

    @property
    def grad(self):
        if self._grad is None:
            self._grad = array(shape=self.shape, dtype=self.dtype)
        return self._grad

    def __getitem__(self, i):
        r = self.a[i]
        return r.view(_Val) if self._inner else r

    def __setitem__(self, i, v):
        self.a[i] = v

    def numpy(self):
        return self.a.copy()

    def fill_(self, v):
        self.a[...] = np.asarray(v, dtype=self.a.dtype)


def from_numpy(arr, dtype=None, device=None, requires_grad=False):
    inner = getattr(dtype, "_wp_shape", ())
    base = getattr(dtype, "_wp_dtype", dtype)
    out = array(shape=np.asarray(arr).shape[:np.asarray(arr).ndim - len(inner)], dtype=dtype)
    out.a[...] = np.asarray(arr).astype(base)
    return out


def copy(dest, src, dest_offset=0, src_offset=0, count=0):
    count = count or len(src.a)
    dest.a[dest_offset:dest_offset + count] = src.a[src_offset:src_offset + count]


# ------------------------------------------------------------------------------------------------ builtins
def outer(a, b):
    return (np.asarray(a)[:, None] * np.asarray(b)[None, :]).view(_Val)


def transpose(m):
    return np.array(np.asarray(m).T).view(_Val)


def identity(n, dtype):
    return np.eye(n, dtype=dtype).view(_Val)


def trace(m):
    m = np.asarray(m)
    return m[0, 0] + m[1, 1] + m[2, 2]


def length_sq(v):
    v = np.asarray(v)
    return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]


def length(v):
    return np.sqrt(length_sq(v))


def cw_mul(a, b):
    return (np.asarray(a) * np.asarray(b)).view(_Val)


def tanh(x):
    return np.tanh(x)


def atomic_add(arr, i, v):
    arr.a[i] += v


def svd3(A, U, sigma, V):
    """Contract of wp.svd3: A = U diag(sigma) V^T with U, V proper rotations (see the module docstring)."""
    a = np.asarray(A)
    u, s, vt = np.linalg.svd(a)
    v = vt.T.copy()
    if np.linalg.det(u) < 0:
        u[:, 2] = -u[:, 2]
        s[2] = -s[2]
    if np.linalg.det(v) < 0:
        v[:, 2] = -v[:, 2]
        s[2] = -s[2]
    U[...] = u.astype(a.dtype)
    sigma[...] = s.astype(a.dtype)
    V[...] = v.astype(a.dtype)


# ------------------------------------------------------------------------------------------------ kernels / launch
_tid = 0


def tid():
    return _tid


def func(f):
    return f


def kernel(f):
    f._wp_kernel = True
    return f


def launch(kernel, dim, inputs, outputs=(), device=None):  # noqa: A002
    """One Python call per thread; scalar arguments are cast to their annotated type as Warp does at launch."""
    global _tid
    args = list(inputs) + list(outputs)
    ann = getattr(kernel, "__annotations__", {})
    names = kernel.__code__.co_varnames[:kernel.__code__.co_argcount]
    for k, name in enumerate(names):
        t = ann.get(name)
        if isinstance(t, type) and issubclass(t, np.floating) and not isinstance(args[k], array):
            args[k] = t(args[k])
    for t in range(int(dim)):
        _tid = t
        kernel(*args)


class Tape:
    """wp.Tape as a no-op context (the forward values do not depend on it)."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def backward(self, *a, **k):
        raise NotImplementedError("reverse mode is outside the fixtures")

    def zero(self):
        pass


# ------------------------------------------------------------------------------------------------ hash grid
QUERY_MODE = "grid"      # "grid" | "brute"; module-level switch read by hash_grid_query


class HashGrid:
    def __init__(self, dim_x, dim_y, dim_z, device=None):
        self.dim = (int(dim_x), int(dim_y), int(dim_z))
        self.id = self
        self.points = None

    def _cell(self, cx, cy, cz):
        o = 1 << 20
        dx, dy, dz = self.dim
        x, y, z = max(0, cx + o) % dx, max(0, cy + o) % dy, max(0, cz + o) % dz
        return z * (dx * dy) + y * dx + x

    def build(self, points, radius):
        self.points = points
        self.cell_width = np.float32(radius)
        self.cell_width_inv = np.float32(1.0) / self.cell_width
        p = points.a.astype(np.float32)
        cells = np.empty(len(p), np.int64)
        for i in range(len(p)):
            c = [int(np.float32(p[i, a] * self.cell_width_inv)) for a in range(3)]   # int(): truncation toward zero
            cells[i] = self._cell(*c)
        order = np.argsort(cells, kind="stable")          # radix sort by cell: ascending id inside a cell
        self.point_ids = order.astype(np.int64)
        ncell = self.dim[0] * self.dim[1] * self.dim[2]
        self.cell_members = [[] for _ in range(ncell)]
        for i in order:
            self.cell_members[cells[i]].append(int(i))


def hash_grid_point_id(grid, t):
    return int(grid.point_ids[t])


def hash_grid_query(grid, point, max_dist):
    if QUERY_MODE == "brute":
        return range(len(grid.points.a))
    p = np.asarray(point, dtype=np.float32)
    r = np.float32(max_dist)
    inv = grid.cell_width_inv
    lo = [int(np.float32((p[a] - r) * inv)) for a in range(3)]
    hi = [min(int(np.float32((p[a] + r) * inv)), lo[a] + grid.dim[a] - 1) for a in range(3)]
    out = []
    for cz in range(lo[2], hi[2] + 1):
        for cy in range(lo[1], hi[1] + 1):
            for cx in range(lo[0], hi[0] + 1):
                out.extend(grid.cell_members[grid._cell(cx, cy, cz)])
    return out


# ------------------------------------------------------------------------------------------------ lifting
def lift(path, ranges, namespace, filename=None):
    """Execute the top-level statements of `path` whose first line lies in one of `ranges` (inclusive, 1-based)
    inside `namespace`.  Decorated functions are taken with their decorators.  Nothing else of the file runs."""
    with open(path) as fh:
        src = fh.read()
    tree = ast.parse(src, filename=filename or path)
    picked = []
    for node in tree.body:
        first = min([node.lineno] + [d.lineno for d in getattr(node, "decorator_list", [])])
        if any(a <= first <= b for a, b in ranges):
            picked.append(node)
    mod = ast.Module(body=picked, type_ignores=[])
    exec(compile(mod, filename or path, "exec"), namespace)
    return [(n.lineno, getattr(n, "name", type(n).__name__)) for n in picked]
