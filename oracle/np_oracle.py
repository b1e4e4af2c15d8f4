"""Independent fp64 numpy restatement of the sim.py step (small n only).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED by the reference (no tests / golden
vectors; Warp and Taichi not installable here).  This restatement exists to pin
the C oracle (mis_oracle.c) from a second direction: brute-force all-pairs
neighbours instead of the hash grid, LAPACK SVD instead of the Jacobi polar
rotation, fp64 instead of fp32, vectorised tensor algebra instead of the
literal per-pair operation order.

Follows sim.py:133-151 (W, nabla_W), :154-167 (volume), :170-209 (A_pq, R, F),
:212-235 (stress, pair force with the F_i quirk at :233), :238-258 (ground
penalty, velocity Verlet).
"""
from __future__ import annotations

import numpy as np


def W(r, h):
    """sim.py:133-141, r = |xij| (array)."""
    q = r / h
    s = 1.0 / (np.pi * h ** 3)
    return np.where(q < 1, s * (1 - 1.5 * q ** 2 + 0.75 * q ** 3),
                    np.where(q < 2, s / 4 * (2 - q) ** 3, 0.0))


def nabla_W(xij, h):
    """sim.py:143-151, xij (..., 3)."""
    r = np.linalg.norm(xij, axis=-1)
    q = r / h
    s = 1.0 / (np.pi * h ** 3)
    with np.errstate(divide="ignore", invalid="ignore"):
        c1 = s * (-3.0 + 2.25 * q) / h ** 2
        c2 = s / 4 * (-3.0) * (2 - q) ** 2 / (q * h * h)
    c = np.where(q < 1, c1, np.where(q < 2, c2, 0.0))
    c = np.where(r > 0, c, 0.0)
    return c[..., None] * xij


def polar_rotation(A):
    """R = U V^T with det U = det V = +1 (reflection pushed into sigma_3), sim.py:185-191."""
    U, s, Vt = np.linalg.svd(A)
    dU = np.linalg.det(U)
    dV = np.linalg.det(Vt)
    U = U.copy()
    Vt = Vt.copy()
    U[..., :, 2] *= np.sign(dU)[..., None]
    Vt[..., 2, :] *= np.sign(dV)[..., None]
    return U @ Vt


class NpOracle:
    def __init__(self, x0, h=0.007, dt=5e-5, damping=1e-6, k_col=3e5, col_range=1e-4):
        self.x0 = np.asarray(x0, np.float64)
        self.n = len(self.x0)
        self.h, self.dt, self.damping, self.k_col, self.col_range = map(float, (h, dt, damping, k_col, col_range))
        n = self.n
        self.mass = np.zeros(n)
        self.E = np.zeros(n)
        self.nu = np.zeros(n)
        self.ratio = np.full(n, 0.5 * np.tanh(-3.0) + 0.5)
        self.fext = np.zeros((n, 3))
        self.free = np.ones((n, 3))
        d0 = self.x0[None, :, :] - self.x0[:, None, :]          # x0_j - x0_i  [i, j]
        self.d0 = d0
        r = np.linalg.norm(d0, axis=-1)
        self.mask = (r / self.h < 2) & ~np.eye(n, dtype=bool)
        self.w = W(r, self.h) * self.mask
        self.gw = nabla_W(-d0, self.h) * self.mask[..., None]   # nabla_W(x0_i - x0_j)

    def set_material(self, E, nu):
        self.E = np.broadcast_to(np.asarray(E, np.float64), (self.n,)).copy()
        self.nu = np.broadcast_to(np.asarray(nu, np.float64), (self.n,)).copy()
        self.mu = self.E / (2 * (1 + self.nu))
        self.lam = self.E * self.nu / ((1 + self.nu) * (1 - 2 * self.nu))

    def set_mass(self, m):
        self.mass = np.broadcast_to(np.asarray(m, np.float64), (self.n,)).copy()
        self.rho = (self.w * self.mass[None, :]).sum(1)
        self.vol = self.mass / self.rho

    def set_design(self, x):
        self.ratio = 0.5 * np.tanh(3.0 * np.broadcast_to(np.asarray(x, np.float64), (self.n,))) + 0.5

    def penalty(self, x):
        pen = np.zeros_like(x)
        below = x[:, 1] < self.col_range
        pen[below, 1] = (self.col_range - x[below, 1]) ** 2 * self.k_col
        return pen

    def eval(self, x):
        x = np.asarray(x, np.float64)
        dx = x[None, :, :] - x[:, None, :]                       # x_j - x_i
        A = np.einsum("ij,ija,ijb->iab", self.w * self.mass[None, :], dx, self.d0)
        R = polar_rotation(A)
        u = np.einsum("iba,ijb->ija", R, dx) - self.d0           # R^T (x_j - x_i) - d0
        N = np.einsum("j,ija,ijb->iab", self.vol, u, self.gw)
        F = np.eye(3)[None] + np.swapaxes(N, 1, 2)
        Es = 0.5 * (np.swapaxes(F, 1, 2) @ F - np.eye(3)[None])
        tr = np.trace(Es, axis1=1, axis2=2)
        S = (2 * self.mu[:, None, None] * Es + (self.lam * tr)[:, None, None] * np.eye(3)[None]) \
            * (200.0 - 199.0 * self.ratio)[:, None, None]
        # force_i = 1/2 V_i [ sum_j V_j R_j F_i S_j gw_ij + R_i F_i S_i sum_j V_j gw_ij ]
        Sg = np.einsum("jab,ijb->ija", S, self.gw)               # S_j gw_ij
        FSg = np.einsum("iab,ijb->ija", F, Sg)                   # F_i S_j gw_ij   (sim.py:233)
        t1 = np.einsum("j,jab,ijb->ia", self.vol, R, FSg)
        g = np.einsum("j,ija->ia", self.vol, self.gw)
        t2 = np.einsum("iab,ib->ia", R @ F @ S, g)
        f = 0.5 * self.vol[:, None] * (t1 + t2)
        return dict(A=A, R=R, F=F, S=S, f=f)

    def startup(self, v0=(0.0, -0.4, 0.0)):
        self.x = self.x0.copy()
        self.v = np.tile(np.asarray(v0, np.float64), (self.n, 1))
        self.fel = self.eval(self.x)["f"]

    def step(self, n_steps=1):
        dt, m = self.dt, self.mass[:, None]
        for _ in range(n_steps):
            f1 = self.fext + self.fel - self.damping * self.v + self.penalty(self.x)
            xn = self.x + (dt * self.v + 0.5 * dt * dt * f1 / m) * self.free
            feln = self.eval(xn)["f"]
            f2 = self.fext + feln - self.damping * self.v + self.penalty(xn)
            self.v = self.v + (dt * (f1 + f2) / (2 * m)) * self.free
            self.x, self.fel = xn, feln
