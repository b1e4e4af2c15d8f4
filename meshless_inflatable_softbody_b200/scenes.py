"""Synthetic scene generators.

The reference loads scanned point clouds from absolute paths that do not exist
(sim.py:27,41-53); all inputs here are synthetic but keep its conventions: the
cloud is lifted to y = +0.07 (sim.py:52), lengths are metres after the 0.01
scale (sim.py:47-48), "outer" particles come first (sim.py:49,53).
"""
from __future__ import annotations

import numpy as np


def jittered_sphere(n_target: int, h: float = 0.007, spacing: float = 0.5, jitter: float = 0.2,
                    seed: int = 0, centre=(0.0, 0.07, 0.0), low_drop: bool = False):
    """Jittered cubic lattice clipped to a sphere.

    spacing is in units of h (0.5 => ~160-220 neighbours/particle, the stable regime of the
    reference defaults; 0.8 => ~57).  Returns (x0 float32 (n,3), out_num) where the first
    out_num particles are the outer shell r > R - 2h (the reference's `out_num`, sim.py:53).
    """
    s = spacing * h
    R = (3.0 * n_target * s ** 3 / (4.0 * np.pi)) ** (1.0 / 3.0)
    m = int(np.ceil(R / s)) + 1
    ax = np.arange(-m, m + 1, dtype=np.float64) * s
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    g = g + rng.uniform(-jitter * s, jitter * s, size=g.shape)
    r = np.linalg.norm(g, axis=1)
    g = g[r <= R]
    r = r[r <= R]
    outer = r > R - 2 * h
    g = np.concatenate([g[outer], g[~outer]], 0)
    out_num = int(outer.sum())
    c = np.asarray(centre, np.float64).copy()
    if low_drop:   # lowest particle 0.6 mm above the ground plane: impact starts at step ~30
        c[1] = 0.0006 - g[:, 1].min()
    x0 = (g + c).astype(np.float32)
    return x0, out_num


def jittered_beam(n_target: int, h: float = 0.007, spacing: float = 0.5, jitter: float = 0.2,
                  seed: int = 0, aspect=(4.0, 1.0, 1.0), centre=(0.0, 0.07, 0.0)):
    """Jittered lattice clipped to an aspect[0]:aspect[1]:aspect[2] box (config 5's beam)."""
    s = spacing * h
    a = np.asarray(aspect, np.float64)
    unit = (n_target * s ** 3 / a.prod()) ** (1.0 / 3.0)
    half = 0.5 * unit * a
    axes = [np.arange(-int(np.ceil(hh / s)), int(np.ceil(hh / s)) + 1) * s for hh in half]
    g = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    g = g + rng.uniform(-jitter * s, jitter * s, size=g.shape)
    keep = np.all(np.abs(g) <= half[None, :], axis=1)
    g = g[keep]
    return (g + np.asarray(centre, np.float64)).astype(np.float32)


def jittered_ellipsoid(n_target: int, h: float = 0.007, spacing: float = 0.5, jitter: float = 0.2, seed: int = 0,
                       aspect=(4.0, 1.0, 1.0), centre=(0.0, 0.07, 0.0), low_drop: bool = False):
    """Jittered lattice clipped to an ellipsoid with semi-axes proportional to `aspect` (long axis x by default):
    the slab-partitioned multi-GPU scene -- smooth surface like the sphere, long enough to cut into slabs."""
    s = spacing * h
    a = np.asarray(aspect, np.float64)
    unit = (3.0 * n_target * s ** 3 / (4.0 * np.pi * a.prod())) ** (1.0 / 3.0)
    semi = unit * a
    axes = [np.arange(-int(np.ceil(r / s)) - 1, int(np.ceil(r / s)) + 2) * s for r in semi]
    g = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    g = g + rng.uniform(-jitter * s, jitter * s, size=g.shape)
    g = g[((g / semi[None, :]) ** 2).sum(1) <= 1.0]
    c = np.asarray(centre, np.float64).copy()
    if low_drop:
        c[1] = 0.0006 - g[:, 1].min()
    return (g + c).astype(np.float32)


def shell_mask(x0: np.ndarray, h: float, centre=None):
    """Particles within 2h of the surface of the (assumed spherical) cloud."""
    c = x0.mean(0) if centre is None else np.asarray(centre)
    r = np.linalg.norm(x0 - c, axis=1)
    return r > r.max() - 2 * h


def plateau_obstacle_state(radius: float, y_top: float, hidden: int = 1024, n_linear: int = 9):
    """State dict (the reference's DeepSDFWithCode keys, deepsdf.py:12-38) of a ReLU MLP that computes EXACTLY

        sdf(p) = max( (|x| + |y| + |z| - radius) / sqrt(3),  y - y_top ),

    an octahedron whose upper tip is cut off by the plane y = y_top: a flat square plateau |x| + |z| <= radius - y_top
    a soft body can rest on, so a contact patch holds tens to hundreds of particles (a sharp tip holds a handful).
    Layer 0: relu(+-x), relu(+-y), relu(+-z); layer 1: c = relu(a - b), and y+ = relu(y), y- = relu(-y) passed through;
    hidden layers pass (c, y+, y-) through; the last layer returns c + y+ - y- - y_top = b + relu(a - b) = max(a, b).
    Every Linear is weight-normalised (W = g v / |v| per row): v = the row, g = its norm (g = 0 silences a row)."""
    if n_linear < 3:
        raise ValueError("needs at least one hidden-to-hidden layer")
    dims = [3] + [hidden] * (n_linear - 1) + [1]
    st = {}
    k = 1.0 / np.sqrt(3.0)
    for l in range(n_linear):
        o, i = dims[l + 1], dims[l]
        W = np.zeros((o, i), np.float64)
        b = np.zeros(o, np.float64)
        if l == 0:
            for a in range(3):
                W[2 * a, a] = 1.0; W[2 * a + 1, a] = -1.0
        elif l == 1:
            W[0, :6] = k; W[0, 2] = k - 1.0; W[0, 3] = k + 1.0; b[0] = y_top - radius * k
            W[1, 2] = 1.0; W[2, 3] = 1.0
        elif l < n_linear - 1:
            for u in range(3):
                W[u, u] = 1.0
        else:
            W[0, 0] = 1.0; W[0, 1] = 1.0; W[0, 2] = -1.0; b[0] = -y_top
        g = np.sqrt((W * W).sum(1, keepdims=True))
        v = W.copy()
        v[g[:, 0] == 0.0, 0] = 1.0            # keep |v| non-zero; g = 0 silences the row
        st[f"network.{3 * l}.parametrizations.weight.original0"] = g.astype(np.float32)
        st[f"network.{3 * l}.parametrizations.weight.original1"] = v.astype(np.float32)
        st[f"network.{3 * l}.bias"] = b.astype(np.float32)
    return st


def plateau_obstacle_bbox(radius: float, y_top: float, margin: float):
    """(min xyz, max xyz) of the region where the plateau obstacle's sdf can be below `margin`."""
    return [-radius - margin, -radius - margin, -radius - margin, radius + margin, y_top + margin, radius + margin]
