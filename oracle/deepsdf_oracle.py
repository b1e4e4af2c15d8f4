"""CPU oracle of the DeepSDF MLP (deepsdf.py:9-41, class DeepSDFWithCode) and of the obstacle-contact law.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg;
the product path (meshless_inflatable_softbody_b200/) never imports it.

Parity pin: tests/golden/deepsdf_seed0.npz holds inputs and outputs produced by the reference's own
`deepsdf.DeepSDFWithCode` (imported from /root/reference by tests/golden/make_deepsdf_golden.py) together
with checksums of its seeded state dict; tests/test_deepsdf_oracle.py checks this restatement against them.

forward() restates nn.Sequential of deepsdf.py:12-38: weight_norm(Linear) (deepsdf.py:3,13: W = g * v / ||v||
with the norm over each output row), ReLU, Dropout(0.0) = identity (deepsdf.py:15), last Linear without ReLU.
"""
from __future__ import annotations

import numpy as np

NETWORK_SIZE = 1024          # deepsdf.py:7
LINEAR_INDICES = (0, 3, 6, 9, 12, 15, 18, 21, 24)   # positions of the Linear modules in the Sequential, deepsdf.py:13-37


def reference_like_module(hidden: int = NETWORK_SIZE, n_linear: int = 9):
    """torch module with the structure (and therefore RNG consumption and state-dict keys) of deepsdf.py:12-38."""
    import torch.nn as nn
    from torch.nn.utils.parametrizations import weight_norm
    layers = []
    dims = [3] + [hidden] * (n_linear - 1) + [1]
    for l in range(n_linear):
        layers.append(weight_norm(nn.Linear(dims[l], dims[l + 1])))
        if l < n_linear - 1:
            layers += [nn.ReLU(), nn.Dropout(0.0)]
    import torch
    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.network = nn.Sequential(*layers)
        def forward(self, coords):
            return self.network(coords)
    return M()


def seeded_state(seed: int = 0, hidden: int = NETWORK_SIZE, n_linear: int = 9):
    """State dict (numpy, fp32) of a default-initialised network under torch.manual_seed(seed):
    what `torch.manual_seed(seed); DeepSDFWithCode().state_dict()` gives for the reference class."""
    import torch
    torch.manual_seed(seed)
    m = reference_like_module(hidden, n_linear)
    return {k: v.detach().cpu().numpy().astype(np.float32) for k, v in m.state_dict().items()}


def layers_from_state(state):
    """[(g [out,1], v [out,in], bias [out])] per Linear, from the reference's state-dict keys."""
    idx = sorted({int(k.split(".")[1]) for k in state if k.startswith("network.")})
    out = []
    for i in idx:
        out.append((np.asarray(state[f"network.{i}.parametrizations.weight.original0"]),
                    np.asarray(state[f"network.{i}.parametrizations.weight.original1"]),
                    np.asarray(state[f"network.{i}.bias"])))
    return out


def effective_weights(state, dtype=np.float32):
    """W = g * v / ||v||_row (torch _weight_norm, dim=0), bias."""
    ws = []
    for g, v, b in layers_from_state(state):
        v = v.astype(dtype); g = g.astype(dtype)
        nrm = np.sqrt((v * v).sum(axis=1, keepdims=True))
        ws.append(((v * (g / nrm)).astype(dtype), b.astype(dtype)))
    return ws


def forward(state, coords, dtype=np.float32):
    """DeepSDFWithCode.forward (deepsdf.py:40-41): [n,3] -> [n,1]."""
    h = np.asarray(coords, dtype=dtype)
    ws = effective_weights(state, dtype)
    for l, (W, b) in enumerate(ws):
        h = h @ W.T + b
        if l < len(ws) - 1:
            h = np.maximum(h, 0)
    return h


def world_to_model(p_world, R, lift):
    """Inverse of the asset placement p_world = p_model @ R + lift (sim.py:46-52); R orthogonal."""
    return (np.asarray(p_world, np.float64) - np.asarray(lift, np.float64)) @ np.asarray(R, np.float64).T


def contact_force(state, p_world, R, lift, k_col, col_range, fd_eps, dtype=np.float64):
    """Obstacle contact (extension of compute_collision_penalty, sim.py:238-244):
    delta = range - sdf(p_model), f = delta^2 k n, n = grad sdf / |grad sdf| rotated to world.
    The gradient is the forward difference of step fd_eps in model space, as the CUDA path takes it."""
    pm = world_to_model(p_world, R, lift)
    s0 = forward(state, pm, dtype)[:, 0]
    g = np.zeros((len(pm), 3))
    for a in range(3):
        e = np.zeros(3); e[a] = fd_eps
        g[:, a] = (forward(state, pm + e, dtype)[:, 0] - s0) / fd_eps
    gw = g @ np.asarray(R, np.float64)          # model -> world for row vectors: p_world = p_model @ R
    nn = np.linalg.norm(gw, axis=1)
    f = np.zeros_like(gw)
    hit = (s0 < col_range) & (nn > 1e-20)
    d = col_range - s0[hit]
    f[hit] = (d * d * k_col / nn[hit])[:, None] * gw[hit]
    return s0, gw, f


def octahedron_state(radius: float, hidden: int = 256, n_linear: int = 4):
    """Hand-set weights that make the ReLU MLP compute (|x|+|y|+|z| - r)/sqrt(3) exactly:
    layer 0 rows = +-e_x, +-e_y, +-e_z; hidden layers pass the 6 units through; the last layer sums them."""
    dims = [3] + [hidden] * (n_linear - 1) + [1]
    st = {}
    for l in range(n_linear):
        o, i = dims[l + 1], dims[l]
        v = np.zeros((o, i), np.float32)
        b = np.zeros(o, np.float32)
        if l == 0:
            for a in range(3):
                v[2 * a, a] = 1.0; v[2 * a + 1, a] = -1.0
            v[6:, 0] = 1.0            # keep the norm non-zero; g = 0 silences these rows
            g = np.zeros((o, 1), np.float32); g[:6] = 1.0
        elif l < n_linear - 1:
            for u in range(6):
                v[u, u] = 1.0
            v[6:, 0] = 1.0
            g = np.zeros((o, 1), np.float32); g[:6] = 1.0
        else:
            v[0, :6] = 1.0
            g = np.full((1, 1), np.sqrt(6.0) / np.sqrt(3.0), np.float32)     # W = g v/|v| = (1/sqrt 3) on the 6 units
            b[0] = -radius / np.sqrt(3.0)
        st[f"network.{3 * l}.parametrizations.weight.original0"] = g
        st[f"network.{3 * l}.parametrizations.weight.original1"] = v
        st[f"network.{3 * l}.bias"] = b
    return st
