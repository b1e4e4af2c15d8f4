"""DeepSDF on the tcgen05 GEMM chain (through the C-ABI) against the oracle and the reference's golden outputs.

Tolerance: the hidden layers run as 3xTF32 with fp32 accumulation (error ~2^-21 per product); the bound is
stated against the fp64 oracle: |sdf - sdf64| <= 4e-6 * max|sdf64| + 1e-7, the same order as the reference's
own fp32-vs-fp64 difference on these inputs (checked in tests/test_deepsdf_oracle.py with 2e-6).
"""
import os

import numpy as np
import pytest
import torch

from oracle import deepsdf_oracle as do
from meshless_inflatable_softbody_b200 import SceneConfig, scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deepsdf_seed0.npz")


def _net(state):
    from meshless_inflatable_softbody_b200 import DeepSDF
    return DeepSDF(state)


@pytest.fixture(scope="module")
def seeded():
    st = do.seeded_state(0)
    return st, _net(st)


def test_golden_outputs_of_the_reference(seeded):
    st, net = seeded
    g = np.load(GOLD)
    got = net(g["points"]).cpu().numpy()
    assert got.shape == (256, 1) and got.dtype == np.float32
    scale = np.abs(g["sdf64"]).max()
    assert np.abs(got - g["sdf64"]).max() <= 4e-6 * scale + 1e-7, np.abs(got - g["sdf64"]).max()
    assert np.abs(got - g["sdf"]).max() <= 6e-6 * scale + 2e-7          # vs the reference's own fp32 outputs


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 70_000])
def test_ragged_sizes_against_oracle(seeded, n):
    st, net = seeded
    rng = np.random.default_rng(n)
    p = rng.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32)
    got = net(p).cpu().numpy()[:, 0]
    sub = np.unique(np.concatenate([np.arange(min(n, 64)), np.arange(max(0, n - 64), n), rng.integers(0, n, 128)]))
    want = do.forward(st, p[sub], np.float64)[:, 0]
    assert np.isfinite(got).all()
    assert np.abs(got[sub] - want).max() <= 4e-6 * np.abs(want).max() + 1e-7


@pytest.mark.parametrize("path,n", [(1, 1), (1, 129), (1, 1000), (1, 1024 + 77), (1, 3000), (2, 1), (2, 129), (2, 1000),
                                    (3, 1), (3, 129), (3, 300), (3, 1000), (3, 3000)])
def test_both_hidden_layer_kernels_at_any_row_count(seeded, path, n):
    """The split-K cluster kernels (few rows: per-step contact; 1 = one launch per layer, 3 = every hidden layer in one cooperative
    launch with a device-wide barrier between layers) and the persistent big-tile kernel (bulk) are all correct for any row count,
    agree bit for bit where they share the reduction order (1 vs 3), and are bit-reproducible."""
    st, net = seeded
    rng = np.random.default_rng(100 + n)
    p = rng.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32)
    net.set_gemm_path(path)
    try:
        got = net(p).cpu().numpy()[:, 0]
        again = net(p).cpu().numpy()[:, 0]
    finally:
        net.set_gemm_path(0)
    sub = np.unique(np.concatenate([np.arange(min(n, 64)), np.arange(max(0, n - 64), n), rng.integers(0, n, 128)]))
    want = do.forward(st, p[sub], np.float64)[:, 0]
    assert np.isfinite(got).all()
    assert np.abs(got[sub] - want).max() <= 4e-6 * np.abs(want).max() + 1e-7
    assert np.array_equal(got, again)                       # fixed reduction order: bit-reproducible
    if path == 3:
        net.set_gemm_path(1)
        try:
            per_layer = net(p).cpu().numpy()[:, 0]
        finally:
            net.set_gemm_path(0)
        assert np.array_equal(got, per_layer)               # same K split, same fixed-order reduction


def test_small_network_and_octahedron_exactness():
    st = do.octahedron_state(0.03, hidden=256, n_linear=4)
    net = _net(st)
    rng = np.random.default_rng(5)
    p = rng.uniform(-0.08, 0.08, size=(3000, 3)).astype(np.float32)
    want = (np.abs(p.astype(np.float64)).sum(1) - 0.03) / np.sqrt(3.0)
    got, grad = net.query(p, grad=True, fd_eps=1e-4)
    assert np.abs(got.cpu().numpy() - want).max() < 8e-8     # a few fp32 ulps at 0.1
    # gradient = sign(p)/sqrt(3) away from the kinks
    away = (np.abs(p) > 2e-4).all(1)
    gw = np.sign(p[away]) / np.sqrt(3.0)
    assert np.abs(grad.cpu().numpy()[away] - gw).max() < 2e-3


def test_world_to_model_transform_and_fd_gradient(seeded):
    from meshless_inflatable_softbody_b200.deepsdf import world_to_model_xform, ASSET_R, ASSET_LIFT
    st, net = seeded
    rng = np.random.default_rng(2)
    pm = rng.uniform(-0.5, 0.5, size=(500, 3))
    pw = (pm @ ASSET_R + ASSET_LIFT).astype(np.float32)                  # sim.py:52
    s_w, g_w = net.query(pw, xform=world_to_model_xform(), grad=True, fd_eps=1e-3)
    s0, gw, _ = do.contact_force(st, pw.astype(np.float64), ASSET_R, ASSET_LIFT, 3e5, 1e-4, 1e-3)
    assert np.abs(s_w.cpu().numpy() - s0).max() <= 4e-6 * np.abs(s0).max() + 1e-7
    # forward difference of fp32-accurate values: error ~ 2 * 4e-6 * |sdf| / eps
    assert np.abs(g_w.cpu().numpy() - gw).max() <= 2 * 5e-6 * np.abs(s0).max() / 1e-3 + 1e-4


def test_design_field_as_in_the_reference(seeded):
    """sim.py:100-101: x = sdf(points); x[:out_num] = clip(x[:out_num], 1, None)."""
    st, net = seeded
    x0, out_num = scenes.jittered_sphere(2000, seed=0)
    pm = (x0 - np.float32([0, 0.07, 0])).astype(np.float32)
    x = net.design_field(pm, out_num).cpu().numpy()
    ref = do.forward(st, pm, np.float64)[:, 0]
    ref[:out_num] = np.clip(ref[:out_num], 1.0, None)
    assert np.abs(x - ref).max() <= 4e-6 * np.abs(ref).max() + 1e-7


def test_plane_obstacle_equals_ground_penalty():
    """An SDF obstacle that is the plane y = 0 must reproduce the reference's ground contact (sim.py:238-244):
    the trajectory with (ground off, plane obstacle on) equals the trajectory with the built-in ground penalty."""
    from meshless_inflatable_softbody_b200 import Simulator
    hidden, n_linear = 256, 3
    st = do.octahedron_state(0.0, hidden=hidden, n_linear=n_linear)
    g0 = np.zeros((hidden, 1), np.float32); g0[:2] = 1.0
    v0 = np.zeros((hidden, 3), np.float32); v0[0, 1] = 1.0; v0[1, 1] = -1.0; v0[2:, 0] = 1.0
    st["network.0.parametrizations.weight.original0"] = g0
    st["network.0.parametrizations.weight.original1"] = v0
    vl = np.zeros((1, hidden), np.float32); vl[0, 0] = 1.0; vl[0, 1] = -1.0
    st[f"network.{3 * (n_linear - 1)}.parametrizations.weight.original1"] = vl
    st[f"network.{3 * (n_linear - 1)}.parametrizations.weight.original0"] = np.full((1, 1), np.sqrt(2.0), np.float32)
    st[f"network.{3 * (n_linear - 1)}.bias"] = np.zeros(1, np.float32)
    net = _net(st)
    x0, _ = scenes.jittered_sphere(1500, seed=0, low_drop=True)
    a = Simulator(x0, SceneConfig())
    b = Simulator(x0, SceneConfig(ground_contact=False))
    b.set_sdf_obstacle(net, bbox_model=[-1, -1, -1, 1, 2e-4, 1], fd_eps=1e-3)
    c = Simulator(x0, SceneConfig(ground_contact=False))                   # control: no contact at all
    a.startup(); b.startup(); c.startup()
    a.step(300); b.step(300); c.step(300)
    xa, va = a.position_velocity(); xb, vb = b.position_velocity(); xc, vc = c.position_velocity()
    assert b.contact_count() > 0
    assert (va - vc).abs().max().item() > 0.05                            # the ground penalty has acted by now
    # sdf = y is exact in the MLP up to fp32 rounding; the two runs differ by rounding of the contact force only,
    # amplified like any fp32 reordering (~1e-7 / ~1e-3 after a few hundred steps)
    assert (xa - xb).abs().max().item() < 3e-7 and (va - vb).abs().max().item() < 3e-3


def test_octahedron_obstacle_contact_force_matches_oracle():
    """Drop a sphere on an octahedron encoded in the reference architecture (9 layers); the contact force the step
    applied at frame k must equal the oracle's contact law at the same positions, and the body must decelerate."""
    from meshless_inflatable_softbody_b200 import Simulator
    r_oct = 0.012
    st = do.octahedron_state(r_oct, hidden=1024, n_linear=9)
    net = _net(st)
    cfg = SceneConfig()
    x0, _ = scenes.jittered_sphere(3000, seed=0)
    x0[:, 1] += (r_oct + 0.0004) - x0[:, 1].min()                        # lowest particle 0.4 mm above the tip
    m = cfg.collision_range + 1e-3
    sim = Simulator(x0, cfg)
    sim.set_sdf_obstacle(net, bbox_model=[-r_oct - m, -r_oct - m, -r_oct - m, r_oct + m, r_oct + m, r_oct + m], fd_eps=1e-4)
    sim.startup()
    hit = False
    for k in range(12):
        sim.step(10)
        x, v = sim.position_velocity()
        f = sim.contact_force().cpu().numpy()
        xn = x.cpu().numpy().astype(np.float64)
        s0, gw, fo = do.contact_force(st, xn, np.eye(3), np.zeros(3), cfg.collision_penalty_stiffness, cfg.collision_range, 1e-4)
        band = s0 < cfg.collision_range
        # particles whose value sits within fp32 rounding of the band edge may fall on either side
        sure = np.abs(s0 - cfg.collision_range) > 2e-7
        assert np.array_equal((np.abs(f).sum(1) > 0)[sure], band[sure])
        if band.any():
            hit = True
            scale = np.abs(fo).max()
            # delta = range - sdf is known to ~1e-8 absolute (fp32 sdf at 0.01): relative force error 2 * 1e-8 / delta for the deepest
            assert np.abs(f - fo)[sure].max() <= 2e-3 * scale + 1e-9, (np.abs(f - fo).max(), scale)
        nb, nc = sim.contact_counts()
        assert nb >= nc >= int(band[sure].sum()) - 2
    assert hit
    assert torch.isfinite(x).all()
