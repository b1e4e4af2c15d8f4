"""The three gather modes (include/mis.h, mis_set_gather_mode) against the oracle and against each other; the bounded row
capacity of the obstacle query; what an fp64 scene refuses."""
import os

import numpy as np
import pytest
import torch

from conftest import make_oracle, deformed
from meshless_inflatable_softbody_b200 import SceneConfig, Simulator, DeepSDF, scenes
from meshless_inflatable_softbody_b200.native import MisError

pytestmark = pytest.mark.gpu
FLOOR_MULT = 4.0


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fields_and_trajectory_in_every_gather_mode(sphere3k, mode):
    x0 = sphere3k
    sim, o = Simulator(x0, SceneConfig(), keep_fields=True), make_oracle(x0)
    sim.set_gather_mode(mode)
    info = sim.gather_info()
    assert info["mode"] == mode and (mode == 0 or info["max_tile"] <= info["cap_deform"])
    x = deformed(x0)
    sim.set_state(x, np.zeros_like(x))
    g = sim.fields(want=("A", "R", "F", "S", "fel"))
    ref = o.eval(x)
    o.set_order(1); rev = o.eval(x); o.set_order(0)
    for k, ko, rel in (("A", "A", 1e-6), ("R", "R", 1e-6), ("F", "F", 1e-6), ("S", "S", 5e-5), ("fel", "f", 2e-5)):
        tol = FLOOR_MULT * np.abs(rev[ko] - ref[ko]).max() + rel * np.abs(ref[ko]).max()
        assert np.abs(_np(g[k]) - ref[ko]).max() <= tol, (mode, k)
    a, b = make_oracle(x0), make_oracle(x0)
    b.set_order(1)
    a.startup(); b.startup(); sim.startup()
    a.step(100); b.step(100); sim.step(100)
    xs, vs = sim.position_velocity()
    assert np.abs(_np(xs) - a.position()).max() <= FLOOR_MULT * np.abs(a.position() - b.position()).max() + 4e-9
    assert np.abs(_np(vs) - a.velocity()).max() <= FLOOR_MULT * np.abs(a.velocity() - b.velocity()).max() + 2e-5


def test_mode_switch_mid_run_continues_the_same_trajectory(sphere3k):
    """Switching the gather kernels re-primes the frame; the state is untouched."""
    a, b = Simulator(sphere3k, SceneConfig()), Simulator(sphere3k, SceneConfig())
    a.startup(); b.startup()
    a.step(30); b.step(30)
    b.set_gather_mode(0)
    a.step(30); b.step(30)
    xa, va = a.position_velocity(); xb, vb = b.position_velocity()
    o, r = make_oracle(sphere3k), make_oracle(sphere3k)
    r.set_order(1); o.startup(); r.startup(); o.step(60); r.step(60)
    assert float((xa - xb).abs().max()) <= FLOOR_MULT * np.abs(o.position() - r.position()).max() + 4e-9
    assert float((va - vb).abs().max()) <= FLOOR_MULT * np.abs(o.velocity() - r.velocity()).max() + 2e-5


def test_lists_are_identical_whichever_build_made_them(sphere3k):
    """The tile-bitmask build and the per-thread 27-cell walk (MIS_BUILD_WALK=1) give the same exact lists, entry for entry."""
    a = Simulator(sphere3k, SceneConfig())
    off_a, nb_a = (_np(t) for t in a.neighbors())
    os.environ["MIS_BUILD_WALK"] = "1"
    try:
        import subprocess, sys, json
        code = ("import numpy as np, sys; sys.path.insert(0, %r); from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes;"
                "x0, _ = scenes.jittered_sphere(3000, seed=0, low_drop=True); s = Simulator(x0, SceneConfig()); off, nb = s.neighbors();"
                "np.save(sys.argv[1] + '_off.npy', off.cpu().numpy()); np.save(sys.argv[1] + '_nb.npy', nb.cpu().numpy())" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        base = os.path.join(os.environ.get("TMPDIR", "/tmp"), "mis_walk_lists")
        subprocess.run([sys.executable, "-c", code, base], check=True, env=dict(os.environ))
        off_b, nb_b = np.load(base + "_off.npy"), np.load(base + "_nb.npy")
    finally:
        os.environ.pop("MIS_BUILD_WALK", None)
    assert np.array_equal(off_a, off_b) and np.array_equal(nb_a, nb_b)


def test_contact_row_capacity_overflow_is_reported():
    """More candidates in the obstacle's bounding box than the chain has rows: a sticky device flag, surfaced as an error."""
    os.environ["MIS_CONTACT_ROWS"] = "384"
    try:
        cfg = SceneConfig()
        x0, _ = scenes.jittered_sphere(3000, seed=2)
        st = scenes.plateau_obstacle_state(0.05, 0.02, hidden=1024, n_linear=9)
        net = DeepSDF(st)
        sim = Simulator(x0, cfg)
        sim.set_sdf_obstacle(net, bbox_model=[-1, -1, -1, 1, 1, 1], fd_eps=1e-4)       # the box holds all 3000 particles
        sim.startup(); sim.step(2)
        with pytest.raises(MisError, match="overflow"):
            sim.contact_counts()
        nb, nc = sim.contact_counts()                                                  # the flag was cleared by the read ...
        sim.step(1)
        with pytest.raises(MisError, match="overflow"):                                # ... and is raised again by the next step
            sim.contact_counts()
    finally:
        os.environ.pop("MIS_CONTACT_ROWS", None)


def test_fp64_scene_refuses_what_it_does_not_implement():
    x0, _ = scenes.jittered_sphere(500, seed=0)
    sim = Simulator(x0.astype(np.float64), SceneConfig(), precision="f64")
    sim.startup(); sim.step(2)
    x, v = sim.position_velocity()
    assert x.dtype == torch.float64 and torch.isfinite(x).all()
    with pytest.raises(MisError, match="fp64"):
        sim.profile_step(1)
    with pytest.raises(MisError, match="fp64"):
        sim.rebuild_neighbors()
    f32 = torch.empty((len(x0), 3), dtype=torch.float32).pin_memory()
    with pytest.raises(MisError, match="fp64"):
        sim.get_state_host(f32, f32.clone())


@pytest.mark.parametrize("n,spacing,expect_tiles", [(8000, 0.4, False), (7000, 0.3, False), (4000, 0.5, True)])
def test_dense_clouds_fall_back_and_stay_exact(n, spacing, expect_tiles):
    """A 27-cell neighbourhood of more than 2816 particles does not fit the force tile (more than 4096: not even the bitmask
    build): the scene runs on the cluster kernels, the lists stay bit-equal to the oracle's, the step stays on the oracle's track."""
    x0, _ = scenes.jittered_sphere(n, seed=5, spacing=spacing)
    cfg = SceneConfig(youngs_modulus=1.5e4)            # 400-800 neighbours per particle: keep the explicit step stable
    sim, o = Simulator(x0, cfg), make_oracle(x0, cfg)
    info = sim.gather_info()
    assert (info["mode"] == 2) == expect_tiles, info
    off, nb = (_np(t) for t in sim.neighbors())
    cnt, ooff, oflat = o.neighbor_lists()
    assert np.array_equal(np.diff(off), cnt)
    rows = np.repeat(np.arange(len(x0)), cnt)
    assert np.array_equal(nb[np.lexsort((nb, rows))], oflat)
    if not expect_tiles:
        with pytest.raises(MisError, match="tile"):
            sim.set_gather_mode(2)
    b = make_oracle(x0, cfg); b.set_order(1)
    sim.startup(); o.startup(); b.startup()
    sim.step(10); o.step(10); b.step(10)
    x, v = sim.position_velocity()
    assert np.abs(_np(x) - o.position()).max() <= FLOOR_MULT * np.abs(o.position() - b.position()).max() + 4e-9
    assert np.abs(_np(v) - o.velocity()).max() <= FLOOR_MULT * np.abs(o.velocity() - b.velocity()).max() + 2e-5


def _env(name, value):
    class _E:
        def __enter__(self):
            self.old = os.environ.get(name)
            os.environ[name] = value
        def __exit__(self, *a):
            if self.old is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = self.old
    return _E()


@pytest.mark.parametrize("n", [3000, 20001])
def test_cell_expansion_equals_the_bitwise_expansion(n):
    """k_tile_expand (one CTA per cell, warp staging buffers) and k_bits_expand (MIS_BUILD_EXPAND_OLD=1) are two expansions of the
    same bitmasks: identical exact lists, and -- because the uint16 tile lists and the cluster union lists then hold the same
    entries in the same order -- bit-identical trajectories."""
    x0, _ = scenes.jittered_sphere(n, seed=3, low_drop=True)
    a = Simulator(x0, SceneConfig())
    with _env("MIS_BUILD_EXPAND_OLD", "1"):
        b = Simulator(x0, SceneConfig())
    for u, w in zip(a.neighbors(), b.neighbors()):
        assert torch.equal(u, w)
    a.startup(); b.startup()
    a.step(25); b.step(25)
    (xa, va), (xb, vb) = a.position_velocity(), b.position_velocity()
    assert torch.equal(xa, xb) and torch.equal(va, vb)
    a.rebuild_neighbors(); a.step(5); b.step(5)                    # a rebuild reproduces the same lists (Total-Lagrangian)
    (xa, va), (xb, vb) = a.position_velocity(), b.position_velocity()
    assert torch.equal(xa, xb) and torch.equal(va, vb)


def test_split_contact_integration_is_bit_identical():
    """MIS_CONTACT_SPLIT: the force gather stores the elastic force and k_integrate follows the contact chain, instead of the
    fused epilogue waiting for it.  Same operands, same operation order: the same bits."""
    cfg = SceneConfig()
    r, top = 0.05, 0.02
    st = scenes.plateau_obstacle_state(r, top, hidden=1024, n_linear=9)
    x0, _ = scenes.jittered_sphere(3000, seed=1)
    x0[:, 1] += (top - 0.003) - x0[:, 1].min()
    out = []
    for split in ("0", "1"):
        with _env("MIS_CONTACT_SPLIT", split):
            net = DeepSDF(st)
            sim = Simulator(x0, cfg)
            sim.set_sdf_obstacle(net, bbox_model=scenes.plateau_obstacle_bbox(r, top, cfg.collision_range * np.sqrt(3.0) + 5e-4), fd_eps=1e-4)
            sim.startup(); sim.step(40)
            x, v = sim.position_velocity()
            nb, nc = sim.contact_counts()
            out.append((x.clone(), v.clone(), sim.contact_force().clone(), nc))
            sim.close()
    assert out[0][3] >= 8, "the scene is not in contact"
    for u, w in zip(out[0][:3], out[1][:3]):
        assert torch.equal(u, w)
