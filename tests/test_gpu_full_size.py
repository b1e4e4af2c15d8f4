"""Oracle parity at the sizes BASELINE.json names (round-1 review: 'no oracle parity at configs[1]/[2] size').

configs[1]: ~100k particles with a DeepSDF obstacle; configs[2]: ~1M particles.  The C oracle (oracle/mis_oracle.c, CACHED mode:
bit-identical to the per-candidate FAITHFUL mode) needs ~1-2 s per force evaluation at 100k and ~20 s at 1M on the box's host
cores, so the checks are a handful of evaluations: every field at a rotated + strained state, a short trajectory, and the contact
force against the fp64 numpy restatement of the MLP.  Tolerances as everywhere: 4 x the oracle's own summation-order floor."""
import numpy as np
import pytest
import torch

from conftest import make_oracle, deformed
from meshless_inflatable_softbody_b200 import SceneConfig, Simulator, DeepSDF, scenes
from oracle import deepsdf_oracle as do

pytestmark = pytest.mark.gpu
FLOOR_MULT = 4.0


def _np(t):
    return t.detach().cpu().numpy()


def _fields_vs_oracle(x0, o, sim):
    rho, vol = o.volume()
    f = sim.fields(want=("rho", "vol"))
    assert np.abs(_np(f["rho"]) - rho).max() < 3e-6 * rho.max()
    assert np.abs(_np(f["vol"]) - vol).max() < 3e-6 * vol.max()
    x = deformed(x0)
    sim.set_state(x, np.zeros_like(x))
    g = sim.fields(want=("A", "R", "F", "S", "fel"))
    ref = o.eval(x)
    o.set_order(1); rev = o.eval(x); o.set_order(0)
    for k, ko, rel in (("A", "A", 1e-6), ("R", "R", 1e-6), ("F", "F", 1e-6), ("S", "S", 5e-5), ("fel", "f", 2e-5)):
        tol = FLOOR_MULT * np.abs(rev[ko] - ref[ko]).max() + rel * np.abs(ref[ko]).max()
        err = np.abs(_np(g[k]) - ref[ko]).max()
        assert err <= tol, (k, err, tol)


def test_configs1_size_fields_and_trajectory_match_oracle():
    """~100k particles (configs[1] size), k ~ 243: neighbour sets, every field, and 6 steps of the rollout."""
    cfg = SceneConfig()
    x0, _ = scenes.jittered_sphere(100_000, seed=0, centre=(0.0, 0.2, 0.0))
    sim, o = Simulator(x0, cfg, keep_fields=True), make_oracle(x0, cfg)
    off, nb = (_np(t) for t in sim.neighbors())
    cnt, ooff, oflat = o.neighbor_lists()
    assert np.array_equal(np.diff(off), cnt)
    rows = np.repeat(np.arange(len(x0)), cnt)
    assert np.array_equal(nb[np.lexsort((nb, rows))], oflat)
    _fields_vs_oracle(x0, o, sim)
    # a short rollout from a strained, moving state (elastic forces act from step 1)
    x = deformed(x0, seed=1, strain=0.01)
    v = np.tile(np.float32([0.0, -0.4, 0.0]), (len(x0), 1))
    steps = 6
    b = make_oracle(x0, cfg); b.set_order(1)
    for s_ in (sim, o, b):
        s_.set_state(x, v)
        s_.step(steps)
    xs, vs = sim.position_velocity()
    fx, fv = np.abs(o.position() - b.position()).max(), np.abs(o.velocity() - b.velocity()).max()
    assert np.abs(_np(xs) - o.position()).max() <= FLOOR_MULT * fx + 4e-9
    assert np.abs(_np(vs) - o.velocity()).max() <= FLOOR_MULT * fv + 2e-5
    assert np.abs(o.velocity() - v).max() > 1e-3                   # the elastic forces did act


def test_configs1_size_obstacle_contact_force_matches_oracle():
    """configs[1]: the ~100k-particle body resting on the plateau obstacle (reference 9 x 1024 MLP).  The contact force the step
    applied equals the oracle's contact law at the same positions; outside the obstacle's bounding box it is exactly zero."""
    cfg = SceneConfig()
    r, top = 0.05, 0.02
    st = scenes.plateau_obstacle_state(r, top, hidden=1024, n_linear=9)
    x0, _ = scenes.jittered_sphere(100_000, seed=0)
    x0[:, 1] += (top - 0.002) - x0[:, 1].min()                     # the lowest cap starts 2 mm inside the plateau
    margin = cfg.collision_range * np.sqrt(3.0) + 5e-4
    bbox = scenes.plateau_obstacle_bbox(r, top, margin)
    sim = Simulator(x0, cfg)
    sim.set_sdf_obstacle(DeepSDF(st), bbox_model=bbox, fd_eps=1e-4)
    sim.startup(); sim.step(4)
    x, v = sim.position_velocity()
    f = _np(sim.contact_force())
    xn = _np(x).astype(np.float64)
    lo, hi = np.asarray(bbox[:3]), np.asarray(bbox[3:])
    inside = np.all((xn >= lo) & (xn <= hi), axis=1)
    assert 64 <= inside.sum() < 30000
    assert not f[~inside].any()
    s0, gw, fo = do.contact_force(st, xn[inside], np.eye(3), np.zeros(3), cfg.collision_penalty_stiffness, cfg.collision_range, 1e-4)
    band = s0 < cfg.collision_range
    sure = np.abs(s0 - cfg.collision_range) > 2e-7
    assert band.sum() >= 32, "the scene is not in contact"
    fi = f[inside]
    assert np.array_equal((np.abs(fi).sum(1) > 0)[sure], band[sure])
    scale = np.abs(fo).max()
    assert np.abs(fi - fo)[sure].max() <= 2e-3 * scale + 1e-9
    nb, nc = sim.contact_counts()
    assert nb >= nc >= int(band[sure].sum()) - 2
    assert torch.isfinite(x).all() and torch.isfinite(v).all()


def test_configs2_size_fields_match_oracle():
    """~1M particles (configs[2] size): volumes and every field of the step at a rotated + strained state."""
    cfg = SceneConfig()
    x0, _ = scenes.jittered_sphere(1_000_000, seed=0, centre=(0.0, 0.3, 0.0))
    sim, o = Simulator(x0, cfg, keep_fields=True), make_oracle(x0, cfg)
    info = sim.neighbor_info()
    cnt, ooff, oflat = o.neighbor_lists()
    assert info.total_pairs == int(cnt.sum())
    off, nb = sim.neighbors()
    assert np.array_equal(_np(off[1:] - off[:-1]), cnt)
    _fields_vs_oracle(x0, o, sim)
