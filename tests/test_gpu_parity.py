"""Parity of the CUDA path (through the C-ABI of include/mis.h) with the CPU oracle.

Integer structures (cell indices, cell-sorted order, neighbour sets) must be bit-exact.
Floating-point state is compared against a tolerance stated per test as a multiple of the
oracle's own fp32 summation-order noise floor (the same oracle run with the candidate walk
reversed), measured on the same fixture and number of steps: an fp32 sum of ~200 terms is
not associative and stiffness 3e7 Pa amplifies position ulps, so no two correct fp32
implementations agree closer than that.
"""
import numpy as np
import pytest
import torch

from conftest import make_oracle, deformed
from meshless_inflatable_softbody_b200 import SceneConfig, scenes

pytestmark = pytest.mark.gpu

FLOOR_MULT = 4.0      # tolerance = FLOOR_MULT x measured reorder floor (+ a tiny absolute term)


def _sim(x0, cfg=None, **kw):
    from meshless_inflatable_softbody_b200 import Simulator
    return Simulator(x0, cfg or SceneConfig(), **kw)


def _np(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ integer structures: bit-exact
@pytest.mark.parametrize("n,spacing,seed", [(800, 0.5, 0), (3000, 0.5, 1), (2000, 0.8, 2), (5000, 0.65, 3)])
def test_cells_and_neighbours_bit_exact(n, spacing, seed):
    x0, _ = scenes.jittered_sphere(n, seed=seed, spacing=spacing)
    sim, o = _sim(x0), make_oracle(x0)
    ci, cc, pm = (_np(t) for t in sim.cells())
    oc, occ, oids = o.grid_arrays()
    assert np.array_equal(ci, oc)                 # wp.HashGrid linear cell index
    assert np.array_equal(cc, occ)                # int(p / cell_width)
    assert np.array_equal(np.sort(pm), np.arange(len(x0)))
    s, e = (_np(t) for t in sim.cell_ranges())
    assert int((e - s).sum()) == len(x0)
    for a, b in zip(s, e):                        # a cell is one contiguous slot range
        assert b <= a or len(np.unique(cc[pm[a:b]], axis=0)) == 1
    # hash_grid_point_id order (stable sort by cell index) follows from the bit-exact cell indices
    assert np.array_equal(np.argsort(ci, kind="stable"), oids)
    off, nb = (_np(t) for t in sim.neighbors())
    cnt, ooff, oflat = o.neighbor_lists()
    assert np.array_equal(np.diff(off), cnt)
    rows = np.repeat(np.arange(len(x0)), cnt)
    got = np.lexsort((nb, rows))
    assert np.array_equal(nb[got], oflat)         # same sets, row by row
    info = sim.neighbor_info()
    assert info.total_pairs == int(cnt.sum()) and info.max_neighbors == int(cnt.max())


@pytest.mark.parametrize("pts", [
    [[0.0, 0.07, 0.0]],                                            # single particle: no neighbours
    [[0.0, 0.07, 0.0], [0.005, 0.07, 0.0]],                        # one pair
    [[0.0, 0.07, 0.0], [0.05, 0.07, 0.0], [0.0, 0.2, 0.0]],        # all isolated
    [[-0.001, 0.07, 0.0], [0.001, 0.07, 0.0], [0.0139, 0.07, 0.0], [0.0141, 0.07, 0.0]],   # straddles x = 0 and the 2h limit
])
def test_ragged_tiny_inputs(pts):
    x0 = np.asarray(pts, np.float32)
    sim, o = _sim(x0), make_oracle(x0)
    off, nb = (_np(t) for t in sim.neighbors())
    cnt, ooff, oflat = o.neighbor_lists()
    assert np.array_equal(np.diff(off), cnt)
    for i in range(len(x0)):
        assert np.array_equal(np.sort(nb[off[i]:off[i + 1]]), oflat[ooff[i]:ooff[i + 1]])
    sim.startup(); o.startup()
    sim.step(5); o.step(5)
    x, v = sim.position_velocity()
    assert np.all(np.isfinite(_np(x))) and np.all(np.isfinite(_np(v)))
    assert np.allclose(_np(x), o.position(), atol=1e-7) and np.allclose(_np(v), o.velocity(), atol=1e-4)


def test_rebuild_is_idempotent(sphere3k):
    sim = _sim(sphere3k)
    before = [_np(t).copy() for t in sim.cells()] + [_np(t).copy() for t in sim.neighbors()]
    sim.rebuild_neighbors(); sim.rebuild_neighbors()
    after = [_np(t) for t in sim.cells()] + [_np(t) for t in sim.neighbors()]
    for a, b in zip(before, after):
        assert np.array_equal(a, b)


# ------------------------------------------------------------------ per-particle fields
@pytest.mark.parametrize("C,G", [(1, 8), (2, 8), (2, 16), (2, 32), (4, 8), (4, 16), (4, 32), (1, 32)])
def test_fields_match_oracle(sphere3k, C, G):
    x0 = sphere3k
    sim, o = _sim(x0, cluster_size=C, lanes_per_particle=G, keep_fields=True), make_oracle(x0)
    f = sim.fields(want=("rho", "vol"))
    rho, vol = o.volume()
    assert np.abs(_np(f["rho"]) - rho).max() < 3e-6 * rho.max()
    assert np.abs(_np(f["vol"]) - vol).max() < 3e-6 * vol.max()
    x = deformed(x0)
    sim.set_state(x, np.zeros_like(x))
    g = sim.fields(want=("A", "R", "F", "S", "fel"))
    ref = o.eval(x)
    o.set_order(1); ref_rev = o.eval(x); o.set_order(0)
    for k, ko, rel_min in (("A", "A", 1e-6), ("R", "R", 1e-6), ("F", "F", 1e-6), ("S", "S", 5e-5), ("fel", "f", 2e-5)):
        floor = np.abs(ref_rev[ko] - ref[ko]).max()
        tol = FLOOR_MULT * floor + rel_min * np.abs(ref[ko]).max()
        err = np.abs(_np(g[k]) - ref[ko]).max()
        assert err <= tol, (k, err, tol, floor)
    fe = _np(sim.eval_forces(x))
    assert np.abs(fe - ref["f"]).max() <= FLOOR_MULT * np.abs(ref_rev["f"] - ref["f"]).max() + 2e-5 * np.abs(ref["f"]).max()


def test_rest_state_and_rigid_motion(sphere3k):
    x0 = sphere3k
    sim = _sim(x0)
    c = x0.mean(0)
    f_ref = _np(sim.eval_forces(((x0 - c) * 1.01 + c).astype(np.float32)))   # 1 % stretch for scale
    f_rest = _np(sim.eval_forces(x0))
    assert np.abs(f_rest).max() < 2e-3 * np.abs(f_ref).max()
    x = deformed(x0, angle=0.7, strain=0.0, noise=0.0) + np.float32([0.01, 0.02, -0.03])
    sim.set_state(x, np.zeros_like(x))
    g = sim.fields(want=("R", "F", "fel"))
    th = 0.7
    Q = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    assert np.abs(_np(g["R"]) - Q).max() < 5e-6
    assert np.abs(_np(g["F"]) - np.eye(3)).max() < 2e-5
    assert np.abs(_np(g["fel"])).max() < 1e-2 * np.abs(f_ref).max()
    assert np.abs(np.linalg.det(_np(g["R"]).astype(np.float64)) - 1).max() < 1e-5


# ------------------------------------------------------------------ trajectories
def _floor_after(x0, cfg, steps, variant="warp", **flags):
    a, b = make_oracle(x0, cfg, variant=variant, **flags), make_oracle(x0, cfg, variant=variant, **flags)
    b.set_order(1)
    a.startup(cfg.initial_velocity); b.startup(cfg.initial_velocity)
    a.step(steps); b.step(steps)
    return a, np.abs(a.position() - b.position()).max(), np.abs(a.velocity() - b.velocity()).max()


@pytest.mark.parametrize("steps", [1, 100, 300])
def test_trajectory_within_noise_floor(sphere3k, steps):
    """Tolerance: |dx| <= 4 x floor + 4e-9 (half an ulp at 0.07 m), |dv| <= 4 x floor + 2e-5 m/s."""
    cfg = SceneConfig()
    o, fx, fv = _floor_after(sphere3k, cfg, steps)
    sim = _sim(sphere3k, cfg)
    sim.startup(); sim.step(steps)
    x, v = sim.position_velocity()
    dx, dv = np.abs(_np(x) - o.position()).max(), np.abs(_np(v) - o.velocity()).max()
    assert dx <= FLOOR_MULT * fx + 4e-9, (dx, fx)
    assert dv <= FLOOR_MULT * fv + 2e-5, (dv, fv)


def test_trajectory_1000_steps_config1():
    """BASELINE config 1 shape (sphere, reference defaults, 1000 steps, impact included) at n ~ 1500."""
    cfg = SceneConfig()
    x0, _ = scenes.jittered_sphere(1500, seed=0, low_drop=True)
    o, fx, fv = _floor_after(x0, cfg, 1000)
    sim = _sim(x0, cfg)
    sim.startup(); sim.step(1000)
    x, v = sim.position_velocity()
    assert o.velocity()[:, 1].mean() > -0.2         # the drop has hit the ground plane and is rebounding
    assert np.abs(_np(x) - o.position()).max() <= FLOOR_MULT * fx + 4e-9
    assert np.abs(_np(v) - o.velocity()).max() <= FLOOR_MULT * fv + 2e-5


def test_config0_full_size_against_committed_oracle_trajectory():
    """BASELINE configs[0] at its full size: ~10k-particle sphere, reference defaults, 1000 steps (impact and rebound included),
    against the oracle trajectory committed under tests/golden/ (made by tests/golden/make_config0_golden.py; ~7 CPU-minutes,
    which is why it is a fixture).  Tolerance as everywhere: 4 x the oracle's own reorder floor at the same checkpoint."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config0_n10k.npz"))
    cfg = SceneConfig()
    x0, _ = scenes.jittered_sphere(10000, seed=0, low_drop=True)
    assert len(x0) == int(gold["n"]) and float(x0.astype(np.float64).sum()) == float(gold["x0_checksum"])
    sim = _sim(x0, cfg)
    sim.startup()
    done = 0
    for cp in (100, 500, 1000):
        sim.step(cp - done); done = cp
        x, v = sim.position_velocity()
        dx, dv = np.abs(_np(x) - gold[f"x_{cp}"]).max(), np.abs(_np(v) - gold[f"v_{cp}"]).max()
        assert dx <= FLOOR_MULT * float(gold[f"floor_x_{cp}"]) + 4e-9, (cp, dx, float(gold[f"floor_x_{cp}"]))
        assert dv <= FLOOR_MULT * float(gold[f"floor_v_{cp}"]) + 2e-5, (cp, dv, float(gold[f"floor_v_{cp}"]))
    assert gold["v_1000"][:, 1].mean() > -0.2        # the golden run did hit the ground plane


def test_ballistic_is_bit_exact():
    """E = 0 removes the only reordered sums: integration + ground penalty must match bit for bit."""
    x0, _ = scenes.jittered_sphere(1000, seed=3, low_drop=True)
    cfg = SceneConfig(youngs_modulus=0.0)
    sim, o = _sim(x0, cfg), make_oracle(x0, cfg)
    sim.startup(); o.startup()
    sim.step(400); o.step(400)
    x, v = sim.position_velocity()
    assert (o.velocity()[:, 1] > 0).any()           # penalty branch exercised: some particles bounced
    assert np.array_equal(_np(x), o.position()) and np.array_equal(_np(v), o.velocity())


def test_graph_chunks_equal_single_launches(sphere3k):
    a, b = _sim(sphere3k, graph_steps=16), _sim(sphere3k, graph_steps=-1)
    a.startup(); b.startup()
    a.step(70)                                      # 4 graph chunks of 16 + 6 direct steps
    for _ in range(70):
        b.step(1)
    xa, va = a.position_velocity(); xb, vb = b.position_velocity()
    assert torch.equal(xa, xb) and torch.equal(va, vb)


def test_set_state_resumes_identically(sphere3k):
    a, b = _sim(sphere3k), _sim(sphere3k)
    a.startup(); a.step(40)
    x, v = a.position_velocity()
    a.step(25)
    b.set_state(x, v, frame=40); b.step(25)
    xa, va = a.position_velocity(); xb, vb = b.position_velocity()
    assert torch.equal(xa, xb) and torch.equal(va, vb)


def test_dirichlet_and_external_force(sphere800):
    cfg = SceneConfig()
    x0 = sphere800
    sim, o = _sim(x0, cfg), make_oracle(x0, cfg)
    top = np.nonzero(x0[:, 1] > np.percentile(x0[:, 1], 90))[0]
    free = np.ones((len(x0), 3), np.float32); free[top] = 0.0
    fext = np.tile(np.float32(cfg.external_force), (len(x0), 1))
    side = np.nonzero(x0[:, 0] < np.percentile(x0[:, 0], 10))[0]
    fext[side] = [2e-3, 0.0, 0.0]
    sim.set_dirichlet(torch.as_tensor(top), [0.0, 0.0, 0.0]); o.set_free_points(free)
    sim.set_external_force(torch.as_tensor(side), [2e-3, 0.0, 0.0]); o.set_external_forces(fext)
    sim.startup(); o.startup()
    sim.step(50); o.step(50)
    x, v = sim.position_velocity()
    assert np.array_equal(_np(x)[top], x0[top])                       # pinned particles never move
    b = make_oracle(x0, cfg); b.set_order(1); b.set_free_points(free); b.set_external_forces(fext); b.startup(); b.step(50)
    fx, fv = np.abs(o.position() - b.position()).max(), np.abs(o.velocity() - b.velocity()).max()
    assert np.abs(_np(x) - o.position()).max() <= FLOOR_MULT * fx + 4e-9
    assert np.abs(_np(v) - o.velocity()).max() <= FLOOR_MULT * fv + 2e-5


def test_design_field_soft_shell(sphere3k):
    """Per-particle design x (sim.py:98-110): soft shell ratio ~ 1, stiff core."""
    cfg = SceneConfig()
    x0 = sphere3k
    design = np.where(scenes.shell_mask(x0, cfg.h), 1.0, -1.0).astype(np.float32)
    sim, o = _sim(x0, cfg), make_oracle(x0, cfg)
    sim.set_design(design); o.set_design(design)
    x = deformed(x0, strain=0.05)
    ref = o.eval(x)["f"]
    o.set_order(1); rev = o.eval(x)["f"]
    got = _np(sim.eval_forces(x))
    assert np.abs(got - ref).max() <= FLOOR_MULT * np.abs(rev - ref).max() + 2e-5 * np.abs(ref).max()


# ------------------------------------------------------------------ sim_taichi.py variant switches
def test_taichi_variant_flags(sphere800):
    cfg = SceneConfig.taichi().with_(h=0.007, time_step=5e-5, mass=1e-4, youngs_modulus=1.5e5,
                                     external_force=(0.0, 0.0, -1e-3), design_x=-1.0)
    x0 = sphere800
    o, fx, fv = _floor_after(x0, cfg, 60, variant="taichi")
    sim = _sim(x0, cfg)
    sim.startup(cfg.initial_velocity); sim.step(60)
    x, v = sim.position_velocity()
    assert np.abs(_np(x) - o.position()).max() <= FLOOR_MULT * fx + 4e-9
    assert np.abs(_np(v) - o.velocity()).max() <= FLOOR_MULT * fv + 2e-5
    # the symmetric pair force conserves linear momentum (no external force, no contact)
    cfg0 = cfg.with_(external_force=(0.0, 0.0, 0.0))
    sim0 = _sim(x0, cfg0)
    f = _np(sim0.eval_forces(deformed(x0, strain=0.03))).astype(np.float64)
    assert np.abs(f.sum(0)).max() < 1e-4 * np.abs(f).sum(0).max()


# ------------------------------------------------------------------ full-size properties (BASELINE config 2 size)
def test_full_size_properties():
    x0, _ = scenes.jittered_sphere(100_000, seed=0, centre=(0.0, 0.2, 0.0))   # radius 0.1 m: clear of the ground
    n = len(x0)
    sim = _sim(x0)
    info = sim.neighbor_info()
    off, nb = sim.neighbors()
    # symmetry of the neighbour relation: (i,j) listed <=> (j,i) listed
    rows = torch.repeat_interleave(torch.arange(n, device=nb.device), (off[1:] - off[:-1]))
    a = torch.sort(rows * n + nb.long()).values
    b = torch.sort(nb.long() * n + rows).values
    assert torch.equal(a, b)
    assert 150 < info.total_pairs / n < 260
    # rigid motion => no force; 1 % stretch => radial restoring force
    c = x0.mean(0)
    f_stretch = _np(sim.eval_forces(((x0 - c) * 1.01 + c).astype(np.float32)))
    x_rot = deformed(x0, angle=1.1, strain=0.0, noise=0.0)
    f_rot = _np(sim.eval_forces(x_rot))
    assert np.abs(f_rot).max() < 1e-2 * np.abs(f_stretch).max()
    # free fall before impact: centre of mass follows the ballistic law, state stays finite
    cfg = SceneConfig()
    sim.startup(); sim.step(200)
    x, v = sim.position_velocity()
    assert torch.isfinite(x).all() and torch.isfinite(v).all()
    t = 200 * cfg.time_step
    com = _np(x).astype(np.float64).mean(0) - x0.astype(np.float64).mean(0)
    assert abs(com[1] - (-0.4 * t + 0.5 * (-1e-3 / 1e-4) * t * t)) < 2e-6
    assert np.abs(com[[0, 2]]).max() < 2e-6


def test_export_targets_format(tmp_path, sphere800):
    sim = _sim(sphere800)
    sim.startup()
    sim.export_targets(str(tmp_path), every=5, count=3)
    for i in (1, 2, 3):
        p = np.load(tmp_path / f"position_{i}.npy"); v = np.load(tmp_path / f"velocity_{i}.npy")
        assert p.shape == (len(sphere800), 3) and p.dtype == np.float32 and v.shape == p.shape
    assert sim.frame == 15


def test_force_change_takes_the_cheap_path_and_equals_a_full_reprime(sphere3k):
    """Changing only the external force between steps (sim.py:279-283) redoes force_1 / part_1 from the stored elastic
    force; the result must equal a full re-evaluation (forced here by re-sending the design field)."""
    a, b = _sim(sphere3k), _sim(sphere3k)
    a.startup(); b.startup()
    a.step(20); b.step(20)
    f = np.tile(np.float32([3e-4, -1e-3, 0.0]), (len(sphere3k), 1))
    la = a.launch_count
    a.set_external_forces(f); a.step(1)
    cheap = a.launch_count - la
    lb = b.launch_count
    b.set_external_forces(f); b.set_design(SceneConfig().design_x); b.step(1)
    full = b.launch_count - lb
    assert cheap < full
    a.step(15); b.step(15)
    xa, va = a.position_velocity(); xb, vb = b.position_velocity()
    assert torch.equal(xa, xb) and torch.equal(va, vb)
