"""fp64 scenes (sim_taichi.py runs in ti.f64, options.py:3) and the reverse pass of the rollout (sim.py:346-372).

* The Taichi prototype's own scene in DOUBLE on the GPU against the fixture produced by executing sim_taichi.py
  (tests/golden/sim_taichi_n171.npz): positions to 1e-12, velocities to 1e-10.
* The sim.py path in double against the fp64 fixtures of the executed sim.py (fields and trajectories).
* mis_rollout_grad: d loss / d design against central finite differences of the same fp64 forward pass -- the reference's own
  self-check (grad_check, sim.py:418-436) -- and the fp32 gradient against the fp64 one.
"""
import numpy as np
import pytest
import torch

from meshless_inflatable_softbody_b200 import SceneConfig, Simulator, scenes
from test_simpy_golden import fields_fixture, trajectory_fixtures, taichi_fixture

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def test_taichi_prototype_in_double_on_the_gpu():
    t = taichi_fixture()
    cfg = SceneConfig.taichi()
    sim = Simulator(t["x0"], cfg, precision="f64", apply_defaults=False)
    sim.set_youngs_modulus(1e5); sim.set_poisson_ratio(0.4); sim.set_mass(1e-2)          # sim_taichi.py:326-328
    sim.set_design(-10.0)                                                                  # sim_taichi.py:75
    sim.set_external_forces(t["external_forces"])
    free = np.ones((len(t["x0"]), 3)); free[t["edge"]] = 0.0
    assert np.array_equal(free, t["free_points"])
    sim.set_dirichlet(torch.as_tensor(t["edge"]), [0.0, 0.0, 0.0])
    f = sim.fields(want=("rho", "vol"))
    assert f["rho"].dtype == torch.float64
    assert np.abs(_np(f["rho"]) - t["rho_i"]).max() <= 1e-13 * t["rho_i"].max()
    assert np.abs(_np(f["vol"]) - t["volume_i"]).max() <= 1e-13 * t["volume_i"].max()
    sim.startup((0.0, 0.0, 0.0))
    done = 0
    for fr in [int(k) for k in t["save_frames"]]:
        if fr == 0:
            continue
        sim.step(fr - done); done = fr
        x, v = sim.position_velocity()
        assert np.abs(_np(x) - t[f"position_{fr}"]).max() <= 1e-12, fr
        assert np.abs(_np(v) - t[f"velocity_{fr}"]).max() <= 1e-10, fr
    g = sim.fields(want=("F", "S", "fel"))
    assert np.abs(_np(g["F"]) - t[f"def_grad_{done}"]).max() <= 1e-12
    assert np.abs(_np(g["S"]) - t[f"sigma_{done}"]).max() <= 1e-9 * np.abs(t[f"sigma_{done}"]).max()
    assert np.abs(_np(g["fel"]) - t[f"elastic_forces_{done}"]).max() <= 1e-10 * np.abs(t[f"elastic_forces_{done}"]).max()
    assert np.array_equal(_np(x)[t["edge"]], t["x0"][t["edge"]])                           # Dirichlet particles never move


def test_sim_py_path_in_double_against_the_fp64_fixtures():
    g = fields_fixture()
    sim = Simulator(g["x0"], SceneConfig(), precision="f64")
    f = sim.fields(want=("rho", "vol"))
    assert np.abs(_np(f["rho"]) - g["f64_rho"]).max() <= 1e-12 * g["f64_rho"].max()
    sim.set_state(g["xdef"].astype(np.float64), np.zeros_like(g["xdef"], dtype=np.float64))
    got = sim.fields(want=("A", "R", "F", "S", "fel"))
    for k, name in (("A", "A_pq"), ("R", "R"), ("F", "def_grad"), ("S", "S"), ("fel", "elastic_forces")):
        ref = g[f"f64_{name}"]
        assert np.abs(_np(got[k]) - ref).max() <= 1e-9 * np.abs(ref).max(), name
    for tr in trajectory_fixtures():
        sim = Simulator(tr["x0"], SceneConfig(), precision="f64")
        sim.startup()
        done = 0
        for fr in [int(k) for k in tr["save_frames"]]:
            sim.step(fr - done); done = fr
            x, v = sim.position_velocity()
            # two fp64 evaluations (Jacobi polar here, LAPACK SVD in the fixture) separate at ~1e-16 per step and the impact amplifies it
            assert np.abs(_np(x) - tr[f"f64_position_{fr}"]).max() <= 1e-11, (len(tr["x0"]), fr)
            assert np.abs(_np(v) - tr[f"f64_velocity_{fr}"]).max() <= 1e-7, (len(tr["x0"]), fr)


def _grad_scene(n=300, seed=4):
    x0, _ = scenes.jittered_sphere(n, seed=seed, low_drop=True)
    x0 = x0.astype(np.float64)
    x0[:, 1] += 0.0003 - x0[:, 1].min()                  # ground impact from step ~10: the penalty is inside the differentiated path
    x0 = x0.astype(np.float32)
    rng = np.random.default_rng(seed)
    design = rng.uniform(-0.4, 0.4, len(x0))
    return x0, design


def _targets(x0, frames, n_targets):
    """Trajectory of the stiff default design (x = -1) at the target frames: what `--set_target` would have written (sim.py:363-369)."""
    sim = Simulator(x0, SceneConfig(), precision="f64")
    sim.startup()
    out = []
    every = frames // n_targets
    for _ in range(n_targets):
        sim.step(every)
        x, v = sim.position_velocity()
        out.append((_np(x).astype(np.float32), _np(v).astype(np.float32)))
    return out


def test_design_gradient_against_finite_differences_in_double():
    """grad_check of sim.py:418-436: central differences of the loss against the reverse pass, here for several particles."""
    x0, design = _grad_scene()
    frames, nt = 30, 3
    targets = _targets(x0, frames, nt)
    sim = Simulator(x0, SceneConfig(), precision="f64")
    sim.set_design(design)
    loss, grad = sim.rollout_grad(targets, frames=frames, checkpoint_every=7)        # 7 does not divide 30: ragged last segment
    grad = _np(grad)
    assert loss > 0 and np.isfinite(grad).all() and np.abs(grad).max() > 0
    # same result with another checkpoint spacing (recomputation is exact in structure, fp64 in value)
    loss2, grad2 = sim.rollout_grad(targets, frames=frames, checkpoint_every=30)
    assert abs(loss2 - loss) <= 1e-12 * loss and np.abs(_np(grad2) - grad).max() <= 1e-9 * np.abs(grad).max()
    order = np.argsort(-np.abs(grad))
    picks = [int(order[0]), int(order[1]), int(order[len(order) // 2]), int(order[5])]
    eps = 1e-5
    for i in picks:
        d = design.copy(); d[i] += eps
        sim.set_design(d); lp, _ = sim.rollout_grad(targets, frames=frames)
        d[i] -= 2 * eps
        sim.set_design(d); lm, _ = sim.rollout_grad(targets, frames=frames)
        num = (lp - lm) / (2 * eps)
        # measured: the largest component agrees to 4e-8 relative; every component carries the same ABSOLUTE finite-difference
        # noise of ~2e-15 = 4e-8 max|grad| (the loss is a 5e-9 number summed over an impact trajectory)
        assert abs(num - grad[i]) <= 1e-5 * abs(grad[i]) + 2e-7 * np.abs(grad).max(), (i, num, grad[i])


def test_fp32_gradient_agrees_with_fp64():
    x0, design = _grad_scene()
    frames, nt = 30, 3
    targets = _targets(x0, frames, nt)
    a = Simulator(x0, SceneConfig(), precision="f64"); a.set_design(design)
    b = Simulator(x0, SceneConfig()); b.set_design(design.astype(np.float32))
    la, ga = a.rollout_grad(targets, frames=frames)
    lb, gb = b.rollout_grad(targets, frames=frames)
    ga, gb = _np(ga), _np(gb).astype(np.float64)
    # the loss is sum |x - xt|^2 with |x - xt| ~ 4e-6 m: fp32 positions (ulp 2e-9 at 0.02 m) and the fp32 trajectory noise put
    # the fp32 value within ~1 % of the fp64 one (measured 0.9 %)
    assert abs(la - lb) <= 5e-2 * la, (la, lb)
    assert np.abs(ga - gb).max() <= 1e-1 * np.abs(ga).max(), (np.abs(ga - gb).max(), np.abs(ga).max())
    assert np.dot(ga, gb) / (np.linalg.norm(ga) * np.linalg.norm(gb)) > 0.99          # same descent direction
    # and the fp32 scene still steps on its hot path afterwards
    b.startup(); b.step(5)
    x, v = b.position_velocity()
    assert torch.isfinite(x).all() and x.dtype == torch.float32
