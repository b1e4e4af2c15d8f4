"""The CUDA path (through the C-ABI) against fixtures produced by executing the reference's own source.

Same fixtures and the same tolerance rule as tests/test_simpy_golden.py (the oracle's twin of these tests):
tests/golden/sim_py_fields_n*.npz, sim_py_n*.npz (sim.py under the numpy `wp` shim, fp32 and fp64) and
sim_taichi_n*.npz (sim_taichi.py under the numpy `ti` shim, fp64).  Tolerance = FLOOR_MULT x the fp32
summation-order noise floor (oracle run twice with the candidate walk reversed; measured here on the fixture) plus
the small absolute term written at each assert.
"""
import numpy as np
import pytest
import torch

from conftest import make_oracle
from meshless_inflatable_softbody_b200 import SceneConfig, Simulator
from test_simpy_golden import FLOOR_MULT, fields_fixture, trajectory_fixtures, taichi_fixture, taichi_oracle

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def test_static_fields_vs_reference_source():
    g = fields_fixture()
    sim = Simulator(g["x0"], SceneConfig())
    f = sim.fields(want=("rho", "vol"))
    assert np.abs(_np(f["rho"]) - g["f64_rho"]).max() <= 3e-6 * g["f64_rho"].max()
    assert np.abs(_np(f["vol"]) - g["f64_volume"]).max() <= 3e-6 * g["f64_volume"].max()


@pytest.mark.parametrize("C,G", [(2, 8), (1, 8), (4, 16)])
def test_fields_vs_reference_source(C, G):
    g = fields_fixture()
    x0, xdef = g["x0"], g["xdef"]
    sim = Simulator(x0, SceneConfig(), cluster_size=C, lanes_per_particle=G, keep_fields=True)
    sim.set_state(xdef, np.zeros_like(xdef))
    got = sim.fields(want=("A", "R", "F", "S", "fel"))
    a, b = make_oracle(x0), make_oracle(x0)
    b.set_order(1)
    e, er = a.eval(xdef), b.eval(xdef)
    for k, ko, name, rel in (("A", "A", "A_pq", 1e-6), ("R", "R", "R", 1e-6), ("F", "F", "def_grad", 1e-6), ("S", "S", "S", 5e-5),
                             ("fel", "f", "elastic_forces", 2e-5)):
        ref = g[f"f64_{name}"]
        tol = FLOOR_MULT * np.abs(e[ko] - er[ko]).max() + rel * np.abs(ref).max()
        err = np.abs(_np(got[k]) - ref).max()
        assert err <= tol, (name, err, tol)
    assert np.abs(np.linalg.det(_np(got["R"]).astype(np.float64)) - 1).max() < 1e-5


def test_trajectories_vs_reference_source():
    """1 / 20 / 100 frames of the reference's own diff_sim loop (ground impact included) on the GPU."""
    for g in trajectory_fixtures():
        x0 = g["x0"]
        a, b = make_oracle(x0), make_oracle(x0)
        b.set_order(1)
        a.startup(); b.startup()
        sim = Simulator(x0, SceneConfig())
        sim.startup()
        done = 0
        for f in [int(f) for f in g["save_frames"]]:
            a.step(f - done); b.step(f - done); sim.step(f - done); done = f
            fx = np.abs(a.position() - b.position()).max()
            fv = np.abs(a.velocity() - b.velocity()).max()
            x, v = sim.position_velocity()
            for tag in ("f32", "f64"):
                ex = np.abs(_np(x) - g[f"{tag}_position_{f}"]).max()
                ev = np.abs(_np(v) - g[f"{tag}_velocity_{f}"]).max()
                assert ex <= FLOOR_MULT * fx + 4e-9, (len(x0), f, tag, ex, fx)
                assert ev <= FLOOR_MULT * fv + 2e-5, (len(x0), f, tag, ev, fv)
        assert (g[f"f64_velocity_{done}"][:, 1] > -0.39).any() or done < 100    # the 100-frame fixture contains the ground impact


def test_taichi_prototype_vs_reference_source():
    """sim_taichi.py's own scene (cantilever: Dirichlet z > 0.85, pull z < 0.5; sim_taichi.py:326-334) with the fp32 kernels."""
    t = taichi_fixture()
    x0 = t["x0"].astype(np.float32)
    cfg = SceneConfig.taichi()
    assert cfg.h == float(t["h"]) and cfg.time_step == float(t["time_step"]) and cfg.damping == float(t["damping"])
    sim = Simulator(x0, cfg)
    sim.set_external_forces(t["external_forces"].astype(np.float32))
    free = torch.as_tensor(t["free_points"].astype(np.float32))
    edge = torch.as_tensor(t["edge"])
    sim.set_dirichlet(edge, [0.0, 0.0, 0.0])
    assert np.array_equal(t["free_points"][t["edge"]], np.zeros((len(t["edge"]), 3))) and free.sum() == 3 * (len(x0) - len(edge))
    f = sim.fields(want=("rho", "vol"))
    assert np.abs(_np(f["rho"]) - t["rho_i"]).max() <= 3e-6 * t["rho_i"].max()
    a, b = taichi_oracle(t), taichi_oracle(t, order=1)
    a.startup((0.0, 0.0, 0.0)); b.startup((0.0, 0.0, 0.0))
    sim.startup((0.0, 0.0, 0.0))
    done = 0
    for fr in [int(f) for f in t["save_frames"]]:
        if fr == 0:
            continue
        a.step(fr - done); b.step(fr - done); sim.step(fr - done); done = fr
        fx = np.abs(a.position() - b.position()).max()
        fv = np.abs(a.velocity() - b.velocity()).max()
        x, v = sim.position_velocity()
        ex = np.abs(_np(x) - t[f"position_{fr}"]).max()
        ev = np.abs(_np(v) - t[f"velocity_{fr}"]).max()
        assert ex <= FLOOR_MULT * fx + 2.5e-7, (fr, ex, fx)                      # positions ~1: one fp32 ulp is 6e-8
        assert ev <= FLOOR_MULT * fv + 1e-4 * max(0.2, np.abs(t[f"velocity_{fr}"]).max()), (fr, ev, fv)
    assert np.array_equal(_np(x)[t["edge"]], x0[t["edge"]])
