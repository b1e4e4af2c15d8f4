"""The oracle pinned to the REFERENCE'S OWN SOURCE.

tests/golden/sim_py_*.npz and sim_taichi_*.npz hold outputs of /root/reference/sim.py and sim_taichi.py executed
statement by statement (ast-lifted, numpy `wp` / `ti` shims: tests/golden/warp_shim.py, taichi_shim.py,
make_simpy_golden.py, make_taichi_golden.py).  Here the C oracle (oracle/mis_oracle.c) is compared with them:

  * static fields (rho, volume, mu, lam, ratio): the fp32 fixture is the same literal arithmetic -> bit-equal
    up to the summation order of rho (<= 2 ulp);
  * per-evaluation fields A_pq, R, def_grad, S, elastic_forces at a rotated + 2 %-strained state: within
    FLOOR_MULT x the oracle's own fp32 summation-order noise floor (same oracle, candidate walk reversed) of the
    fixture computed in DOUBLE (the exact-arithmetic reading of the same source);
  * trajectories (position, velocity) after 1 / 20 / 100 frames of the reference's own diff_sim loop, with ground
    impact, against both the fp32 and the fp64 fixture, same rule;
  * the Taichi prototype (fp64, h = 0.1, symplectic Euler, identity rotation, symmetric pair force, Dirichlet +
    pull scene of sim_taichi.py:326-334) against the oracle's variant="taichi".
The GPU twins of these tests are in tests/test_gpu_golden.py.
"""
import glob
import os

import numpy as np
import pytest

from conftest import make_oracle
from oracle import c_oracle as co

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLOOR_MULT = 4.0


def _load(pattern):
    files = sorted(glob.glob(os.path.join(GOLD, pattern)))
    assert files, f"missing golden fixture {pattern}: run tests/golden/make_simpy_golden.py / make_taichi_golden.py"
    return [np.load(f) for f in files]


def fields_fixture():
    return _load("sim_py_fields_n*.npz")[0]


def trajectory_fixtures():
    return _load("sim_py_n*.npz")


def taichi_fixture():
    return _load("sim_taichi_n*.npz")[0]


def taichi_oracle(t, order=0, precision="f32"):
    x0 = t["x0"].astype(np.float32)
    o = co.Oracle(x0, h=float(t["h"]), dt=float(t["time_step"]), damping=float(t["damping"]), variant="taichi", precision=precision)
    o.set_order(order)
    o.set_external_forces(t["external_forces"])
    o.set_free_points(t["free_points"])
    o.set_youngs_modulus(1e5); o.set_poisson_ratio(0.4); o.set_mass(1e-2)      # sim_taichi.py:326-328
    o.set_design(np.full(len(x0), -10.0))                                      # sim_taichi.py:75
    return o


def test_static_fields_bit_level():
    g = fields_fixture()
    o = make_oracle(g["x0"])
    rho, vol = o.volume()
    mu, lam, ratio = o.lame()
    assert np.array_equal(mu, g["f32_mu"]) and np.array_equal(lam, g["f32_lam"])
    assert np.abs(ratio - g["f32_ratio"]).max() <= 1e-9            # tanhf vs numpy tanh: <= 1 ulp of 2.5e-3
    # rho is a sum over ~200 neighbours: same terms, the fixture's walk order is the shim's restated hash grid
    assert np.abs(rho - g["f32_rho"]).max() <= 4e-7 * rho.max()
    assert np.abs(vol - g["f32_volume"]).max() <= 4e-7 * vol.max()
    assert np.abs(rho - g["f64_rho"]).max() <= 2e-6 * rho.max()


@pytest.mark.parametrize("mode", [co.FAITHFUL, co.CACHED])
def test_fields_match_reference_source(mode):
    g = fields_fixture()
    x0, xdef = g["x0"], g["xdef"]
    a, b = make_oracle(x0), make_oracle(x0)
    b.set_order(1)
    e, er = a.eval(xdef, mode=mode), b.eval(xdef, mode=mode)
    for k, name in (("A", "A_pq"), ("R", "R"), ("F", "def_grad"), ("S", "S"), ("f", "elastic_forces")):
        ref64, ref32 = g[f"f64_{name}"], g[f"f32_{name}"]
        scale = np.abs(ref64).max()
        floor = np.abs(e[k] - er[k]).max()
        tol = FLOOR_MULT * floor + 1e-6 * scale
        assert np.abs(e[k] - ref64).max() <= tol, (name, "fp64", np.abs(e[k] - ref64).max(), tol)
        assert np.abs(e[k] - ref32).max() <= tol, (name, "fp32", np.abs(e[k] - ref32).max(), tol)
    # R is a proper rotation in the fixture as well (the contract the shim's svd3 states)
    assert np.abs(np.linalg.det(g["f64_R"]) - 1).max() < 1e-12


def test_trajectories_match_reference_source():
    for g in trajectory_fixtures():
        x0 = g["x0"]
        a, b = make_oracle(x0), make_oracle(x0)
        b.set_order(1)
        a.startup(); b.startup()
        done = 0
        for f in [int(f) for f in g["save_frames"]]:
            # the first frame through the literal per-candidate path (sim.py:218-235 as written), the rest hoisted (bit-identical)
            a.step(f - done, mode=co.FAITHFUL if f <= 1 else co.CACHED); b.step(f - done); done = f
            fx = np.abs(a.position() - b.position()).max()
            fv = np.abs(a.velocity() - b.velocity()).max()
            for tag in ("f32", "f64"):
                ex = np.abs(a.position() - g[f"{tag}_position_{f}"]).max()
                ev = np.abs(a.velocity() - g[f"{tag}_velocity_{f}"]).max()
                assert ex <= FLOOR_MULT * fx + 4e-9, (len(x0), f, tag, ex, fx)
                assert ev <= FLOOR_MULT * fv + 2e-5, (len(x0), f, tag, ev, fv)
            # fields of that frame, evaluated at the fixture's own position (no trajectory drift in the comparison)
            e = a.eval(g[f"f32_position_{f}"])
            er = b.eval(g[f"f32_position_{f}"])
            for k, name in (("A", "A_pq"), ("R", "R"), ("F", "def_grad")):
                ref = g[f"f32_{name}_{f}"]
                tol = FLOOR_MULT * np.abs(e[k] - er[k]).max() + 1e-6 * np.abs(ref).max()
                assert np.abs(e[k] - ref).max() <= tol, (name, f)
            ref = g[f"f32_elastic_forces_{f}"]
            scale = max(np.abs(ref).max(), np.abs(er["f"]).max())
            assert np.abs(e["f"] - ref).max() <= FLOOR_MULT * np.abs(e["f"] - er["f"]).max() + 1e-5 * scale, ("force", f)


def test_target_export_of_the_reference_loop():
    """sim.py:363-369 executed by the lifted diff_sim: position_{i}.npy / velocity_{i}.npy, i = 1..target_frames, at frames
    (frames // target_frames) * i, (n,3) fp32 -- the naming and frame selection export.py / Simulator.export_targets follow."""
    for g in trajectory_fixtures():
        files = [str(f) for f in g["target_files"]]
        assert files == sorted([f"position_{i}.npy" for i in range(1, 6)] + [f"velocity_{i}.npy" for i in range(1, 6)])
        stride = int(g["target_frame_stride"])
        for i in range(1, 6):
            f = stride * i
            if f in [int(s) for s in g["save_frames"]]:
                assert np.array_equal(g[f"target_position_{i}"], g[f"f32_position_{f}"])
                assert np.array_equal(g[f"target_velocity_{i}"], g[f"f32_velocity_{f}"])


def test_taichi_prototype_matches_reference_source():
    t = taichi_fixture()
    a, b = taichi_oracle(t), taichi_oracle(t, order=1)
    rho, vol = a.volume()
    assert np.abs(rho - t["rho_i"]).max() <= 2e-6 * rho.max()          # self-inclusive density, sim_taichi.py:97
    assert np.abs(vol - t["volume_i"]).max() <= 2e-6 * vol.max()
    mu, lam, ratio = a.lame()
    assert np.allclose(mu, t["mu"], rtol=1e-6) and np.allclose(lam, t["lam"], rtol=1e-6)
    a.startup((0.0, 0.0, 0.0)); b.startup((0.0, 0.0, 0.0))             # sim_taichi.py:203-207
    done = 0
    for f in [int(f) for f in t["save_frames"]]:
        if f == 0:
            continue
        a.step(f - done); b.step(f - done); done = f
        fx = np.abs(a.position() - b.position()).max()
        fv = np.abs(a.velocity() - b.velocity()).max()
        ex = np.abs(a.position() - t[f"position_{f}"]).max()
        ev = np.abs(a.velocity() - t[f"velocity_{f}"]).max()
        # an fp32 engine against the fp64 prototype: positions are ~1 (h = 0.1 scene), one fp32 ulp is 6e-8, and the rounding
        # of x every step (not the summation order the floor measures) is what separates the two: 1e-4 of max|v|
        assert ex <= FLOOR_MULT * fx + 2.5e-7, (f, ex, fx)
        assert ev <= FLOOR_MULT * fv + 1e-4 * max(0.2, np.abs(t[f"velocity_{f}"]).max()), (f, ev, fv)
    # Dirichlet particles never moved, pulled ones did (sim_taichi.py:329-334)
    x = a.position()
    assert np.array_equal(x[t["edge"]], t["x0"].astype(np.float32)[t["edge"]])
    assert np.abs(x[t["pull"]] - t["x0"][t["pull"]]).max() > 1e-3
    # fields forward(f) computes, at the fixture's own state
    f = int(t["save_frames"][-1])
    e = a.eval(t[f"position_{f}"].astype(np.float32))
    ref = t[f"elastic_forces_{f}"]
    assert np.abs(e["f"] - ref).max() <= 2e-4 * np.abs(ref).max()
    assert np.abs(e["F"] - t[f"def_grad_{f}"]).max() <= 2e-6
    assert np.abs(e["S"] - t[f"sigma_{f}"]).max() <= 2e-4 * np.abs(t[f"sigma_{f}"]).max()


def test_taichi_prototype_in_double_matches_reference_source_tightly():
    """The same oracle source built with real = double (options.py:3: real = ti.f64): the prototype's trajectory to ~1e-13."""
    t = taichi_fixture()
    a = taichi_oracle(t, precision="f64")
    rho, vol = a.volume()
    assert rho.dtype == np.float64
    assert np.abs(rho - t["rho_i"]).max() <= 1e-13 * rho.max() and np.abs(vol - t["volume_i"]).max() <= 1e-13 * vol.max()
    a.startup((0.0, 0.0, 0.0))
    done = 0
    for f in [int(f) for f in t["save_frames"]]:
        if f == 0:
            continue
        a.step(f - done); done = f
        assert np.abs(a.position() - t[f"position_{f}"]).max() <= 1e-13, f
        assert np.abs(a.velocity() - t[f"velocity_{f}"]).max() <= 1e-11, f
    e = a.eval(t[f"position_{done}"])
    for k, name, tol in (("F", "def_grad", 1e-13), ("S", "sigma", 1e-9), ("f", "elastic_forces", 1e-12)):
        assert np.abs(e[k] - t[f"{name}_{done}"]).max() <= tol * max(1.0, np.abs(t[f"{name}_{done}"]).max()), name


def test_double_build_of_the_oracle_matches_the_fp64_fixture_of_sim_py():
    """real = double on the sim.py path: A_pq, R (Jacobi polar here, LAPACK in the fixture), def_grad, S, force to ~1e-10 relative."""
    g = fields_fixture()
    o = make_oracle(g["x0"], precision="f64")
    e = o.eval(g["xdef"])
    for k, name in (("A", "A_pq"), ("R", "R"), ("F", "def_grad"), ("S", "S"), ("f", "elastic_forces")):
        ref = g[f"f64_{name}"]
        assert np.abs(e[k] - ref).max() <= 1e-9 * np.abs(ref).max(), (name, np.abs(e[k] - ref).max() / np.abs(ref).max())
