"""Known-answer tests of the CPU oracle, derived from the reference's own mathematics
(SURVEY 4, items 1-6).  The reference ships no tests or golden vectors (parity unpinned);
these properties, the fp64 numpy cross-check (test_oracle_vs_numpy.py) and brute-force
neighbour search are the oracle's pins."""
import numpy as np
import pytest

from conftest import make_oracle, deformed
from oracle import c_oracle as co
from meshless_inflatable_softbody_b200 import SceneConfig, scenes

H = 0.007


def test_W_values_and_support():
    # sim.py:133-141: W(0) = 1/(pi h^3), W(q=1) = 1/(4 pi h^3), exactly 0 for q >= 2
    s = 1.0 / (np.pi * H ** 3)
    assert co.W([0, 0, 0], H) == pytest.approx(s, rel=1e-6)
    assert co.W([H, 0, 0], H) == pytest.approx(s / 4, rel=1e-5)
    assert co.W([2 * H, 0, 0], H) == 0.0
    assert co.W([0, 2.5 * H, 0], H) == 0.0
    assert np.all(co.nabla_W([0, 0, 2 * H], H) == 0.0)
    # continuity at q = 1
    a, b = co.W([H * (1 - 1e-4), 0, 0], H), co.W([H * (1 + 1e-4), 0, 0], H)
    assert a == pytest.approx(b, rel=1e-3)
    ga, gb = co.nabla_W([H * (1 - 1e-4), 0, 0], H), co.nabla_W([H * (1 + 1e-4), 0, 0], H)
    assert ga[0] == pytest.approx(gb[0], rel=1e-3)


def test_W_integrates_to_one():
    m = 40
    ax = (np.arange(-m, m) + 0.5) * (2 * H / m)
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    r = np.linalg.norm(g, axis=1)
    from oracle.np_oracle import W as Wnp
    total = Wnp(r, H).sum() * (2 * H / m) ** 3
    assert total == pytest.approx(1.0, abs=2e-3)
    # and the C version agrees with the numpy formula point-wise
    idx = np.random.default_rng(0).choice(len(g), 200, replace=False)
    for k in idx:
        assert co.W(g[k], H) == pytest.approx(float(Wnp(r[k], H)), rel=2e-5, abs=1e-3)


def test_nabla_W_is_gradient_of_W():
    rng = np.random.default_rng(1)
    for _ in range(50):
        x = rng.uniform(-1.2 * H, 1.2 * H, 3)
        eps = 1e-3 * H   # fp32 W: use a generous step
        fd = np.array([(co.W(x + eps * e, H) - co.W(x - eps * e, H)) / (2 * eps) for e in np.eye(3)])
        g = co.nabla_W(x, H)
        assert np.allclose(g, fd, rtol=2e-2, atol=2e-3 * np.abs(fd).max() + 1e3)


def test_polar_recovers_rotation_times_spd():
    rng = np.random.default_rng(2)
    for _ in range(100):
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] *= -1
        b = rng.standard_normal((3, 3)); spd = b @ b.T + 0.5 * np.eye(3)
        R = co.polar(q @ spd)
        assert np.abs(R - q).max() < 5e-6
        assert np.linalg.det(R.astype(np.float64)) == pytest.approx(1.0, abs=1e-5)
    # reflection is pushed into sigma_3: det R = +1 even for det A < 0
    A = np.diag([2.0, 1.0, -0.5])
    R = co.polar(A)
    assert np.linalg.det(R.astype(np.float64)) == pytest.approx(1.0, abs=1e-5)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-5
    # A = 0 falls back to identity
    assert np.array_equal(co.polar(np.zeros((3, 3))), np.eye(3, dtype=np.float32))


def test_sigma_svk():
    # sim.py:212-216 with F = diag(1+e,1,1)
    e, mu, lam, ratio = 0.01, 3.0, 5.0, 0.25
    S = co.sigma(np.diag([1 + e, 1, 1]), mu, lam, ratio)
    E11 = 0.5 * ((1 + e) ** 2 - 1)
    k = 200 - 199 * ratio
    assert S[0, 0] == pytest.approx((2 * mu * E11 + lam * E11) * k, rel=1e-4)
    assert S[1, 1] == pytest.approx(lam * E11 * k, rel=1e-4)
    assert S[0, 1] == 0.0


def test_lame_and_ratio(sphere800):
    o = make_oracle(sphere800)
    mu, lam, ratio = o.lame()
    E, nu = 1.5e5, 0.4
    assert mu[0] == pytest.approx(E / (2 * (1 + nu)), rel=1e-6)
    assert lam[0] == pytest.approx(E * nu / ((1 + nu) * (1 - 2 * nu)), rel=1e-5)
    assert ratio[0] == pytest.approx(0.5 * np.tanh(-3.0) + 0.5, rel=1e-4)   # 0.00247 => factor 199.5


def test_rest_state_has_no_force(sphere800):
    o = make_oracle(sphere800)
    ev = o.eval(sphere800)
    assert np.abs(ev["R"] - np.eye(3)).max() < 2e-6
    assert np.abs(ev["F"] - np.eye(3)).max() < 5e-6
    # fp32 floor: |F - I| ~ 1e-6 times stiffness 3e7 Pa; compare with a 1 % stretch
    stretched = (sphere800 - sphere800.mean(0)) * 1.01 + sphere800.mean(0)
    f_ref = np.abs(o.eval(stretched.astype(np.float32))["f"]).max()
    assert np.abs(ev["f"]).max() < 2e-3 * f_ref


def test_rigid_motion_invariance(sphere800):
    o = make_oracle(sphere800)
    x = deformed(sphere800, angle=0.7, strain=0.0, noise=0.0) + np.float32([0.01, 0.02, -0.03])
    ev = o.eval(x)
    th = 0.7
    Q = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    assert np.abs(ev["R"] - Q).max() < 5e-6
    assert np.abs(ev["F"] - np.eye(3)).max() < 2e-5
    stretched = (sphere800 - sphere800.mean(0)) * 1.01 + sphere800.mean(0)
    f_ref = np.abs(o.eval(stretched.astype(np.float32))["f"]).max()
    assert np.abs(ev["f"]).max() < 1e-2 * f_ref


def test_ballistic_with_zero_stiffness():
    # E = 0: velocity Verlet reproduces x0 + v0 t + a t^2 / 2 for a constant force (damping 1e-6 negligible)
    x0, _ = scenes.jittered_sphere(300, seed=3)
    cfg = SceneConfig(youngs_modulus=0.0)
    o = make_oracle(x0, cfg)
    o.startup()
    steps = 200
    o.step(steps)
    t = steps * cfg.time_step
    a = np.array(cfg.external_force) / cfg.mass
    expect = x0 + np.array(cfg.initial_velocity) * t + 0.5 * a * t * t
    assert np.abs(o.position() - expect).max() < 5e-7
    # damping 1e-6 is present: -damping*v/m*t = 4.5e-5 m/s over the run
    assert np.abs(o.velocity() - (np.array(cfg.initial_velocity) + a * t)).max() < 1e-4


def test_ground_penalty_single_particle():
    # sim.py:238-244: f_y = (1e-4 - y)^2 * 3e5 below y = 1e-4, else 0
    cfg = SceneConfig(external_force=(0.0, 0.0, 0.0), initial_velocity=(0.0, 0.0, 0.0), damping=0.0)
    for y, expect in ((5e-5, (1e-4 - 5e-5) ** 2 * 3e5), (-1e-3, (1e-4 + 1e-3) ** 2 * 3e5), (2e-4, 0.0)):
        x0 = np.array([[0.0, y, 0.0]], np.float32)
        o = make_oracle(x0, cfg)
        o.startup(v0=(0, 0, 0))
        o.step(1)
        # x' = x + 0.5 dt^2 f / m
        dy = float(o.position()[0, 1]) - float(np.float32(y))
        assert dy == pytest.approx(0.5 * cfg.time_step ** 2 * expect / cfg.mass, rel=1e-3, abs=1e-12)


def test_momentum_symmetric_pair_only(sphere800):
    # sim.py's pair force uses F_i with S_j (sim.py:233) and is not antisymmetric; the Taichi form is.
    cfg = SceneConfig(external_force=(0, 0, 0), initial_velocity=(0, 0, 0), ground_contact=False)
    x = deformed(sphere800, strain=0.03, noise=2e-5)
    o_sym = make_oracle(sphere800, cfg, symmetric_pair=1, no_contact=1)
    f = o_sym.eval(x)["f"].astype(np.float64)
    assert np.abs(f.sum(0)).max() < 1e-4 * np.abs(f).sum(0).max()
    o_ref = make_oracle(sphere800, cfg, no_contact=1)
    f2 = o_ref.eval(x)["f"].astype(np.float64)
    assert np.abs(f2.sum(0)).max() > 10 * np.abs(f.sum(0)).max()


def test_faithful_equals_cached_bitwise():
    x0, _ = scenes.jittered_sphere(400, seed=5, low_drop=True)
    a = make_oracle(x0); b = make_oracle(x0)
    a.startup(mode=co.FAITHFUL); b.startup(mode=co.CACHED)
    a.step(5, mode=co.FAITHFUL); b.step(5, mode=co.CACHED)
    assert np.array_equal(a.position(), b.position())
    assert np.array_equal(a.velocity(), b.velocity())
    assert np.array_equal(a.elastic_forces(), b.elastic_forces())


def test_threads_do_not_change_results(sphere800):
    a = make_oracle(sphere800); b = make_oracle(sphere800)
    a.set_threads(1); b.set_threads(4)
    a.startup(); b.startup(); a.step(3); b.step(3)
    assert np.array_equal(a.position(), b.position()) and np.array_equal(a.velocity(), b.velocity())
