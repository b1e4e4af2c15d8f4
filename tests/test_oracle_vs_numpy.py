"""C oracle (fp32, hash grid, Jacobi polar) against the independent fp64 numpy restatement
(brute-force neighbours, LAPACK SVD)."""
import numpy as np

from conftest import make_oracle, deformed
from oracle.np_oracle import NpOracle
from meshless_inflatable_softbody_b200 import SceneConfig, scenes


def _np_oracle(x0, cfg):
    p = NpOracle(x0.astype(np.float64), h=cfg.h, dt=cfg.time_step, damping=cfg.damping,
                 k_col=cfg.collision_penalty_stiffness, col_range=cfg.collision_range)
    p.set_material(cfg.youngs_modulus, cfg.poisson_ratio)
    p.set_mass(cfg.mass)
    p.set_design(cfg.design_x)
    p.fext[:] = cfg.external_force
    return p


def test_fields_match_fp64():
    cfg = SceneConfig()
    x0, _ = scenes.jittered_sphere(500, seed=1)
    o, p = make_oracle(x0, cfg), _np_oracle(x0, cfg)
    rho, vol = o.volume()
    assert np.abs(vol - p.vol).max() < 2e-6 * p.vol.max()
    x = deformed(x0)
    ec, ep = o.eval(x), p.eval(x.astype(np.float64))
    for k, tol in (("A", 3e-6), ("R", 2e-6), ("F", 2e-6), ("S", 1e-4), ("f", 1e-4)):
        err = np.abs(ec[k] - ep[k]).max() / np.abs(ep[k]).max()
        assert err < tol, (k, err)


def test_neighbour_sets_equal_brute_force():
    for seed, spacing in ((0, 0.5), (1, 0.8), (2, 0.65)):
        x0, _ = scenes.jittered_sphere(400, seed=seed, spacing=spacing)
        o = make_oracle(x0)
        p = NpOracle(x0.astype(np.float64))
        cnt, off, flat = o.neighbor_lists()
        # brute force in fp32 with the reference's predicate sqrt(|d|^2)/h < 2
        d = x0[:, None, :] - x0[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        q = np.sqrt(d2) / np.float32(0.007)
        mask = (q < np.float32(2.0)) & ~np.eye(len(x0), dtype=bool)
        for i in range(len(x0)):
            assert np.array_equal(flat[off[i]:off[i + 1]], np.nonzero(mask[i])[0])


def test_trajectory_matches_fp64_within_noise():
    cfg = SceneConfig()
    x0, _ = scenes.jittered_sphere(400, seed=2, low_drop=True)
    o, p = make_oracle(x0, cfg), _np_oracle(x0, cfg)
    o.startup(); p.startup()
    o.step(60); p.step(60)
    assert np.abs(o.position() - p.x).max() < 3e-7
    assert np.abs(o.velocity() - p.v).max() < 5e-3
