import os
import sys

import numpy as np
import pytest

# a lost peer must fail a multi-GPU halo test quickly instead of stalling every step for the default 20 s
os.environ.setdefault("MIS_HALO_TIMEOUT_MS", "5000")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; without a device they are skipped, never silently passed
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def make_oracle(x0, cfg=None, variant="warp", **kw):
    """Oracle configured like main() (sim.py:441-444) from a SceneConfig."""
    from meshless_inflatable_softbody_b200 import SceneConfig
    from oracle import c_oracle as co
    cfg = cfg or SceneConfig()
    o = co.Oracle(x0, h=cfg.h, dt=cfg.time_step, damping=cfg.damping, k_col=cfg.collision_penalty_stiffness,
                  col_range=cfg.collision_range, variant=variant, **kw)
    o.set_all_external_force(cfg.external_force)
    o.set_youngs_modulus(cfg.youngs_modulus)
    o.set_poisson_ratio(cfg.poisson_ratio)
    o.set_mass(cfg.mass)
    o.set_design(cfg.design_x)
    return o


@pytest.fixture(scope="session")
def sphere800():
    from meshless_inflatable_softbody_b200 import scenes
    x0, out_num = scenes.jittered_sphere(800, seed=0, low_drop=True)
    return x0


@pytest.fixture(scope="session")
def sphere3k():
    from meshless_inflatable_softbody_b200 import scenes
    x0, out_num = scenes.jittered_sphere(3000, seed=0, low_drop=True)
    return x0


def deformed(x0, seed=0, angle=0.3, strain=0.02, noise=1e-5):
    rng = np.random.default_rng(seed)
    Q = np.array([[np.cos(angle), -np.sin(angle), 0], [np.sin(angle), np.cos(angle), 0], [0, 0, 1]])
    G = Q @ (np.eye(3) + strain * rng.standard_normal((3, 3)))
    c = x0.mean(0)
    return ((x0 - c) @ G.T + c + noise * rng.standard_normal(x0.shape)).astype(np.float32)
