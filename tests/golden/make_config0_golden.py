"""Golden trajectory of BASELINE configs[0] (single ~10k-particle sphere, reference defaults, 1000 steps) from the CPU oracle.

    python tests/golden/make_config0_golden.py        # ~7 min on 8 cores

Writes tests/golden/config0_n10k.npz: the oracle's position / velocity (fp32, caller order) after 100, 500 and 1000 steps and, per
checkpoint, the oracle's own fp32 summation-order noise floor (same oracle, candidate walk reversed).  The scene is regenerated from
the seed by the test (scenes.jittered_sphere(10000, seed=0)); only the results are stored.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from meshless_inflatable_softbody_b200 import SceneConfig, scenes   # noqa: E402
from oracle import c_oracle as co                                     # noqa: E402

CHECKPOINTS = (100, 500, 1000)


def make(order, x0, cfg):
    o = co.Oracle(x0, h=cfg.h, dt=cfg.time_step, damping=cfg.damping, k_col=cfg.collision_penalty_stiffness, col_range=cfg.collision_range)
    o.set_order(order)
    o.set_all_external_force(cfg.external_force); o.set_youngs_modulus(cfg.youngs_modulus)
    o.set_poisson_ratio(cfg.poisson_ratio); o.set_mass(cfg.mass); o.set_design(cfg.design_x)
    o.startup(cfg.initial_velocity)
    return o


def main():
    cfg = SceneConfig()
    x0, _ = scenes.jittered_sphere(10000, seed=0, low_drop=True)      # low drop: ground impact starts around step 30
    a, b = make(0, x0, cfg), make(1, x0, cfg)
    out = {"n": len(x0), "x0_checksum": np.float64(x0.astype(np.float64).sum())}
    done = 0
    t0 = time.time()
    for cp in CHECKPOINTS:
        a.step(cp - done); b.step(cp - done); done = cp
        x, v = a.position(), a.velocity()
        out[f"x_{cp}"] = x.astype(np.float32); out[f"v_{cp}"] = v.astype(np.float32)
        out[f"floor_x_{cp}"] = np.float64(np.abs(x - b.position()).max())
        out[f"floor_v_{cp}"] = np.float64(np.abs(v - b.velocity()).max())
        print(cp, out[f"floor_x_{cp}"], out[f"floor_v_{cp}"], "%.0f s" % (time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "config0_n10k.npz"), **out)


if __name__ == "__main__":
    main()
