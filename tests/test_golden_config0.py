"""The committed oracle trajectory of BASELINE configs[0] (tests/golden/config0_n10k.npz) is what the committed oracle source
produces: re-run the first checkpoint (100 steps of the ~10k-particle sphere, ~40 s on 8 cores) and compare.  Keeps the fixture the
GPU test relies on tied to oracle/mis_oracle.c."""
import os

import numpy as np

from conftest import make_oracle
from meshless_inflatable_softbody_b200 import SceneConfig, scenes


def test_oracle_reproduces_the_committed_config0_checkpoint():
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config0_n10k.npz"))
    cfg = SceneConfig()
    x0, _ = scenes.jittered_sphere(10000, seed=0, low_drop=True)
    assert len(x0) == int(gold["n"]) and float(x0.astype(np.float64).sum()) == float(gold["x0_checksum"])
    o = make_oracle(x0, cfg)
    o.startup(cfg.initial_velocity)
    o.step(100)
    # deterministic per-particle summation order: identical on the same toolchain; a libm that rounds tanhf (set_design) differently
    # would still stay inside the reorder floor stored with the fixture
    assert np.abs(o.position() - gold["x_100"]).max() <= float(gold["floor_x_100"])
    assert np.abs(o.velocity() - gold["v_100"]).max() <= float(gold["floor_v_100"])
    # the checkpoints record a drop that hits the ground plane and rebounds (sim.py:238-244 acts)
    assert gold["v_100"][:, 1].mean() < -0.3 and gold["v_1000"][:, 1].mean() > 0.0
