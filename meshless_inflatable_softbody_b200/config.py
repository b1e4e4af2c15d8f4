"""Scene configuration: the reference's module-level constants as one dataclass.

The reference has no config object; every value below is a module constant or a
literal in main() (sim.py:25-26, 63-69, 266, 441-444; Taichi variant:
options.py:3-9, sim_taichi.py:28-29, 326-328).  Names follow the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Tuple


@dataclass
class SceneConfig:
    # --- sim.py (Warp) defaults: the primary spec --------------------------------
    h: float = 0.007                       # sim.py:25  kernel radius (support 2h)
    damping: float = 1e-6                  # sim.py:26
    time_step: float = 5e-5                # sim.py:65
    frames: int = 3000                     # sim.py:63
    target_frames: int = 100               # sim.py:64
    collision_penalty_stiffness: float = 3e5   # sim.py:68
    collision_range: float = 1e-4          # sim.py:69
    youngs_modulus: float = 1.5e5          # sim.py:442
    poisson_ratio: float = 0.4             # sim.py:443
    mass: float = 1e-4                     # sim.py:444
    external_force: Tuple[float, float, float] = (0.0, -1e-3, 0.0)   # sim.py:441
    initial_velocity: Tuple[float, float, float] = (0.0, -0.4, 0.0)  # sim.py:266
    design_x: float = -1.0                 # sim.py:99  x.fill_(-1.)
    tanh_k: float = 3.0                    # sim.py:110 ratio = 0.5*tanh(3x)+0.5
    stiffness_a: float = 200.0             # sim.py:215 factor = 200 - 199*ratio
    stiffness_b: float = 199.0
    # --- variant switches (sim_taichi.py deltas) ---------------------
    symmetric_pair: bool = False           # f_ij uses F_j   (sim_taichi.py:157)
    identity_rotation: bool = False        # R = I           (sim_taichi.py:129)
    self_density: bool = False             # rho includes j==i (sim_taichi.py:97)
    euler: bool = False                    # symplectic Euler (sim_taichi.py:167-172)
    ground_contact: bool = True            # sim.py:238-244 (absent in sim_taichi.py)
    # --- build knobs (ours) -------------------------------------------------------
    rebuild_every: int = 0                 # re-sort/re-bin every K steps (idempotent: queries are on x0)

    @staticmethod
    def warp() -> "SceneConfig":
        return SceneConfig()

    @staticmethod
    def taichi() -> "SceneConfig":
        """options.py:3-9 + sim_taichi.py:28-29,81,151,326-328 (fp32 arithmetic, not the script's f64)."""
        return SceneConfig(h=0.1, damping=1e-5, time_step=4e-4, youngs_modulus=1e5, poisson_ratio=0.4,
                           mass=1e-2, external_force=(0.0, 0.0, 0.0), initial_velocity=(0.0, 0.0, 0.0),
                           design_x=-10.0, tanh_k=5.0, stiffness_a=1.0, stiffness_b=1.0,
                           symmetric_pair=True, identity_rotation=True, self_density=True, euler=True,
                           ground_contact=False)

    def with_(self, **kw) -> "SceneConfig":
        return replace(self, **kw)
