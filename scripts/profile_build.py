"""Short single-GPU command for ncu: the HBM-bound kernels of the path at n = 1e6 -- Morton/radix cell binning (mis_build_neighbors),
the stand-alone integrate kernel (k_reintegrate: force_1 + part_1 after an external-force change) and the state export."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cfg = SceneConfig()
x0, _ = scenes.jittered_sphere(n, seed=0, low_drop=True)
sim = Simulator(x0, cfg, graph_steps=-1)
sim.startup(); sim.step(2)
sim.rebuild_neighbors()                                   # the profiled build
sim.set_all_external_force([0.0, -2e-3, 0.0]); sim.step(0)   # k_gather_vec3 + k_reintegrate
x, v = sim.position_velocity()                            # k_export_vec3 x 2
sim.synchronize()
print("ok", len(x0), float(x[:, 1].min()))
