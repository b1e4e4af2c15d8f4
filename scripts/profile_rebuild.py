"""Short single-GPU command for ncu: one neighbour rebuild at n particles (launch list of the build kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x0, _ = scenes.jittered_sphere(n, seed=0, centre=(0.0, 0.3, 0.0))
sim = Simulator(x0, SceneConfig())
sim.synchronize()
sim.rebuild_neighbors()
sim.synchronize()
print("ok", len(x0), sim.gather_info())
