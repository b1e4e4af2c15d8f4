"""Short single-GPU command for ncu: build a scene, prime, run a few un-graphed steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cluster = int(sys.argv[4]) if len(sys.argv) > 4 else 0
x0, _ = scenes.jittered_sphere(n, seed=0, low_drop=True)
sim = Simulator(x0, SceneConfig(), lanes_per_particle=lanes, cluster_size=cluster, graph_steps=-1)
sim.startup()
for _ in range(steps):
    sim.step(1)
sim.synchronize()
x = sim.position()
print("ok", len(x0), float(x[:, 1].min()))
