"""Short single-GPU command for ncu: the configs[1] workload (100k sphere on the DeepSDF plateau obstacle), a few un-graphed steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, DeepSDF

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = SceneConfig()
x0 = bench.sphere_on_obstacle(n, 0)
sim = Simulator(x0, cfg, graph_steps=-1)
net = DeepSDF(bench.obstacle_state())
sim.set_sdf_obstacle(net, bbox_model=bench.obstacle_bbox(cfg), fd_eps=1e-4)
sim.startup()
sim.step(60)          # the sphere has landed on the plateau: ~100 particles in the contact band
for _ in range(steps):
    sim.step(1)
sim.synchronize()
print("ok", len(x0), sim.contact_counts())
