import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from meshless_inflatable_softbody_b200 import SceneConfig, Simulator
x0 = bench.beam_scene(200_000, 0, 2)
print("n", len(x0), "bbox", x0.min(0), x0.max(0))
sim = Simulator(x0, SceneConfig())
print("k", sim.neighbor_info().total_pairs / len(x0))
sim.startup()
for k in range(8):
    sim.step(50)
    x, v = sim.position_velocity()
    print(k, "finite", torch.isfinite(x).all().item(), "dv", (v - v.mean(0)).abs().max().item(), "ymin", x[:, 1].min().item(), flush=True)
