"""Does the multi-GPU bench body (thick elongated body, reference defaults, low drop) stay finite on ONE GPU over ~1000 steps?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
lift = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
x0 = bench.beam_scene(n, seed=0, world=1)
x0[:, 1] += np.float32(lift)
sim = Simulator(x0, SceneConfig())
sim.startup()
for k in range(12):
    sim.step(100)
    x, v = sim.position_velocity()
    print(100 * (k + 1), bool(torch.isfinite(x).all()), float(torch.nan_to_num(v).abs().max()), float(torch.nan_to_num(x)[:, 1].min()), flush=True)
