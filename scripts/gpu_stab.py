"""Noise growth of a free-falling sphere on the CUDA path (diagnostic; run under gpurun).
Prints max |v - mean v| every 10 steps for the deform variants, to compare with the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes

def run(n, steps, **kw):
    x0, _ = scenes.jittered_sphere(n, seed=0, centre=(0.0, 0.2, 0.0))
    sim = Simulator(x0, SceneConfig(), **kw)
    sim.startup()
    out = []
    for k in range(steps // 10):
        sim.step(10)
        x, v = sim.position_velocity()
        dv = (v - v.mean(0)).abs().max().item()
        out.append(dv)
        if not np.isfinite(dv):
            break
    print(n, kw, " ".join("%.2e" % d for d in out), flush=True)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    for kw in ({"two_pass_deform": False}, {"two_pass_deform": True}, {"two_pass_deform": False, "graph_steps": -1}):
        run(n, steps, **kw)
