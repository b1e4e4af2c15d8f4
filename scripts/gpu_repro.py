import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import deformed

def run(tag, do_eval, graph_steps, do_nbrs, steps=200):
    x0, _ = scenes.jittered_sphere(100_000, seed=0, centre=(0.0, 0.2, 0.0))
    sim = Simulator(x0, SceneConfig(), graph_steps=graph_steps)
    if do_nbrs:
        off, nb = sim.neighbors()
    if do_eval:
        c = x0.mean(0)
        f = sim.eval_forces(((x0 - c) * 1.01 + c).astype(np.float32))
        f2 = sim.eval_forces(deformed(x0, angle=1.1, strain=0.0, noise=0.0))
        print(tag, "eval finite", torch.isfinite(f).all().item(), torch.isfinite(f2).all().item(), f.abs().max().item(), f2.abs().max().item())
    sim.startup(); sim.step(steps)
    x, v = sim.position_velocity()
    print(tag, "finite", torch.isfinite(x).all().item(), torch.isfinite(v).all().item(), "dv", (v - v.mean(0)).abs().max().item(), flush=True)

run("eval+graph", True, 0, True)
run("noeval+graph", False, 0, False)
run("eval+nograph", True, -1, False)
run("eval+graph,nonbr", True, 0, False)
