"""Per-kernel device time of the step for each (cluster_size, lanes) combination (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
combos = [(1, 8), (2, 8), (2, 16), (4, 8), (4, 16), (4, 32), (2, 32)]
if len(sys.argv) > 2:
    combos = [tuple(int(v) for v in a.split("x")) for a in sys.argv[2:]]
x0, _ = scenes.jittered_sphere(n, seed=0, low_drop=True)
ref = None
for C, G in combos:
    sim = Simulator(x0, SceneConfig(), cluster_size=C, lanes_per_particle=G)
    info = sim.neighbor_info()
    sim.startup(); sim.step(20); sim.synchronize()
    d, f = sim.profile_step(50)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sim.stream):
        e0.record(); sim.step(256); e1.record()
    sim.synchronize()
    ms = e0.elapsed_time(e1) / 256
    x, v = sim.position_velocity()
    if ref is None:
        ref = (x.clone(), v.clone())
    print("C=%d G=%2d  deform %.1f us  force %.1f us  step %.1f us  %.3e particle-steps/s  k=%.1f union=%.2f  dx=%.2e dv=%.2e finite=%s" % (
        C, G, 1e3 * d / 50, 1e3 * f / 50, 1e3 * ms, len(x0) / (ms * 1e-3), info.total_pairs / len(x0), info.union_entries * C / info.total_pairs,
        (x - ref[0]).abs().max().item(), (v - ref[1]).abs().max().item(), bool(torch.isfinite(x).all())), flush=True)
    sim.close()
