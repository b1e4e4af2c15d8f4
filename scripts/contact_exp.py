"""Steady-state device time per step of the bench workload (sphere on the DeepSDF octahedron); knobs come from the environment
(MIS_SERIAL_CONTACT, MIS_SDF_L2_PIN, MIS_DEFORM_CARVEOUT, MIS_SK_CARVEOUT).  Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, DeepSDF

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
cfg = SceneConfig()
x0 = bench.sphere_on_obstacle(n, 0)
net = DeepSDF(bench.obstacle_state())
sim = Simulator(x0, cfg)
sim.set_sdf_obstacle(net, bbox_model=bench.obstacle_bbox(cfg), fd_eps=1e-4)
sim.startup(); sim.step(64); sim.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(sim.stream):
    e0.record(); sim.step(512); e1.record()
sim.synchronize()
knobs = {k: os.environ[k] for k in ("MIS_SERIAL_CONTACT", "MIS_SDF_L2_PIN", "MIS_DEFORM_CARVEOUT", "MIS_SK_CARVEOUT") if k in os.environ}
print("%.1f us/step  counts=%s  %s" % (1e3 * e0.elapsed_time(e1) / 512, sim.contact_counts(), knobs), flush=True)
