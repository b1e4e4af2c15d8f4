"""Per-step device time of the bench workload with / without the DeepSDF obstacle, and small-row GEMM layer times (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, DeepSDF

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
cfg = SceneConfig()
x0 = bench.sphere_on_obstacle(n, 0)
net = DeepSDF(bench.obstacle_state())
for m in (64, 128, 392, 512, 1024):
    for path in (1, 2):
        net.set_gemm_path(path)
        print("gemm rows=%5d path=%d: %.2f us/layer" % (m, path, 1e3 * net.profile_gemm(m, reps=50)), flush=True)
net.set_gemm_path(0)
for obstacle in (False, True):
    sim = Simulator(x0, cfg)
    if obstacle:
        sim.set_sdf_obstacle(net, bbox_model=bench.obstacle_bbox(cfg), fd_eps=1e-4)
    sim.startup(); sim.step(64); sim.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sim.stream):
        e0.record(); sim.step(256); e1.record()
    sim.synchronize()
    print("obstacle=%s: %.1f us/step  counts=%s" % (obstacle, 1e3 * e0.elapsed_time(e1) / 256, sim.contact_counts() if obstacle else None), flush=True)
    sim.close()
