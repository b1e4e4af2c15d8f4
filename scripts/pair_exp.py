"""Union-list overhead and gather-kernel times (CUDA events per launch), with / without in-cell pairing (MIS_PAIR_CELLS). Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 0
x0, _ = scenes.jittered_sphere(n, seed=0, low_drop=True)
sim = Simulator(x0, SceneConfig(), cluster_size=C)
info = sim.neighbor_info()
sim.startup(); sim.step(20); sim.synchronize()
a, b = sim.profile_step(50)
x = sim.position()
import torch
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sim.rebuild_neighbors(); sim.synchronize()
with torch.cuda.stream(sim.stream):
    e0.record()
    for _ in range(5):
        sim.rebuild_neighbors()
    e1.record()
sim.synchronize()
print("rebuild %.3f ms (merge=%s)" % (e0.elapsed_time(e1) / 5, os.environ.get("MIS_MERGE_LISTS", "1")), end="  ")
print("pair=%s C=%d n=%d k=%.1f union/k=%.3f deform %.1f us force %.1f us finite=%s" % (
    os.environ.get("MIS_PAIR_CELLS", "1"), info.cluster_size, sim.n, info.total_pairs / sim.n,
    info.union_entries * info.cluster_size / info.total_pairs, 1e3 * a / 50, 1e3 * b / 50, bool(x.isfinite().all())))
