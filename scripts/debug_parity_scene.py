"""Debug: the N = 8 parity-check scene (200k ellipsoid 12.8:1:1, low drop) single-domain in every gather mode."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes
cfg = SceneConfig()
aspect = float(sys.argv[1]) if len(sys.argv) > 1 else 12.8
x0 = scenes.jittered_ellipsoid(400_000, seed=7, aspect=(aspect, 1.0, 1.0), low_drop=True).astype(np.float32)
fext = np.tile(np.float32(cfg.external_force), (len(x0), 1))
print("n", len(x0), "extent", x0.min(0), x0.max(0))
for mode, kw in ((0, {}), (2, {}), (0, dict(cluster_size=4, lanes_per_particle=16)), (1, {})):
    sim = Simulator(x0, cfg, **kw)
    sim.set_gather_mode(mode)
    sim.set_external_forces(fext)
    print(mode, kw, sim.gather_info(), sim.neighbor_info().total_pairs / len(x0), sim.neighbor_info().max_neighbors)
    sim.startup()
    for k in range(12):
        sim.step(10)
        x, v = sim.position_velocity()
        fin = bool(torch.isfinite(x).all() and torch.isfinite(v).all())
        print("  step", 10 * (k + 1), "finite", fin, "vmax", float(v.abs().max()) if fin else None, "ymin", float(x[:, 1].min()) if fin else None)
        if not fin:
            bad = (~torch.isfinite(x).all(1)).nonzero()[:5, 0].cpu().numpy()
            print("  first bad ids", bad, x0[bad])
            break
    sim.close()
