"""A/B of the gather modes (0: register-tiled cluster kernels, 1: shared-memory cell tiles, 2: tiles for the deformation pass only) on one GPU.
    python scripts/tile_ab.py [n ...]
Per n: neighbour/tile info, per-kernel CUDA-event times (mis_profile_step), chained step time, and the difference of the two
trajectories after 40 steps (must be at the fp32 reorder floor)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes

cfg = SceneConfig()
for n in [int(a) for a in sys.argv[1:]] or [100_000, 1_000_000]:
    x0, _ = scenes.jittered_sphere(n, seed=0, low_drop=True)
    sim = Simulator(x0, cfg)
    out = {"n": len(x0), "mean_k": sim.neighbor_info().total_pairs / len(x0)}
    state = {}
    for mode in (0, 1, 2):
        t0 = time.time()
        sim.set_gather_mode(mode)
        sim.synchronize()
        out[f"mode{mode}_switch_s"] = time.time() - t0
        sim.startup(); sim.step(40)
        x, v = sim.position_velocity()
        state[mode] = (x.clone(), v.clone())
        sim.step(64); sim.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sim.stream):
            e0.record(); sim.step(128); e1.record()
        sim.synchronize()
        d, f = sim.profile_step(50)
        out[f"mode{mode}"] = {"info": sim.gather_info(), "names": sim.kernel_names(), "deform_us": 1e3 * d / 50, "force_us": 1e3 * f / 50,
                              "step_us_chained": 1e3 * e0.elapsed_time(e1) / 128}
    out["dx_between_modes"] = max(float((state[0][0] - state[m][0]).abs().max()) for m in (1, 2))
    out["dv_between_modes"] = max(float((state[0][1] - state[m][1]).abs().max()) for m in (1, 2))
    out["finite"] = bool(torch.isfinite(state[1][0]).all())
    print(json.dumps(out), flush=True)
    sim.close()
