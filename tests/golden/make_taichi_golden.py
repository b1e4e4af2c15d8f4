"""Golden fixtures of the Taichi prototype (sim_taichi.py) produced by EXECUTING the reference's own source.

    python tests/golden/make_taichi_golden.py        # needs /root/reference; ~10 min

Lifted with `ast` (warp_shim.lift) and run under tests/golden/taichi_shim.py (a numpy `ti`):
  options.py:3-9 (real = f64, h = 0.1, damping = 1e-5), utils.py:25-43 (W, nabla_W),
  sim_taichi.py:28-29 (frames, time_step), 32-86 (every field), 78-81 compute_ratio, 93-207 (compute_v_i ... startup),
  210-213 compute_loss, 240-294 setters, 317-321 loss().
The scene mirrors main() (sim_taichi.py:326-337): E = 1e5, nu = 0.4, m = 1e-2, Dirichlet where z > 0.85, pull force
(0, 0, -0.5) where z < 0.5, then startup() and the reference's own `loss(time_step)` loop (forward() per frame).
`frames` (sim_taichi.py:28) is overridden from 3000.  The point cloud (the reference loads ./pcd/spot_*.ply, absent)
is a synthetic cantilever: a jittered 3 x 3 x 19 lattice at spacing 0.5 h along z in [0, 0.9].

Outputs (tests/golden/sim_taichi_n{n}.npz): x0, rho_i, volume_i, mu, lam, ratio, free_points, external_forces, and per
saved frame f: position, velocity (state entering frame f) and the fields forward(f) computes from it -- A_pq, def_grad,
sigma, elastic_forces.  Everything fp64 (options.py:3).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import taichi_shim as ti                                  # noqa: E402
from warp_shim import lift                                # noqa: E402

REF = "/root/reference"
FRAMES = 40
SAVE = (0, 1, 10, 40)


def cantilever(seed=0):
    h = 0.1
    s = 0.5 * h
    ax = np.arange(3) * s - s
    az = np.arange(19) * s
    g = np.stack(np.meshgrid(ax, ax, az, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    g = g + rng.uniform(-0.2 * s, 0.2 * s, size=g.shape)
    g[:, 2] = np.clip(g[:, 2], 0.0, None)
    return g.astype(np.float32).astype(np.float64)        # fp32-representable so an fp32 engine starts from the same state


def main():
    pts = cantilever()
    ns = {"ti": ti, "np": np, "__name__": "sim_taichi_lifted"}
    lift(REF + "/options.py", [(3, 9)], ns)
    lift(REF + "/utils.py", [(25, 43)], ns)
    lift(REF + "/sim_taichi.py", [(28, 29)], ns)
    ns["frames"] = FRAMES + 1                              # sim_taichi.py:28 says 3000
    ns["points_np"] = pts                                  # what sim_taichi.py:20-25 would load
    ns["n_points"] = pts.shape[0]
    ns["tqdm"] = lambda it: it
    got = lift(REF + "/sim_taichi.py", [(32, 86), (93, 213), (240, 294), (317, 317)], ns)
    names = [n for _, n in got]
    for need in ("compute_ratio", "compute_v_i", "compute_A_pq", "svd", "compute_R_i", "compute_nabla_u", "compute_elastic_forces",
                 "compute_damping_forces", "advance", "forward", "startup", "compute_loss", "set_external_force", "set_dirichlet",
                 "set_youngs_modulus", "set_poisson_ratio", "set_mass", "set_target", "loss"):
        assert need in names, need
    n = ns["n_points"]
    # main(), sim_taichi.py:326-337
    ns["set_youngs_modulus"](1e5)
    ns["set_poisson_ratio"](0.4)
    ns["set_mass"](1e-2)
    edge = np.where(pts[:, 2] > 0.85)[0]
    for i in edge:
        ns["set_dirichlet"](i, ti.Vector([0., 0., 0.]))
    pull = np.where(pts[:, 2] < 0.5)[0]
    for i in pull:
        ns["set_external_force"](i, ti.Vector([0., 0., -5e-1]))
    ns["set_target"]()
    ns["startup"]()
    t0 = time.time()
    ns["loss"](ns["time_step"])                            # sim_taichi.py:339 / 317-321
    print("loss(): %d frames, n=%d, %.0f s, l=%g" % (ns["frames"], n, time.time() - t0, float(ns["l"])), flush=True)
    out = {"x0": pts, "frames": np.int64(FRAMES), "save_frames": np.array(SAVE), "time_step": np.float64(ns["time_step"]),
           "h": np.float64(ns["h"]), "damping": np.float64(ns["damping"]), "edge": edge, "pull": pull, "loss": np.float64(ns["l"])}
    for k in ("rho_i", "volume_i", "mu", "lam", "ratio", "free_points", "external_forces", "mass"):
        out[k] = np.array(ns[k])
    for f in SAVE:
        for k in ("position", "velocity", "A_pq", "def_grad", "sigma", "elastic_forces"):
            out[f"{k}_{f}"] = np.array(ns[k][f])
    path = os.path.join(HERE, f"sim_taichi_n{n}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; |x - x0|max at frame %d = %.3e" %
          (FRAMES, np.abs(out[f"position_{FRAMES}"] - pts).max()))


if __name__ == "__main__":
    main()
