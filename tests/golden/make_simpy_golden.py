"""Golden fixtures of the particle path produced by EXECUTING the reference's own source.

    python tests/golden/make_simpy_golden.py [--n 260] [--frames 100]      # needs /root/reference; ~25 min

What runs: the top-level statements of /root/reference/sim.py at the line ranges in LIFT below -- constants (21-26,
63-69), every state array (72-99, 105), the kernels and @wp.func helpers (107-110, 133-273), the setters (279-308),
clear_grads (310-319), the hash-grid construction (123-127) and the rollout `diff_sim` itself (341-372) -- lifted with
`ast` and executed under tests/golden/warp_shim.py (a numpy `wp`).  The scene set-up mirrors main() (sim.py:441-444).
What does not run: argparse, asset loading, the DeepSDF checkpoint (sim.py:29-60,100-104), none of which exist here.

Inputs the reference would take from its assets are synthetic: a jittered-lattice sphere, spacing 0.5 h,
placed so that ground impact starts around step 10.  `frames` (sim.py:63) is overridden from 3000 to --frames and
`target_frames` from 100 to 5 so the reference's own target export (sim.py:363-369) writes position_{i}.npy /
velocity_{i}.npy for frames 20, 40, ... which are stored too (they pin export naming and frame selection).

Outputs (tests/golden/sim_py_n{n}.npz), for the fp32 run (real = wp.float32, as sim.py:22) and an fp64 run (real,
vec, mat overridden to the double types -- the "exact arithmetic" reading of the same source):
  x0, rho, volume, mu, lam, ratio; per saved frame f: position, velocity, A_pq, def_grad, elastic_forces and
  R = compute_R_i(A_pq[f][i]), S = compute_sigma(def_grad[f][i], ...) evaluated by the lifted functions.
Also the same fp32 run with candidate generator "brute" (every index) at the first checkpoint, to show that the
hash-grid restatement in the shim changes nothing but summation order.
"""
import argparse
import os
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import warp_shim as wp                                            # noqa: E402
from meshless_inflatable_softbody_b200 import scenes              # noqa: E402

REF = "/root/reference/sim.py"
LIFT_TYPES = [(21, 26)]
LIFT_CONSTS = [(63, 69)]
LIFT_STATE = [(72, 99), (105, 110), (112, 115)]
LIFT_GRID = [(123, 127)]
LIFT_KERNELS = [(133, 273), (279, 319), (341, 341)]
SAVE_FRAMES = (1, 20, 100)


def scene(n, seed=3):
    x0, _ = scenes.jittered_sphere(n, seed=seed, low_drop=True)
    x0 = x0.astype(np.float64)
    x0[:, 1] += 0.0003 - x0[:, 1].min()          # lowest particle 0.3 mm above y = 0: impact from step ~10
    return x0.astype(np.float32)


def deformed(x0, seed=0, angle=0.3, strain=0.02, noise=1e-5):
    """Rotation + 2 % random strain + noise: a state in which every field (A, R, F, S, force) is far from trivial."""
    rng = np.random.default_rng(seed)
    Q = np.array([[np.cos(angle), -np.sin(angle), 0], [np.sin(angle), np.cos(angle), 0], [0, 0, 1]])
    G = Q @ (np.eye(3) + strain * rng.standard_normal((3, 3)))
    c = x0.mean(0)
    return ((x0 - c) @ G.T + c + noise * rng.standard_normal(x0.shape)).astype(np.float32)


def namespace(x0, frames, double, query):
    """The lifted module: returns the namespace after set-up as in main() (sim.py:441-444)."""
    wp.QUERY_MODE = query
    ns = {"wp": wp, "np": np, "__name__": "sim_lifted"}
    wp.lift(REF, LIFT_TYPES, ns)
    if double:
        ns["real"], ns["vec"], ns["mat"] = wp.float64, wp.vec3d, wp.mat33d
        ns["h"], ns["damping"] = ns["real"](0.007), ns["real"](1e-6)          # sim.py:25-26 re-evaluated in double
    wp.lift(REF, LIFT_CONSTS, ns)
    if double:
        ns["time_step"] = ns["real"](5e-5)                                     # sim.py:65,68,69 in double
        ns["collision_penalty_stiffness"] = ns["real"](3e5)
        ns["collision_range"] = ns["real"](1e-4)
    ns["frames"] = frames                          # sim.py:63 says 3000
    ns["target_frames"] = 5                        # sim.py:64 says 100
    # what sim.py:41-53 would have produced from the assets
    ns["points_np"] = x0.astype(np.float64)
    ns["n_points"] = x0.shape[0]
    ns["args"] = types.SimpleNamespace(name="fixture", set_target=True, render=False, debug=False, init=False)
    ns["tqdm"] = lambda it: it
    ns["create_folder"] = lambda p, exist_ok=True: os.makedirs(p, exist_ok=True)
    wp.lift(REF, LIFT_STATE, ns)
    wp.lift(REF, LIFT_GRID, ns)
    lifted = wp.lift(REF, LIFT_KERNELS, ns)
    names = [n for _, n in lifted]
    for need in ("W", "nabla_W", "compute_v_i", "compute_A_pq", "compute_R_i", "compute_nabla_u", "compute_sigma",
                 "compute_elastic_forces", "compute_collision_penalty", "part_1", "part_2", "startup", "compute_loss",
                 "set_all_external_force", "set_youngs_modulus", "set_poisson_ratio", "set_mass", "clear_grads", "diff_sim"):
        assert need in names, need
    n = ns["n_points"]
    assert "compute_ratio" in ns
    # main(), sim.py:441-444
    ns["set_all_external_force"](ns["vec"]([0., -1e-3, 0.]))
    wp.launch(kernel=ns["set_youngs_modulus"], dim=n, inputs=[1.5e5, ns["youngs_modulus"], ns["poisson_ratio"], ns["mu"], ns["lam"]])
    wp.launch(kernel=ns["set_poisson_ratio"], dim=n, inputs=[0.4, ns["youngs_modulus"], ns["poisson_ratio"], ns["mu"], ns["lam"]])
    ns["set_mass"](1e-4)
    return ns


def run_fields(x0, xdef, double=False, query="grid"):
    """One force evaluation at a prescribed deformed position: the three launches of sim.py:349-351 (frame slot 0)."""
    ns = namespace(x0, 1, double, query)
    n = ns["n_points"]
    t0 = time.time()
    ns["position"][0] = wp.from_numpy(xdef, dtype=ns["vec"])
    wp.launch(kernel=ns["compute_ratio"], dim=n, inputs=[ns["x"], ns["ratio"]])
    wp.launch(kernel=ns["compute_A_pq"], dim=n, inputs=[ns["grid"].id, ns["position"][0], ns["init_position"], ns["mass"], ns["A_pq"][0], ns["h"], 0])
    wp.launch(kernel=ns["compute_nabla_u"], dim=n, inputs=[ns["grid"].id, ns["position"][0], ns["init_position"], ns["volume"], ns["A_pq"][0], ns["def_grad"][0], ns["h"], 0])
    wp.launch(kernel=ns["compute_elastic_forces"], dim=n, inputs=[ns["grid"].id, ns["position"][0], ns["init_position"], ns["volume"], ns["A_pq"][0], ns["def_grad"][0], ns["mu"], ns["lam"], ns["ratio"], ns["elastic_forces"][0], ns["h"], 0])
    print("fields(n=%d, double=%s, query=%s): %.0f s" % (n, double, query, time.time() - t0), flush=True)
    out = {k: ns[k].numpy() for k in ("rho", "volume", "mu", "lam", "ratio")}
    out["A_pq"], out["def_grad"], out["elastic_forces"] = ns["A_pq"][0].numpy(), ns["def_grad"][0].numpy(), ns["elastic_forces"][0].numpy()
    R = np.empty_like(out["A_pq"])
    S = np.empty_like(R)
    for i in range(n):
        R[i] = ns["compute_R_i"](ns["A_pq"][0][i])
        S[i] = ns["compute_sigma"](ns["def_grad"][0][i], ns["mu"][i], ns["lam"][i], ns["ratio"][i])
    out["R"], out["S"] = R, S
    return out


def run(x0, frames, double=False, query="grid", save_frames=SAVE_FRAMES, workdir=None):
    ns = namespace(x0, frames, double, query)
    n = ns["n_points"]
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        t0 = time.time()
        ns["diff_sim"](False)                      # sim.py:450: forward rollout, the reference's own loop (341-372)
        print("diff_sim(%d frames, n=%d, double=%s, query=%s): %.0f s" % (frames, n, double, query, time.time() - t0), flush=True)
    finally:
        os.chdir(cwd)
    out = {k: ns[k].numpy() for k in ("rho", "volume", "mu", "lam", "ratio", "mass")}
    for f in save_frames:
        if f > frames:
            continue
        out[f"position_{f}"] = ns["position"][f].numpy()
        out[f"velocity_{f}"] = ns["velocity"][f].numpy()
        out[f"A_pq_{f}"] = ns["A_pq"][f].numpy()
        out[f"def_grad_{f}"] = ns["def_grad"][f].numpy()
        out[f"elastic_forces_{f}"] = ns["elastic_forces"][f].numpy()
        R = np.empty_like(out[f"A_pq_{f}"])
        S = np.empty_like(R)
        for i in range(n):
            R[i] = ns["compute_R_i"](ns["A_pq"][f][i])
            S[i] = ns["compute_sigma"](ns["def_grad"][f][i], ns["mu"][i], ns["lam"][i], ns["ratio"][i])
        out[f"R_{f}"], out[f"S_{f}"] = R, S
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=260)
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--fields-n", type=int, default=0, help="only write the one-evaluation fields fixture sim_py_fields_n*.npz")
    a = ap.parse_args()
    if a.fields_n:
        x0, _ = scenes.jittered_sphere(a.fields_n, seed=a.seed)
        xdef = deformed(x0)
        res = {"x0": x0, "xdef": xdef}
        for tag, dbl in (("f32", False), ("f64", True)):
            for k, v in run_fields(x0, xdef, double=dbl).items():
                res[f"{tag}_{k}"] = v
        path = os.path.join(HERE, f"sim_py_fields_n{len(x0)}.npz")
        np.savez_compressed(path, **res)
        print("wrote", path, os.path.getsize(path), "bytes")
        return
    x0 = scene(a.n, a.seed)
    print("particles:", len(x0), flush=True)
    res = {"x0": x0, "frames": np.int64(a.frames), "save_frames": np.array([f for f in SAVE_FRAMES if f <= a.frames])}
    with tempfile.TemporaryDirectory() as tmp:
        r32 = run(x0, a.frames, double=False, query="grid", workdir=tmp)
        for k, v in r32.items():
            res["f32_" + k] = v
        # the reference's own target export (sim.py:363-369), as written to disk by the lifted diff_sim
        tdir = os.path.join(tmp, "target", "fixture")
        files = sorted(os.listdir(tdir))
        res["target_files"] = np.array(files)
        step = a.frames // 5
        for i in range(1, 6):
            p = np.load(os.path.join(tdir, f"position_{i}.npy"))
            v = np.load(os.path.join(tdir, f"velocity_{i}.npy"))
            assert p.dtype == np.float32 and p.shape == (len(x0), 3)
            res[f"target_position_{i}"] = p
            res[f"target_velocity_{i}"] = v
        res["target_frame_stride"] = np.int64(step)
    with tempfile.TemporaryDirectory() as tmp:
        first = int(res["save_frames"][0])
        rb = run(x0, first, double=False, query="brute", save_frames=(first,), workdir=tmp)
        for k in ("position", "velocity", "A_pq", "def_grad", "elastic_forces"):
            d = np.abs(rb[f"{k}_{first}"] - r32[f"{k}_{first}"]).max()
            s = np.abs(r32[f"{k}_{first}"]).max()
            print(f"grid vs brute candidates, frame {first}: {k}: max|d| = {d:.3e} (scale {s:.3e})", flush=True)
            res[f"brute_minus_grid_{k}"] = np.float64(d)
        assert np.array_equal(rb["rho"], rb["rho"])
    with tempfile.TemporaryDirectory() as tmp:
        r64 = run(x0, a.frames, double=True, query="grid", workdir=tmp)
        for k, v in r64.items():
            res["f64_" + k] = v
    path = os.path.join(HERE, f"sim_py_n{len(x0)}.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
