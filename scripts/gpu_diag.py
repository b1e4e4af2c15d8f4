"""Stage-by-stage comparison of the CUDA path against the CPU oracle (run under gpurun).

Prints one line per stage; never raises on a mismatch so a single GPU call shows
everything.  Diagnostic tooling, not a test: tests/ holds the asserting versions.
"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes
from oracle import c_oracle as co

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max()), float(np.abs(b).max())

def main(n=3000, G=8, steps=50):
    print("device", torch.cuda.get_device_name(0))
    x0, out_num = scenes.jittered_sphere(n, seed=0, low_drop=True)
    n = len(x0)
    cfg = SceneConfig()
    t = time.time()
    sim = Simulator(x0, cfg, lanes_per_particle=G, keep_fields=True)
    sim.synchronize(); print("create+build %.3fs n=%d" % (time.time() - t, n))
    o = co.Oracle(x0)
    o.set_all_external_force(cfg.external_force); o.set_youngs_modulus(cfg.youngs_modulus)
    o.set_poisson_ratio(cfg.poisson_ratio); o.set_mass(cfg.mass); o.set_design(cfg.design_x)
    # --- cells
    ci, cc, pm = sim.cells()
    oc, occ, oids = o.grid_arrays()
    print("cell_index equal:", np.array_equal(ci.cpu().numpy(), oc), " cell_coords equal:", np.array_equal(cc.cpu().numpy(), occ))
    pm = pm.cpu().numpy()
    print("perm is permutation:", np.array_equal(np.sort(pm), np.arange(n)))
    info = sim.neighbor_info()
    print("info pairs", info.total_pairs, "max_k", info.max_neighbors, "dim", list(info.cell_dim), "min", list(info.cell_min))
    # within a cell ascending caller id
    s, e = sim.cell_ranges(); s = s.cpu().numpy(); e = e.cpu().numpy()
    print("hash-grid order from cell_index == oracle point_ids:", np.array_equal(np.argsort(ci.cpu().numpy(), kind="stable"), oids),
          " cells covered:", int((e - s).sum()) == n)
    # --- neighbours
    off, nb = sim.neighbors(); off = off.cpu().numpy(); nb = nb.cpu().numpy()
    cnt, ooff, oflat = o.neighbor_lists()
    same = np.array_equal(np.diff(off), cnt)
    if same:
        for i in range(n):
            if not np.array_equal(np.sort(nb[off[i]:off[i+1]]), oflat[ooff[i]:ooff[i+1]]): same = False; break
    print("neighbour lists bit-exact:", same, " mean k %.1f" % cnt.mean())
    # --- volume
    f = sim.fields(want=("rho", "vol"))
    rho, vol = o.volume()
    print("rho err %.3e / %.3e   vol err %.3e / %.3e" % (*rel(f["rho"].cpu().numpy(), rho), *rel(f["vol"].cpu().numpy(), vol)))
    # --- one force evaluation at a deformed configuration
    rng = np.random.default_rng(0)
    th = 0.3; Q = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    Gm = Q @ (np.eye(3) + 0.02 * rng.standard_normal((3, 3)))
    c = x0.mean(0)
    x = ((x0 - c) @ Gm.T + c + 1e-5 * rng.standard_normal(x0.shape)).astype(np.float32)
    sim.set_state(x, np.zeros_like(x))
    fg = sim.fields(want=("A", "R", "F", "S", "fel"))
    eo = o.eval(x)
    for k, ko in (("A", "A"), ("R", "R"), ("F", "F"), ("S", "S"), ("fel", "f")):
        print("eval %s: err %.3e / max %.3e" % ((k,) + rel(fg[k].cpu().numpy(), eo[ko])))
    o.set_order(1); eo1 = o.eval(x); o.set_order(0)
    print("oracle reorder floor: f %.3e  F %.3e" % (rel(eo1["f"], eo["f"])[0], rel(eo1["F"], eo["F"])[0]))
    fe = sim.eval_forces(x).cpu().numpy()
    print("eval_forces: err %.3e / %.3e" % rel(fe, eo["f"]))
    # --- trajectory
    sim.startup(); o.startup()
    for k in (1, 9, 40, steps):
        sim.step(k); o.step(k)
        xg, vg = sim.position_velocity()
        print("after +%d steps (frame %d): dx %.3e  dv %.3e   |v|max %.3f  ymin %.5f" % (
            k, sim.frame, np.abs(xg.cpu().numpy() - o.position()).max(), np.abs(vg.cpu().numpy() - o.velocity()).max(),
            float(np.abs(o.velocity()).max()), float(o.position()[:, 1].min())))
    # oracle own noise floor over the same number of steps
    o2 = co.Oracle(x0); o2.set_order(1)
    o2.set_all_external_force(cfg.external_force); o2.set_youngs_modulus(cfg.youngs_modulus)
    o2.set_poisson_ratio(cfg.poisson_ratio); o2.set_mass(cfg.mass); o2.set_design(cfg.design_x)
    o2.startup(); o2.step(sim.frame)
    print("oracle reorder floor after %d steps: dx %.3e dv %.3e" % (sim.frame, np.abs(o2.position() - o.position()).max(), np.abs(o2.velocity() - o.velocity()).max()))
    # --- ballistic bit-exactness (E = 0)
    sim.set_youngs_modulus(0.0); o.set_youngs_modulus(0.0)
    sim.startup(); o.startup(); sim.step(100); o.step(100)
    xg, vg = sim.position_velocity()
    print("ballistic (E=0) bit-exact: x", np.array_equal(xg.cpu().numpy(), o.position()), " v", np.array_equal(vg.cpu().numpy(), o.velocity()))
    # --- timing
    sim.set_youngs_modulus(cfg.youngs_modulus); sim.startup(); sim.step(64); sim.synchronize()
    for label, k in (("graph", 256),):
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sim.stream):
            ev0.record(); sim.step(k); ev1.record()
        sim.synchronize()
        ms = ev0.elapsed_time(ev1)
        print("timing n=%d G=%d: %.3f us/step  %.3e particle-steps/s" % (n, G, 1e3 * ms / k, n * k / (ms * 1e-3)))
    print("launches", sim.launch_count)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    main(n, G)
