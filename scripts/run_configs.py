"""BASELINE.json configs[0], [2] and [3] on ONE B200 (configs[1] is bench.py at N = 1, configs[4] is bench.py --gpus 8).
Prints one JSON line per config; run under gpurun.  Device time from CUDA events on the library's streams.

  configs[0]  single ~10k-particle sphere, reference defaults, 1000 steps; checked against the committed oracle trajectory
              (tests/golden/config0_n10k.npz) at steps 100 / 500 / 1000.
  configs[2]  ~1M particles, soft shell (design x = +1 on the outer 2h), outward dead load on the shell, the neighbour structure
              (Morton sort, cell table, exact lists, cluster lists) REBUILT every step: rebuild and step timed separately.
  configs[3]  batched independent scenes, one GPU's share (8 of the 64 ~10k-particle scenes, seeds 0..7) stepping concurrently
              on 8 streams.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes

which = sys.argv[1:] or ["0", "2", "3"]
cfg = SceneConfig()
dev = torch.device("cuda:0")


def ev():
    return torch.cuda.Event(enable_timing=True)


def config0():
    x0, _ = scenes.jittered_sphere(10000, seed=0, low_drop=True)
    sim = Simulator(x0, cfg)
    info = sim.neighbor_info()
    sim.startup(); sim.step(0); sim.synchronize()
    gold_path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "config0_n10k.npz")
    gold = np.load(gold_path) if os.path.exists(gold_path) else None
    done, ms, parity = 0, 0.0, {}
    for cp in (100, 500, 1000):
        e0, e1 = ev(), ev()
        with torch.cuda.stream(sim.stream):
            e0.record(); sim.step(cp - done); e1.record()
        sim.synchronize()
        ms += e0.elapsed_time(e1); done = cp
        if gold is not None:
            x, v = (t.cpu().numpy() for t in sim.position_velocity())
            parity[str(cp)] = {"dx": float(np.abs(x - gold[f"x_{cp}"]).max()), "floor_x": float(gold[f"floor_x_{cp}"]),
                               "dv": float(np.abs(v - gold[f"v_{cp}"]).max()), "floor_v": float(gold[f"floor_v_{cp}"])}
    n = len(x0)
    print(json.dumps({"config": "configs[0]: single inflatable sphere, reference defaults, 1000 steps", "n_particles": n,
                      "mean_neighbors": info.total_pairs / n, "steps": 1000, "ms_per_step": ms / 1000,
                      "particle_steps_per_s": n * 1000 / (ms * 1e-3), "parity_vs_oracle_golden": parity}), flush=True)
    sim.close()


def config2(n=1_000_000, steps=50):
    os.environ["MIS_MERGE_LISTS"] = "1"      # rebuild-heavy use: cluster lists merged from the exact lists (faster rebuild, 3 % slower force kernel)
    x0, _ = scenes.jittered_sphere(n, seed=0, low_drop=True)     # resting just above the ground plane (the default centre would bury a 0.2 m sphere)
    n = len(x0)
    t = time.perf_counter()
    sim = Simulator(x0, cfg, graph_steps=-1)                      # a rebuild invalidates captured graphs: plain launches
    sim.synchronize()
    t_create = time.perf_counter() - t
    info = sim.neighbor_info()
    c = x0.mean(0)
    r = np.linalg.norm(x0 - c, axis=1)
    shell = r > r.max() - 2.0 * cfg.h
    xd = np.where(shell, 1.0, -1.0).astype(np.float32)           # soft shell: ratio ~ 0.9975 => stress factor ~1.5 instead of ~199.5
    sim.set_design(xd)
    f = np.tile(np.asarray(cfg.external_force, np.float32), (n, 1))
    f[shell] += 2e-3 * ((x0[shell] - c) / r[shell, None]).astype(np.float32)      # outward dead load: the reference API's stand-in for pressure
    sim.set_external_forces(f)
    sim.startup(); sim.step(2); sim.synchronize()
    ms_build = ms_step = 0.0
    for _ in range(steps):
        e0, e1, e2 = ev(), ev(), ev()
        with torch.cuda.stream(sim.stream):
            e0.record()
            sim.rebuild_neighbors()                               # idempotent: queries are centred on x0 (sim.py:161,178,203,224)
            e1.record()
            sim.step(1)
            e2.record()
        sim.synchronize()
        ms_build += e0.elapsed_time(e1); ms_step += e1.elapsed_time(e2)
    x = sim.position().cpu().numpy()
    print(json.dumps({"config": "configs[2]: ~1M particles, soft shell + outward dead load, neighbour structure rebuilt every step",
                      "n_particles": n, "mean_neighbors": info.total_pairs / n, "shell_particles": int(shell.sum()), "steps": steps,
                      "ms_rebuild_per_step": ms_build / steps, "ms_step": ms_step / steps, "create_s": t_create,
                      "particle_steps_per_s_with_rebuild": n * steps / ((ms_build + ms_step) * 1e-3),
                      "particle_steps_per_s_step_only": n * steps / (ms_step * 1e-3), "finite": bool(np.isfinite(x).all()),
                      "max_radial_displacement": float(np.abs(np.linalg.norm(x - x.mean(0), axis=1) - r).max())}), flush=True)
    sim.close()
    os.environ.pop("MIS_MERGE_LISTS", None)


def config3(scenes_per_gpu=8, n=10000, steps=512):
    sims = []
    for seed in range(scenes_per_gpu):
        x0, _ = scenes.jittered_sphere(n, seed=seed, low_drop=True)
        sims.append(Simulator(x0, cfg))
    for s in sims:
        s.startup(); s.step(64)
    torch.cuda.synchronize()
    total = sum(s.n for s in sims)
    gate = ev()
    gate.record(torch.cuda.current_stream())
    ends = []
    for s in sims:
        s.stream.wait_event(gate)
    for s in sims:
        s.step(steps)
        e = ev(); e.record(s.stream); ends.append(e)
    torch.cuda.synchronize()
    ms_conc = max(gate.elapsed_time(e) for e in ends)
    # the same scenes one after the other
    ms_seq = 0.0
    for s in sims:
        e0, e1 = ev(), ev()
        with torch.cuda.stream(s.stream):
            e0.record(); s.step(steps); e1.record()
        s.synchronize()
        ms_seq += e0.elapsed_time(e1)
    ok = all(bool(torch.isfinite(s.position()).all()) for s in sims)
    print(json.dumps({"config": "configs[3]: batched independent scenes, one GPU's share (8 of 64 scenes of ~10k particles)",
                      "scenes": scenes_per_gpu, "n_particles_total": total, "steps": steps,
                      "ms_per_step_all_scenes_concurrent": ms_conc / steps, "ms_per_step_all_scenes_sequential": ms_seq / steps,
                      "particle_steps_per_s_concurrent": total * steps / (ms_conc * 1e-3),
                      "particle_steps_per_s_sequential": total * steps / (ms_seq * 1e-3), "finite": ok}), flush=True)
    for s in sims:
        s.close()


if "0" in which:
    config0()
if "2" in which:
    config2()
if "3" in which:
    config3()
