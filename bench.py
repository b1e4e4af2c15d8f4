#!/usr/bin/env python
"""bench.py -- particle-steps/s of the meshless inflatable soft-body step on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU arm (oracle port)

One bench "step" = one simulation step (loop body sim.py:352-358: part_1, compute_A_pq,
compute_nabla_u, compute_elastic_forces, part_2) over all n particles of the scene.
Workload = BASELINE.json configs[1] scaled to what exists so far: ~100k-particle dense sphere
(spacing 0.5 h, ~215 neighbours/particle), reference defaults, ground-plane contact (the
reference's own per-step contact law, sim.py:238-244); see config.workload in the output.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "particle-steps/sec"
UNIT = "particle-steps/s"
# algorithmic bytes per particle per launch (SURVEY 8d / BASELINE.md 3.5, reference data-flow)
BYTES_FORCE = 112            # compute_elastic_forces: x0 12, V 4, A_pq 36, def_grad 36, mu/lam/ratio 12, force 12
BYTES_DEFORM = 64 + 100      # compute_A_pq 64 + compute_nabla_u 100
FLOP_PER_PAIR_FORCE = 70     # SURVEY 8d algorithmic flops
FLOP_PER_PAIR_DEFORM = 45


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_scene(n, seed):
    from meshless_inflatable_softbody_b200 import scenes
    x0, out_num = scenes.jittered_sphere(n, seed=seed, low_drop=True)
    return x0


def oracle_for(x0, cfg, threads=0):
    from oracle import c_oracle as co           # bench.py's cpu_baseline / reference arm: allowed importer
    o = co.Oracle(x0, h=cfg.h, dt=cfg.time_step, damping=cfg.damping, k_col=cfg.collision_penalty_stiffness,
                  col_range=cfg.collision_range, threads=threads)
    o.set_all_external_force(cfg.external_force); o.set_youngs_modulus(cfg.youngs_modulus)
    o.set_poisson_ratio(cfg.poisson_ratio); o.set_mass(cfg.mass); o.set_design(cfg.design_x)
    return o, co


def cpu_faithful_rate(cfg, n_sample, steps, warmup, seed=0):
    """Reference CPU path = oracle in FAITHFUL mode (27-cell walk, per-candidate svd3 + stress,
    5 passes per step as sim.py:353-358), OpenMP over particles on all host threads."""
    x0 = make_scene(n_sample, seed)
    o, co = oracle_for(x0, cfg)
    cores = co.max_threads()
    o.startup(cfg.initial_velocity, mode=co.FAITHFUL)
    if warmup:
        o.step(warmup, mode=co.FAITHFUL)
    t = time.perf_counter()
    o.step(steps, mode=co.FAITHFUL)
    dt = time.perf_counter() - t
    return len(x0) * steps / dt, dt, cores, len(x0)


def size_cpu_sample(cfg, n_full, total_steps, budget_s):
    """Pick a sample size whose total_steps faithful steps fit the time budget."""
    rate, dt, cores, n0 = cpu_faithful_rate(cfg, 4000, 1, 0)
    n_fit = int(rate * budget_s / max(1, total_steps))
    return max(2000, min(n_full, n_fit)), rate


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    n_s, _ = size_cpu_sample(cfg, args.n, args.steps + args.warmup, budget_s=60.0)
    rate, dt, cores, n_used = cpu_faithful_rate(cfg, n_s, args.steps, args.warmup)
    sample = (f"{args.steps} steps (+{args.warmup} warm-up) of a {n_used}-particle sphere, same spacing/params as the "
              f"GPU workload, oracle FAITHFUL mode (27-cell walk, per-candidate svd3), {cores} OpenMP threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_used, None),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference has no CPU implementation (device='cuda' hard-coded, Warp absent): this arm times our "
                "CPU restatement of sim.py:133-258,341-358 on the host cores",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n, mean_k):
    return {
        "workload": "BASELINE configs[1] (~100k-particle inflatable body, 1 B200) with the reference's per-step "
                    "ground-plane contact (sim.py:238-244); DeepSDF is start-up-only in the reference (sim.py:100)",
        "n_particles": int(n), "mean_neighbors": mean_k, "spacing_h": 0.5, "h": 0.007, "dt": 5e-5,
        "scene": "jittered-lattice sphere, low drop, reference defaults E=1.5e5 nu=0.4 m=1e-4 x=-1",
        "l2": "flushed between timed steps (256 MiB device write outside the event pairs); "
              "steady_state keys give the un-flushed chained-step figure",
    }


def run_ours(args, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from meshless_inflatable_softbody_b200 import Simulator

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    x0 = make_scene(args.n, seed=rank)
    n = len(x0)
    sim = Simulator(x0, cfg, device=str(dev), lanes_per_particle=args.lanes)
    info = sim.neighbor_info()
    mean_k = info.total_pairs / n
    sim.startup()
    sim.step(args.warmup)
    sim.synchronize()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    K = args.steps

    # ---- timed region A (`value`): K steps, L2 flushed between steps, one event pair per step
    sampler = ClockSampler(local_rank)
    launches0 = sim.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    sampler.start()
    with torch.cuda.stream(sim.stream):
        for a, b in evs:
            flush.fill_(1)                 # not timed: sits between the previous end event and this start event
            a.record()
            sim.step(1)
            b.record()
    barrier()
    ms_flushed = sum(a.elapsed_time(b) for a, b in evs)
    launches = sim.launch_count - launches0
    # ---- timed region B (steady state): K chained steps, CUDA-graph chunks, L2 warm
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(sim.stream):
        e0.record(); sim.step(K); e1.record()
    barrier()
    ms_steady = e0.elapsed_time(e1)
    clocks = sampler.stop()
    ms_flushed, ms_steady = max_over_ranks(ms_flushed), max_over_ranks(ms_steady)

    # ---- e2e: the public API with host buffers; per step H2D of the external-force field (the per-step
    # input of the reference API, sim.py:94,279-283) and D2H of position + velocity (sim.py:334,368-369)
    fext_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    fext_host[:] = torch.tensor(cfg.external_force)
    x_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    v_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    Ke = max(1, min(K, args.e2e_steps))
    for _ in range(3):
        sim.set_external_forces_host(fext_host); sim.step(1); sim.get_state_host(x_host, v_host); sim.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(sim.stream):
        g0.record()
        for _ in range(Ke):
            sim.set_external_forces_host(fext_host)
            sim.step(1)
            sim.get_state_host(x_host, v_host)
            sim.synchronize()              # the host reads the result of every step
        g1.record()
    barrier()
    ms_e2e = max_over_ranks(g0.elapsed_time(g1))
    assert bool(torch.isfinite(x_host).all()), "state diverged"

    # ---- roofline of the dominant kernel: CUDA events around every launch (same stream), L2 warm
    Kp = min(K, 50)
    ms_def, ms_for = sim.profile_step(Kp)
    ms_def, ms_for = ms_def / Kp, ms_for / Kp
    peak, peak_src = measured_peaks()
    kern = {
        "k_force": {"ms": ms_for, "bytes_per_particle": BYTES_FORCE, "gbs": n * BYTES_FORCE / (ms_for * 1e-3) / 1e9,
                    "pairs_per_s": info.total_pairs / (ms_for * 1e-3),
                    "fp32_tflops_algorithmic": FLOP_PER_PAIR_FORCE * info.total_pairs / (ms_for * 1e-3) / 1e12},
        "k_deform": {"ms": ms_def, "bytes_per_particle": BYTES_DEFORM, "gbs": n * BYTES_DEFORM / (ms_def * 1e-3) / 1e9,
                     "pairs_per_s": 2 * info.total_pairs / (ms_def * 1e-3),
                     "fp32_tflops_algorithmic": FLOP_PER_PAIR_DEFORM * info.total_pairs / (ms_def * 1e-3) / 1e12},
    }
    dom = "k_force" if ms_for >= ms_def else "k_deform"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom, {}).get(str(args.n))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": n * kern[dom]["bytes_per_particle"],
                "launch_ms": kern[dom]["ms"],
                "note": "at ~215 neighbours/particle the gather kernels are FP32-pipe / L1 bound, not HBM bound "
                        "(SURVEY 8d): fp32_fraction is the binding figure",
                "fp32_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12,
                "fp32_fraction": kern[dom]["fp32_tflops_algorithmic"] / (148 * 128 * 2 * 1.965e9 / 1e12),
                "kernels": kern}

    total_particles = n * world
    value = total_particles * K / (ms_flushed * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_flushed / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, n, mean_k),
        "pairs_per_sec": value * mean_k,
        "steady_state": {"value": total_particles * K / (ms_steady * 1e-3), "ms_per_step": ms_steady / K,
                         "note": "K chained steps in CUDA-graph chunks, no L2 flush (working set ~110 MB < 126 MB L2)"},
        "e2e": {"value": total_particles * Ke / (ms_e2e * 1e-3), "unit": UNIT, "steps": Ke,
                "h2d_bytes_per_step": n * 12, "d2h_bytes_per_step": n * 24,
                "what": "per step: mis_set_ext_force_host (pinned H2D) + mis_step(1) + mis_get_state_host (x, v D2H) + host sync"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        n_s, _ = size_cpu_sample(cfg, n, 1, budget_s=args.cpu_budget)
        rate, dt, cores, n_used = cpu_faithful_rate(cfg, n_s, 1, 0)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 step of a {n_used}-particle sphere (same spacing/params; {dt:.1f} s), oracle FAITHFUL mode "
                      f"(27-cell walk, per-candidate svd3, sim.py:353-358), {cores} OpenMP threads"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    sim.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=100_000, help="particles per GPU")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per particle (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from meshless_inflatable_softbody_b200 import SceneConfig
    cfg = SceneConfig()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        run_ours(args, cfg, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
