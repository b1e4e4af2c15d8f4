#!/usr/bin/env python
"""bench.py -- particle-steps/s of the meshless inflatable soft-body step on B200.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path of this repo
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU arm (oracle port on the host cores)

One bench "step" = one simulation step (loop body sim.py:352-358: part_1, compute_A_pq, compute_nabla_u,
compute_elastic_forces, part_2) over all particles, plus the per-step DeepSDF obstacle-contact query.

ONE workload family for every N (so that value(N) / (N * value(1)) is a scaling efficiency):

  default (--mode strong): BASELINE.json configs[4] -- one ~10M-particle elongated inflatable body (dense jittered
      ellipsoid 12.8 : 1 : 1, reference defaults) whose underside lands on a DeepSDF-encoded obstacle (the reference's
      9 x 1024 architecture, deepsdf.py:12-38, analytic truncated-octahedron weights: a flat plateau, so that a contact
      patch of >= 64 particles forms).  N = 1: single domain.  N > 1: static slab partition along x (slab.py), the
      ghost particles' new positions exchanged every step by the fused P2P push over NVLink peer memory (or NCCL
      send/recv with --halo nccl); the total particle count is fixed => "scaling": "strong".  The timed region starts
      with the scene IN CONTACT (pre-advanced untimed until the contact band holds >= 64 particles).
  The N = 1 line carries BASELINE.json configs[1] (~100k-particle sphere on the same obstacle, 1 B200) as the sub-record
  "configs1" (value, e2e, contact counts, per-kernel roofline at that size).
  --mode weak   : --n particles per GPU (default 1.25 M), body length grows with N.
  --mode batch  : independent ~--n-particle scenes, one per GPU per stream slot, no communication (configs[3]).
  --mode rebuild: configs[2] shape -- 1 M particles, soft shell with an outward dead load, neighbour structure rebuilt
                  every step (step + rebuild timed together).
  --mode configs1: configs[1] alone as the headline (N = 1).

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
# algorithmic bytes per particle per launch (reference data-flow, each array touched once per kernel)
BYTES_FORCE = 112 + 100      # compute_elastic_forces 112 + part_2 100 (fused into the force kernel; part_1 of the next step shares its reads)
BYTES_DEFORM = 64 + 100      # compute_A_pq 64 + compute_nabla_u 100
FLOP_PER_PAIR_FORCE = 70     # SURVEY 8d algorithmic flops
FLOP_PER_PAIR_DEFORM = 45
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12          # 148 SMs x 128 lanes x FMA x max SM clock
OBST_R, OBST_TOP = 0.05, 0.02                              # truncated octahedron |x|+|y|+|z| = OBST_R cut at y = OBST_TOP (m)
BAND_TARGET = 64                                            # particles in the contact band before any timed region starts
BODY_ASPECT = 12.8                                          # the 10 M body: 12.8 : 1 : 1 (1.6 body radii of length per GPU at N = 8)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1500.0, "bf16_tflops_sustained": 1500.0}, "fallback (B200_PROFILING.md)"


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


# ---------------------------------------------------------------------------------------------------- scenes
def sphere_on_obstacle(n, seed):
    """~n-particle dense sphere whose lowest particle starts 0.4 mm above the obstacle's plateau."""
    from meshless_inflatable_softbody_b200 import scenes
    x0, out_num = scenes.jittered_sphere(n, seed=seed)
    x0[:, 1] += (OBST_TOP + 0.0004) - x0[:, 1].min()
    return x0.astype(np.float32)


def body_scene(n_total, seed, aspect):
    """One elongated body (ellipsoid, long axis x) whose underside starts 0.4 mm above the obstacle's plateau.  The ground plane
    (sim.py:238-244) is 2 cm below: not reached within a bench run (2e-5 m per step)."""
    from meshless_inflatable_softbody_b200 import scenes
    x0 = scenes.jittered_ellipsoid(n_total, seed=seed, aspect=(aspect, 1.0, 1.0), low_drop=True)
    x0[:, 1] += (OBST_TOP + 0.0004) - x0[:, 1].min()
    return x0.astype(np.float32)


def obstacle_state():
    """Truncated-octahedron SDF in the reference's architecture: 9 weight-normalised Linear layers, width 1024
    (deepsdf.py:7,12-38).  Workload data (scenes.plateau_obstacle_state), not a checker."""
    from meshless_inflatable_softbody_b200 import scenes
    return scenes.plateau_obstacle_state(OBST_R, OBST_TOP, hidden=1024, n_linear=9)


def obstacle_bbox(cfg):
    from meshless_inflatable_softbody_b200 import scenes
    return scenes.plateau_obstacle_bbox(OBST_R, OBST_TOP, cfg.collision_range * np.sqrt(3.0) + 5e-4)   # band + a margin the body cannot cross in one step


def workload_config(mode, world, n_nominal, n_per_gpu, obstacle):
    """`config` of the JSON line.  Built from the mode and the sizes only, so the CUDA arm and the reference arm print the SAME dict."""
    obst = ("lands on a DeepSDF obstacle (9 x 1024 MLP, deepsdf.py:12-38, truncated-octahedron weights) + ground plane sim.py:238-244; "
            "timed with >= %d particles in the contact band" % BAND_TARGET) if obstacle else "dropped on the ground plane (sim.py:238-244), no obstacle"
    if mode == "strong":
        wl = ("BASELINE configs[4]: one ~%d-particle inflatable body (dense ellipsoid %.1f:1:1, reference defaults), %s; "
              "%s" % (n_nominal, BODY_ASPECT, obst,
                      "single domain on 1 GPU" if world == 1 else
                      "slab-partitioned across %d GPUs, halo exchange of ghost positions every step" % world))
    elif mode == "weak":
        wl = ("one elongated inflatable body of %d particles per GPU (%d GPUs, length grows with N), %s; slab-partitioned, halo exchange of "
              "ghost positions every step (BASELINE configs[4] mechanism at fixed per-GPU size)" % (n_per_gpu, world, obst))
    elif mode == "batch":
        wl = "BASELINE configs[3] shape: independent ~%d-particle inflatable spheres, %s; scenes sharded over %d GPU(s), no communication" % (n_per_gpu, obst, world)
    elif mode == "rebuild":
        wl = ("BASELINE configs[2]: ~%d-particle body, soft shell (x=+1) with an outward dead load on the shell, neighbour structure "
              "(Morton radix sort + cell table + exact lists) rebuilt every step, 1 GPU" % n_nominal)
    else:
        wl = "BASELINE configs[1]: ~%d-particle inflatable body (dense sphere, reference defaults), %s, 1 B200" % (n_nominal, obst)
    return {"workload": wl, "n_particles": int(n_nominal), "spacing_h": 0.5, "h": 0.007, "dt": 5e-5,
            "scene": "reference defaults E=1.5e5 nu=0.4 m=1e-4 x=-1, v0=(0,-0.4,0), f_ext=(0,-1e-3,0)", "mode": mode,
            "l2": "flushed between timed steps (256 MiB device write outside the event pairs); steady_state = un-flushed chained steps"}


# ---------------------------------------------------------------------------------------------------- reference arm
def oracle_for(x0, cfg, threads):
    from oracle import c_oracle as co           # bench.py's cpu_baseline / reference arm: allowed importer
    o = co.Oracle(x0, h=cfg.h, dt=cfg.time_step, damping=cfg.damping, k_col=cfg.collision_penalty_stiffness,
                  col_range=cfg.collision_range, threads=threads)
    o.set_threads(threads)                      # explicit: torchrun exports OMP_NUM_THREADS=1
    o.set_all_external_force(cfg.external_force); o.set_youngs_modulus(cfg.youngs_modulus)
    o.set_poisson_ratio(cfg.poisson_ratio); o.set_mass(cfg.mass); o.set_design(cfg.design_x)
    return o, co


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_faithful_rate(cfg, n_sample, steps, warmup, seed=0, with_obstacle=True):
    """Reference CPU path = oracle in FAITHFUL mode (27-cell walk, per-candidate svd3 + stress, 5 passes per step as
    sim.py:353-358), OpenMP over particles on ALL host threads (set explicitly, whatever OMP_NUM_THREADS says); the obstacle
    query is the oracle's fp32 MLP (torch CPU, all threads: what `sdf(points)` of sim.py:100 costs on a CPU) on the same
    fraction of particles the GPU run sends through the broad phase."""
    from meshless_inflatable_softbody_b200 import scenes
    cores = host_threads()
    x0, _ = scenes.jittered_sphere(n_sample, seed=seed, low_drop=True)
    o, co = oracle_for(x0, cfg, cores)
    o.startup(cfg.initial_velocity, mode=co.FAITHFUL)
    if warmup:
        o.step(warmup, mode=co.FAITHFUL)
    t = time.perf_counter()
    o.step(steps, mode=co.FAITHFUL)
    dt = time.perf_counter() - t
    dt_mlp = 0.0
    if with_obstacle:
        import torch
        from oracle import deepsdf_oracle as do
        torch.set_num_threads(cores)
        m = do.reference_like_module()
        m.load_state_dict({k: torch.as_tensor(v) for k, v in obstacle_state().items()})
        n_q = max(64, int(0.002 * len(x0)))                  # ~0.2 % of the particles pass the broad phase in the GPU run
        pts = torch.as_tensor(x0[:n_q])
        with torch.no_grad():
            m(pts)
            t = time.perf_counter()
            for _ in range(steps):
                m(pts)
            dt_mlp = time.perf_counter() - t
    return len(x0) * steps / (dt + dt_mlp), dt + dt_mlp, cores, len(x0), dt_mlp


def cpu_side_rates(cfg, n_small=3000, n_hoisted=20000, seed=0):
    """Two more CPU figures: the FAITHFUL path on ONE thread (what Warp's serial device='cpu' kernels would do) and the
    oracle's CACHED mode on all threads (R_j, S_j once per particle: the reference's redundant per-candidate SVDs hoisted), so that
    the GPU ratio is not read as a credit for removing that redundancy alone.  Small samples, one step each."""
    from meshless_inflatable_softbody_b200 import scenes
    out = {}
    x0, _ = scenes.jittered_sphere(n_small, seed=seed, low_drop=True)
    o, co = oracle_for(x0, cfg, threads=1)
    o.startup(cfg.initial_velocity, mode=co.FAITHFUL)
    t = time.perf_counter(); o.step(1, mode=co.FAITHFUL); dt = time.perf_counter() - t
    out["faithful_single_thread"] = {"value": len(x0) / dt, "unit": UNIT, "sample": f"1 step, {len(x0)} particles, 1 thread"}
    x0, _ = scenes.jittered_sphere(n_hoisted, seed=seed, low_drop=True)
    cores = host_threads()
    o, co = oracle_for(x0, cfg, cores)
    o.startup(cfg.initial_velocity)
    t = time.perf_counter(); o.step(2); dt = time.perf_counter() - t
    out["hoisted_all_threads"] = {"value": 2 * len(x0) / dt, "unit": UNIT, "sample": f"2 steps, {len(x0)} particles, {cores} threads, "
                                  "oracle CACHED mode (per-particle R_j, S_j; candidates with q >= 2 skipped)"}
    return out


def size_cpu_sample(cfg, n_full, total_steps, budget_s):
    """Particles the timed sample may hold so that `total_steps` steps take ~budget_s: calibrated on one step of a 16k-particle
    sphere after one warm-up step (a 4k-particle probe under-reads the rate 2-3x -- thread start-up, first-touch -- and would
    shrink the sample, and with it the CPU's rate, below what the budget allows)."""
    rate, dt, cores, n0, _ = cpu_faithful_rate(cfg, 16000, 1, 1, with_obstacle=False)
    n_fit = int(rate * budget_s / max(1, total_steps))
    return max(8000, min(n_full, n_fit)), rate


def resolve_mode(args, world):
    mode = args.mode
    if mode in ("configs1", "rebuild") and world > 1:
        raise SystemExit(f"--mode {mode} is a single-GPU workload")
    if mode == "strong":
        n_nominal, n_per = args.n_total, args.n_total // world
    elif mode == "weak":
        n_per = args.n or 1_250_000
        n_nominal = n_per * world
    elif mode == "batch":
        n_per = args.n or 10_000
        n_nominal = n_per * args.scenes * world
    elif mode == "rebuild":
        n_per = n_nominal = args.n or 1_000_000
    else:
        n_per = n_nominal = args.n or 100_000
    return mode, n_nominal, n_per


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    mode, n_nominal, n_per = resolve_mode(args, world)
    n_s, _ = size_cpu_sample(cfg, n_nominal, args.steps + args.warmup, budget_s=args.ref_budget)
    with_obst = not args.no_obstacle and mode != "rebuild" and (mode != "batch" or args.obstacle)
    rate, dt, cores, n_used, dt_mlp = cpu_faithful_rate(cfg, n_s, args.steps, args.warmup, with_obstacle=with_obst)
    sample = (f"{args.steps} steps (+{args.warmup} warm-up) of a {n_used}-particle dense sphere cut from the workload's material (same spacing, "
              f"kernel radius and parameters: ~235 neighbours/particle), oracle FAITHFUL mode (27-cell walk, per-candidate svd3) on {cores} "
              f"OpenMP threads" + (f" + the 9x1024 MLP on 0.2% of the particles per step (torch CPU fp32, {dt_mlp:.2f} s of the {dt:.2f} s)" if with_obst else "")
              + f"; the full workload ({n_nominal} particles) would take {n_nominal / rate:.0f} s per step on these cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps * (n_nominal / n_used),
        "higher_is_better": True, "scaling": "strong" if mode == "strong" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(mode, world, n_nominal, n_per, with_obst),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "sample_particles": n_used,
                         "sample_ms_per_step": 1e3 * dt / args.steps},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference has no CPU implementation (device='cuda' hard-coded, Warp/Taichi absent, SURVEY 8c): this arm times our CPU "
                "restatement of sim.py:133-258,341-358 (pinned to the reference's own source by tests/golden/sim_py_*.npz) and of "
                "deepsdf.py:9-41 on the host cores; particle-steps/s is size-independent on a CPU (work per particle is fixed by the "
                "neighbour count), ms_per_step is scaled to the full workload",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- our arm
class Harness:
    """Timing plumbing shared by every workload of the CUDA arm."""

    def __init__(self, dev, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.dev, self.world, self.local_rank = torch, dist, dev, world, local_rank
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v):
        if self.world > 1:
            t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def sum_over_ranks(self, vals):
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]


def enter_contact(stepper, core, hz, has_obstacle, any_obstacle, warmup, cfg, max_steps=600):
    """Untimed: restart the scene and advance it until the contact band holds >= BAND_TARGET particles (summed over ranks).
    Every rank takes the same number of steps (the decision is made on the all-reduced count)."""
    stepper.startup()
    stepper.step(max(3, warmup))
    done = max(3, warmup)
    band = 0.0
    if any_obstacle:
        while done < max_steps:
            c = core.contact_counts() if has_obstacle else (0, 0)
            band = hz.sum_over_ranks([c[1]])[0]
            if band >= BAND_TARGET:
                break
            stepper.step(8); done += 8
    core.synchronize()
    return done, band


def measure(args, cfg, hz, stepper, core, n_total, n_io, has_obstacle, any_obstacle, K, full=True):
    """Timed regions of one workload: A = K steps, L2 flushed between steps, per-step CUDA events (-> value); B = K chained
    steps (steady state); e2e through the host-buffer API (streamed and lock-step).  Returns a dict."""
    torch, dev = hz.torch, hz.dev
    cur = torch.cuda.current_stream(dev)
    out = {}
    pre_steps, band0 = enter_contact(stepper, core, hz, has_obstacle, any_obstacle, args.warmup, cfg)
    # ---- A
    launches0 = core.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    cand_sum = band_sum = 0
    hz.barrier()
    for a, b in evs:
        hz.flush.fill_(1)                  # not timed: sits between the previous end event and this start event
        core.stream.wait_stream(cur)
        with torch.cuda.stream(core.stream):
            a.record()
            stepper.step(1)
            b.record()
        cur.wait_stream(core.stream)
        if has_obstacle:                   # host read of the device counters of the step just timed (outside the event pair)
            c = core.contact_counts()
            cand_sum += c[0]; band_sum += c[1]
    hz.barrier()
    ms_own = sum(a.elapsed_time(b) for a, b in evs)
    if hz.world > 1:                        # every rank's own device time (the halo flag waits are inside it): who paces the step
        t = torch.zeros(hz.world, device=dev, dtype=torch.float64)
        t[hz.dist.get_rank()] = ms_own / K
        hz.dist.all_reduce(t, op=hz.dist.ReduceOp.SUM)
        out["ms_per_step_by_rank"] = [round(float(v), 4) for v in t.tolist()]
    ms_flushed = hz.max_over_ranks(ms_own)
    out["launches"] = core.launch_count - launches0
    cand_avg, band_avg = [v / K for v in hz.sum_over_ranks([cand_sum, band_sum])]
    out["contact"] = {"broad_phase_candidates_avg": cand_avg, "in_contact_band_avg": band_avg,
                      "pre_advance_steps": pre_steps, "in_contact_band_at_start": band0} if any_obstacle else None
    out["ms_flushed"] = ms_flushed
    out["value"] = n_total * K / (ms_flushed * 1e-3)
    x, v = core.position_velocity()
    out["state_finite"] = bool(torch.isfinite(x).all()) and bool(torch.isfinite(v).all())
    # ---- B
    enter_contact(stepper, core, hz, has_obstacle, any_obstacle, args.warmup, cfg)
    stepper.step(64)                       # untimed: the 32-step graph chunks are captured and instantiated on first use
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hz.barrier()
    with torch.cuda.stream(core.stream):
        e0.record(); stepper.step(K); e1.record()
    hz.barrier()
    ms_steady = hz.max_over_ranks(e0.elapsed_time(e1))
    out["steady_state"] = {"value": n_total * K / (ms_steady * 1e-3), "ms_per_step": ms_steady / K,
                           "note": "K chained steps (CUDA-graph chunks of 32), no L2 flush"}
    # ---- e2e: the public API with host buffers; per step H2D of the external-force field (the per-step input of the
    # reference API, sim.py:94,279-283) and D2H of position + velocity (sim.py:334,368-369)
    fext_host = torch.empty((n_io, 3), dtype=torch.float32).pin_memory()
    fext_host[:] = torch.tensor(cfg.external_force)
    x_host = [torch.empty((n_io, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    v_host = [torch.empty((n_io, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    Ke = max(1, min(K, args.e2e_steps))

    def timed(fn, finish):
        enter_contact(stepper, core, hz, has_obstacle, any_obstacle, args.warmup, cfg)
        for k in range(3):
            fn(k)
        finish()
        hz.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(core.stream):
            g0.record()
        for k in range(Ke):
            fn(k)
        finish()                           # every result is on the host
        with torch.cuda.stream(core.stream):
            g1.record()
        hz.barrier()
        return hz.max_over_ranks(max(g0.elapsed_time(g1), 0.0))

    # (a) streaming: the host hands over the force field of step k and receives the state of step k - 1 while step k runs
    # (uploads / downloads on the library's copy stream, double-buffered; the host still reads every step's result)
    def e2e_stream(k):
        stepper.set_external_forces_host(fext_host)
        stepper.step(1)
        core.get_state_host_async(x_host[k & 1], v_host[k & 1])
        core.wait_state_host(1)            # the state of step k - 1 has landed in x_host / v_host [(k - 1) & 1]
    ms_e2e = timed(e2e_stream, lambda: core.wait_state_host(0))
    out["e2e"] = {"value": n_total * Ke / (ms_e2e * 1e-3), "unit": UNIT, "steps": Ke,
                  "h2d_bytes_per_step": n_io * 12 * hz.world, "d2h_bytes_per_step": n_io * 24 * hz.world,
                  "what": "per step: mis_set_ext_force_host (pinned H2D of the force field) + step(1) + mis_get_state_host_async (x, v D2H to pinned "
                          "memory); transfers run on the library's copy stream and overlap the next step; the host waits for and owns the state "
                          "of step k-1 before it submits step k+1"}
    if full:
        # (b) lock-step: upload, step, download, host sync -- nothing overlaps
        def e2e_sync(k):
            stepper.set_external_forces_host(fext_host)
            stepper.step(1)
            core.get_state_host(x_host[0], v_host[0])
            core.synchronize()
        ms_sync = timed(e2e_sync, core.synchronize)
        out["e2e"]["lock_step"] = {"value": n_total * Ke / (ms_sync * 1e-3), "what": "same traffic, host sync after every step (no overlap)"}
    out["state_finite_at_end"] = bool(torch.isfinite(x_host[0]).all()) and bool(torch.isfinite(x_host[1]).all())
    return out


def kernel_roofline(core, info, Kp, peaks, peak_src):
    """CUDA events around every launch of the two gather kernels (library stream), L2 warm, obstacle removed."""
    hbm_peak = float(peaks["hbm_gbs"])
    ms_def, ms_for = core.profile_step(Kp)
    ms_def, ms_for = ms_def / Kp, ms_for / Kp
    nloc = core.n
    kern = {
        "k_force": {"ms": ms_for, "bytes_per_particle": BYTES_FORCE, "gbs": nloc * BYTES_FORCE / (ms_for * 1e-3) / 1e9,
                    "pairs_per_s": info.total_pairs / (ms_for * 1e-3),
                    "fp32_tflops_algorithmic": FLOP_PER_PAIR_FORCE * info.total_pairs / (ms_for * 1e-3) / 1e12},
        "k_deform": {"ms": ms_def, "bytes_per_particle": BYTES_DEFORM, "gbs": nloc * BYTES_DEFORM / (ms_def * 1e-3) / 1e9,
                     "pairs_per_s": 2 * info.total_pairs / (ms_def * 1e-3),
                     "fp32_tflops_algorithmic": FLOP_PER_PAIR_DEFORM * info.total_pairs / (ms_def * 1e-3) / 1e12},
    }
    dom = "k_force" if ms_for >= ms_def else "k_deform"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj, kname = json.load(open(tpath)), core.kernel_names()[dom]
            t = tj.get(kname) or tj.get(kname.replace("k_force_p", "k_force_c"), {})     # k_force_p streams k_force_c's lists and records
            per_particle = t.get("bytes_per_particle") or (t["bytes_per_launch_n100k"] / 99991.0 if "bytes_per_launch_n100k" in t else None)
            traffic = per_particle * nloc if per_particle else None      # ncu dram bytes per particle of the capture x particles of this launch
        except Exception:
            traffic = None
    names = core.kernel_names()
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": kern[dom]["gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kern[dom]["gbs"] / hbm_peak, "frac_of_spec_8000": kern[dom]["gbs"] / 8000.0, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": nloc * kern[dom]["bytes_per_particle"], "launch_ms": kern[dom]["ms"], "particles_per_launch": nloc,
                "note": "at ~240 neighbours/particle the gather kernels are FP32-pipe / shared-memory bound, not HBM bound (SURVEY 8d): "
                        "fp32_fraction (algorithmic flops / FFMA peak) is the binding figure",
                "fp32_peak_tflops": FP32_PEAK_TFLOPS, "fp32_fraction": kern[dom]["fp32_tflops_algorithmic"] / FP32_PEAK_TFLOPS,
                "kernels": {names[k]: v for k, v in kern.items()}}
    # BASELINE.json's metric names the force kernel's HBM fraction explicitly: always present, whichever kernel dominates
    roofline["force_kernel"] = {"kernel": names["k_force"], "achieved": kern["k_force"]["gbs"], "unit": "GB/s",
                                "frac": kern["k_force"]["gbs"] / hbm_peak, "frac_of_spec_8000": kern["k_force"]["gbs"] / 8000.0,
                                "algorithmic_bytes_per_particle": BYTES_FORCE,
                                "fp32_fraction": kern["k_force"]["fp32_tflops_algorithmic"] / FP32_PEAK_TFLOPS}
    return roofline


def slab_parity_check(args, cfg, hz, rank, world, dev):
    """Every N > 1 line: a >= 200k-particle body dropped on the ground plane (impact from step ~30, elastic waves cross every cut)
    through the SAME slab machinery, against a single-domain run on rank 0.  Tolerance: 4 x the single-domain fp32 summation-order
    floor (four cluster shapes / gather modes) + (4e-9, 2e-5)."""
    torch = hz.torch
    from meshless_inflatable_softbody_b200 import Simulator, scenes
    from meshless_inflatable_softbody_b200.slab import SlabSimulator
    steps = 80
    # a 2:1:1 ellipsoid dropped on the ground plane (impact from step ~30): compact enough to be stable at the reference defaults
    # (a 200k-particle body stretched to 12.8:1:1 has tips one lattice spacing wide -- rank-deficient moment matrices -- and
    # diverges within ~50 steps on any number of GPUs), long enough that every slab is thicker than its ghost depth
    n_par = max(args.parity_n, 50_000 * world)
    x0 = scenes.jittered_ellipsoid(n_par, seed=7, aspect=(2.0, 1.0, 1.0), low_drop=True).astype(np.float32)
    fext = np.tile(np.float32(cfg.external_force), (len(x0), 1))
    slab = SlabSimulator(x0, cfg, rank=rank, world_size=world, device=str(dev), halo=args.halo)
    slab.sim.set_external_forces(fext[slab.plan.local_ids])
    slab.startup(); slab.step(steps)
    X, V = slab.gather_global()
    ok_halo = slab.halo_ok() if slab.halo == "p2p" else True
    res = None
    if rank == 0:
        # the fp32 summation-order floor of THIS scene: four single-domain runs that differ only in the order of the neighbour sums
        # (cluster shape, lanes per cluster, gather mode); the slab run differs from them in the same way (other cell grid origin,
        # other list order near every cut)
        variants = [dict(), dict(cluster_size=4, lanes_per_particle=16, mode=0), dict(cluster_size=1, lanes_per_particle=8, mode=1),
                    dict(cluster_size=2, lanes_per_particle=32, mode=0)]
        states = []
        for kw in variants:
            kw = dict(kw)
            mode = kw.pop("mode", None)
            s_ = Simulator(x0, cfg, device=str(dev), **kw)
            if mode is not None:
                s_.set_gather_mode(mode)
            s_.set_external_forces(fext); s_.startup(); s_.step(steps)
            states.append(s_.position_velocity())
            s_.close()
        xa, va = states[0]
        fx = max(float((states[i][0] - states[j][0]).abs().max()) for i in range(4) for j in range(i))
        fv = max(float((states[i][1] - states[j][1]).abs().max()) for i in range(4) for j in range(i))
        dx, dv = float((X - xa).abs().max()), float((V - va).abs().max())
        t = steps * cfg.time_step
        elastic = float((va[:, 1] - (cfg.initial_velocity[1] + cfg.external_force[1] / cfg.mass * t)).abs().max())   # departure from free fall
        finite = bool(torch.isfinite(X).all() and torch.isfinite(xa).all())
        res = {"n_particles": len(x0), "steps": steps, "max_abs_dx": dx, "max_abs_dv": dv, "floor_dx": fx, "floor_dv": fv,
               "rule": "|dx| <= 4 floor_dx + 4e-9, |dv| <= 4 floor_dv + 2e-5 (floor = largest difference between four single-domain runs with other cluster shapes / gather modes)",
               "elastic_velocity_change": elastic, "halo": slab.halo,
               "ok": bool(finite and elastic > 1e-3 and dx <= 4 * fx + 4e-9 and dv <= 4 * fv + 2e-5 and ok_halo)}
    slab.close()
    hz.barrier()
    return res


def run_ours(args, cfg, rank, world, local_rank):
    import torch
    from meshless_inflatable_softbody_b200 import Simulator, DeepSDF, scenes
    from meshless_inflatable_softbody_b200.slab import SlabSimulator

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    hz = Harness(dev, world, local_rank)
    mode, n_nominal, n_per = resolve_mode(args, world)
    # configs[3] (batched independent scenes) uses the reference's own contact, the ground plane (sim.py:238-244), unless --obstacle
    # is given: every scene would otherwise run its own cooperative MLP chain per step and the chains of the scenes serialise
    use_obstacle = not args.no_obstacle and mode != "rebuild" and (mode != "batch" or args.obstacle)
    K = args.steps
    peaks, peak_src = measured_peaks()
    sim_kw = dict(lanes_per_particle=args.lanes, cluster_size=args.cluster)
    extra = {}
    parity = None

    if mode == "batch":
        return run_batch(args, cfg, hz, rank, world, dev, n_per, n_nominal, use_obstacle, peaks, peak_src)
    if mode == "rebuild":
        return run_rebuild(args, cfg, hz, dev, n_per, peaks, peak_src)

    if mode in ("strong", "weak"):
        aspect = BODY_ASPECT if mode == "strong" else 1.6 * world
        x0 = body_scene(n_nominal, seed=0, aspect=aspect)
    else:
        x0 = sphere_on_obstacle(n_per, seed=0)
    n_total = len(x0)
    if world > 1:
        parity = slab_parity_check(args, cfg, hz, rank, world, dev)
        extra_cost = None
        weights = None
        if args.balance == "pairs":         # cut by estimated pair work (27-cell occupancy), not by particle count
            from meshless_inflatable_softbody_b200.slab import neighbour_weights
            weights = neighbour_weights(x0, cfg.h)
        if use_obstacle and args.contact_cost > 0:
            # static load balance: the ranks whose slab lies under the obstacle run its MLP query every step (a latency-bound
            # ~0.2 ms that does not shrink with the slab); they own that many particle-equivalents fewer
            from meshless_inflatable_softbody_b200.slab import SlabPartition
            cuts = SlabPartition.build(x0, cfg.h, world, weights=weights).cuts
            bb = np.asarray(obstacle_bbox(cfg), np.float64).reshape(2, 3)
            lo, hi = bb[0, 0] - 0.03, bb[1, 0] + 0.03
            extra_cost = np.array([args.contact_cost if (cuts[r] <= hi and cuts[r + 1] >= lo) else 0.0 for r in range(world)])
        stepper = SlabSimulator(x0, cfg, rank=rank, world_size=world, device=str(dev), halo=args.halo, extra_cost=extra_cost, weights=weights, **sim_kw)
        core = stepper.sim
        extra = {"partition_extra_cost_particles": extra_cost.tolist() if extra_cost is not None else None, "partition_balance": args.balance,
                 "halo": ("fused P2P push over NVLink peer memory from the force kernel's epilogue + epoch flags, inside the step graph"
                          if stepper.halo == "p2p" else "NCCL send/recv after every step"),
                 "owned_per_gpu": stepper.n_owned, "ghosts_per_gpu": core.n - stepper.n_owned,
                 "halo_bytes_per_step_per_gpu": (16 if stepper.halo == "p2p" else 12) * int(sum(len(v) for v in stepper.plan.send.values()))}
        near = stepper.obstacle_nearby(np.asarray(obstacle_bbox(cfg)), margin=0.03) if use_obstacle else False
    else:
        stepper = core = Simulator(x0, cfg, device=str(dev), **sim_kw)
        near = use_obstacle
    del x0
    net = None
    if near:
        net = DeepSDF(obstacle_state(), device=str(dev))
        stepper.set_sdf_obstacle(net, bbox_model=obstacle_bbox(cfg), fd_eps=1e-4)
    if use_obstacle:
        extra["ranks_with_obstacle_query"] = int(hz.sum_over_ranks([1 if near else 0])[0])
    info = core.neighbor_info()
    pairs_total, n_local_sum = hz.sum_over_ranks([info.total_pairs, core.n])
    mean_k = pairs_total / n_local_sum

    sampler = ClockSampler(local_rank)
    sampler.start()
    m = measure(args, cfg, hz, stepper, core, n_total, core.n, near, use_obstacle, K, full=True)
    clocks = sampler.stop()
    if mode in ("strong", "weak") and world > 1:
        assert stepper.halo_ok(), "a halo flag wait timed out"

    # ---- roofline of the dominant kernel
    if net:
        stepper.set_sdf_obstacle(None, None)           # per-kernel timing of the two gather kernels alone
    roofline = kernel_roofline(core, info, min(K, 50), peaks, peak_src)
    if world > 1:                           # every rank's gather-kernel time per step (no obstacle, no waits): the partition's balance
        t = torch.zeros((world, 3), device=dev, dtype=torch.float64)
        kk = roofline["kernels"]
        names = core.kernel_names()
        t[rank, 0] = kk[names["k_deform"]]["ms"]; t[rank, 1] = kk[names["k_force"]]["ms"]; t[rank, 2] = 1.0 if near else 0.0
        hz.dist.all_reduce(t, op=hz.dist.ReduceOp.SUM)
        extra["gather_ms_by_rank"] = [{"deform": round(float(a), 4), "force": round(float(b), 4), "obstacle": bool(c)} for a, b, c in t.tolist()]
    if net:
        m_rows = 16384
        ms_gemm = net.profile_gemm(m_rows, reps=20)
        flop = 2.0 * m_rows * 1024 * 1024
        tf32_peak = float(peaks.get("bf16_tflops", 1500.0)) / 2.0
        roofline["sdf_mlp"] = {"kernel": "k_sdf_gemm (tcgen05.mma kind::tf32, 3 products per fp32-accurate product)", "rows": m_rows,
                               "ms_per_layer": ms_gemm, "tflops_fp32_equivalent": flop / ms_gemm / 1e9, "tflops_tf32_issued": 3 * flop / ms_gemm / 1e9,
                               "tf32_peak_tflops": tf32_peak, "tensor_frac": 3 * flop / ms_gemm / 1e9 / tf32_peak,
                               "peak_note": "TF32 dense peak taken as half the measured bf16 burst peak (" + peak_src + ")"}
    value = m["value"]
    cfg_line = workload_config(mode, world, n_nominal, n_per, use_obstacle)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": m["ms_flushed"] / K, "higher_is_better": True, "scaling": "strong" if mode == "strong" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_line,
        "n_particles_actual": n_total, "mean_neighbors": mean_k, "pairs_per_sec": value * mean_k,
        "contact": m["contact"], "steady_state": m["steady_state"], "e2e": m["e2e"],
        "gpu_launches": int(m["launches"]), "clocks": clocks, "roofline": roofline,
        "state_finite": m["state_finite"], "state_finite_at_end": m["state_finite_at_end"],
    }
    if "ms_per_step_by_rank" in m:
        line["ms_per_step_by_rank"] = m["ms_per_step_by_rank"]
    line.update(extra)
    if world > 1:
        line["parity_check"] = parity
    stepper.close()
    del stepper, core
    torch.cuda.empty_cache()

    if world == 1 and mode == "strong" and not args.no_configs1:
        # ---- sub-record: BASELINE configs[1] (100k sphere + obstacle) on this GPU
        x1 = sphere_on_obstacle(args.n or 100_000, seed=0)
        s1 = Simulator(x1, cfg, device=str(dev), **sim_kw)
        net1 = None
        if use_obstacle:
            net1 = DeepSDF(obstacle_state(), device=str(dev))
            s1.set_sdf_obstacle(net1, bbox_model=obstacle_bbox(cfg), fd_eps=1e-4)
        i1 = s1.neighbor_info()
        K1 = max(K, 100)
        m1 = measure(args, cfg, hz, s1, s1, len(x1), s1.n, use_obstacle, use_obstacle, K1, full=False)
        if net1:
            s1.set_sdf_obstacle(None, None)
        r1 = kernel_roofline(s1, i1, 50, peaks, peak_src)
        line["configs1"] = {"config": workload_config("configs1", 1, args.n or 100_000, args.n or 100_000, use_obstacle),
                            "n_particles_actual": len(x1), "mean_neighbors": i1.total_pairs / s1.n, "steps": K1,
                            "value": m1["value"], "unit": UNIT, "ms_per_step": m1["ms_flushed"] / K1, "contact": m1["contact"],
                            "steady_state": m1["steady_state"], "e2e": m1["e2e"], "gpu_launches": int(m1["launches"]),
                            "state_finite": m1["state_finite"], "roofline": r1}
        s1.close()

    if rank == 0 and world == 1 and not args.no_cpu:
        n_s, _ = size_cpu_sample(cfg, n_total, 1, budget_s=args.cpu_budget)
        rate, dt, cores, n_used, dt_mlp = cpu_faithful_rate(cfg, n_s, 1, 0, with_obstacle=use_obstacle)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 step of a {n_used}-particle dense sphere of the workload's material (same spacing/params; {dt:.1f} s, of which {dt_mlp:.2f} s the "
                      f"9x1024 MLP on 0.2% of the particles), oracle FAITHFUL mode (27-cell walk, per-candidate svd3, sim.py:353-358), {cores} OpenMP threads"}
        try:
            line["cpu_baseline"]["other_modes"] = cpu_side_rates(cfg)
        except Exception as e:                       # a reported extra, never a reason to lose the bench line
            line["cpu_baseline"]["other_modes"] = {"error": str(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_batch(args, cfg, hz, rank, world, dev, n_per, n_nominal, use_obstacle, peaks, peak_src):
    """configs[3]: args.scenes independent scenes per GPU, each on its own stream, no communication."""
    torch = hz.torch
    from meshless_inflatable_softbody_b200 import Simulator, DeepSDF
    S, K = args.scenes, args.steps
    sims, nets, n_sum = [], [], 0
    from meshless_inflatable_softbody_b200 import scenes
    for k in range(S):
        if use_obstacle:
            x0 = sphere_on_obstacle(n_per, seed=rank * S + k)
        else:
            x0, _ = scenes.jittered_sphere(n_per, seed=rank * S + k, low_drop=True)      # BASELINE configs[0] scene: impact on the ground plane from step ~30
        s = Simulator(x0, cfg, device=str(dev), lanes_per_particle=args.lanes, cluster_size=args.cluster)
        if use_obstacle:
            nets.append(DeepSDF(obstacle_state(), device=str(dev)))      # one network per scene: activations are per-network scratch
            s.set_sdf_obstacle(nets[-1], bbox_model=obstacle_bbox(cfg), fd_eps=1e-4)
        s.startup(); s.step(max(3, args.warmup))
        sims.append(s); n_sum += len(x0)
    n_total = hz.sum_over_ranks([n_sum])[0]
    for s in sims:
        s.step(64); s.synchronize()
    cur = torch.cuda.current_stream(dev)
    sampler = ClockSampler(hz.local_rank); sampler.start()
    l0 = sum(s.launch_count for s in sims)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hz.barrier()
    hz.flush.fill_(1)
    e0.record()
    for s in sims:
        s.stream.wait_stream(cur)
    for s in sims:
        s.step(K)
    for s in sims:
        cur.wait_stream(s.stream)
    e1.record()
    hz.barrier()
    ms = hz.max_over_ranks(e0.elapsed_time(e1))
    launches = sum(s.launch_count for s in sims) - l0
    clocks = sampler.stop()
    # e2e: every scene uploads its force field and downloads its state every step
    Ke = max(1, min(K, args.e2e_steps))
    bufs = []
    for s in sims:
        f = torch.empty((s.n, 3), dtype=torch.float32).pin_memory(); f[:] = torch.tensor(cfg.external_force)
        bufs.append((f, [torch.empty((s.n, 3), dtype=torch.float32).pin_memory() for _ in range(2)],
                     [torch.empty((s.n, 3), dtype=torch.float32).pin_memory() for _ in range(2)]))
    def sweep(k):
        for s, (f, xs, vs) in zip(sims, bufs):
            s.set_external_forces_host(f); s.step(1); s.get_state_host_async(xs[k & 1], vs[k & 1])
        for s in sims:
            s.wait_state_host(1)
    for k in range(3):
        sweep(k)
    for s in sims:
        s.wait_state_host(0)
    hz.barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for k in range(Ke):
        sweep(k)
    for s in sims:
        s.wait_state_host(0); cur.wait_stream(s.stream)
    g1.record()
    hz.barrier()
    ms_e2e = hz.max_over_ranks(max(g0.elapsed_time(g1), 1e3 * (time.perf_counter() - t0)))
    finite = all(bool(torch.isfinite(b[1][0]).all()) for b in bufs)
    value = n_total * K / (ms * 1e-3)
    info = sims[0].neighbor_info()
    if use_obstacle:
        for s in sims:
            s.set_sdf_obstacle(None, None)
    roofline = kernel_roofline(sims[0], info, min(K, 50), peaks, peak_src)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config("batch", world, n_nominal, n_per, use_obstacle), "n_particles_actual": int(n_total),
            "scenes_per_gpu": S, "scenes_total": S * world, "mean_neighbors": info.total_pairs / sims[0].n,
            "e2e": {"value": n_total * Ke / (ms_e2e * 1e-3), "unit": UNIT, "steps": Ke, "h2d_bytes_per_step": int(n_total) * 12,
                    "d2h_bytes_per_step": int(n_total) * 24, "what": "per scene and step: force field H2D, step(1), streamed x, v D2H"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "state_finite": finite,
            "timing": "one CUDA-event pair around K chained steps of all scenes of a GPU running concurrently on their own streams (L2 flushed once before); max over ranks"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    for s in sims:
        s.close()


def run_rebuild(args, cfg, hz, dev, n, peaks, peak_src):
    """configs[2]: soft shell + outward dead load on the shell, full neighbour rebuild (sort + cell table + lists) every step."""
    torch = hz.torch
    from meshless_inflatable_softbody_b200 import Simulator, scenes
    K = args.steps
    x0, out_num = scenes.jittered_sphere(n, seed=0, centre=(0.0, 0.3, 0.0))
    sim = Simulator(x0, cfg, device=str(dev), lanes_per_particle=args.lanes, cluster_size=args.cluster)
    shell = scenes.shell_mask(x0, cfg.h)
    sim.set_design(np.where(shell, 1.0, -1.0).astype(np.float32))
    c = x0.mean(0)
    rad = (x0 - c) / np.maximum(np.linalg.norm(x0 - c, axis=1, keepdims=True), 1e-9)
    f = np.tile(np.float32(cfg.external_force), (len(x0), 1)); f[shell] += (2e-3 * rad[shell]).astype(np.float32)
    sim.set_external_forces(f)
    sim.startup(); sim.step(max(3, args.warmup)); sim.synchronize()
    cur = torch.cuda.current_stream(dev)
    sampler = ClockSampler(hz.local_rank); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l0 = sim.launch_count
    hz.barrier()
    for a, b, c2 in ev:
        hz.flush.fill_(1)
        sim.stream.wait_stream(cur)
        with torch.cuda.stream(sim.stream):
            a.record(); sim.rebuild_neighbors(); b.record(); sim.step(1); c2.record()
        cur.wait_stream(sim.stream)
    hz.barrier()
    ms_build = sum(a.elapsed_time(b) for a, b, _ in ev); ms_step = sum(b.elapsed_time(c2) for _, b, c2 in ev)
    clocks = sampler.stop()
    x, v = sim.position_velocity()
    info = sim.neighbor_info()
    roofline = kernel_roofline(sim, info, min(K, 20), peaks, peak_src)
    value = len(x0) * K / ((ms_build + ms_step) * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": args.warmup, "ms_per_step": (ms_build + ms_step) / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config("rebuild", 1, n, n, False), "n_particles_actual": len(x0), "mean_neighbors": info.total_pairs / sim.n,
            "rebuild_ms": ms_build / K, "step_ms": ms_step / K, "gpu_launches": int(sim.launch_count - l0), "clocks": clocks, "roofline": roofline,
            "state_finite": bool(torch.isfinite(x).all()) and bool(torch.isfinite(v).all()),
            "note": "the rebuild includes two host round trips (list sizes) and re-primes the step graph; queries are on x0, so the rebuilt lists are identical"}
    print(json.dumps(line), flush=True)
    sim.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="strong", choices=["strong", "weak", "batch", "rebuild", "configs1"])
    ap.add_argument("--n-total", type=int, default=10_000_000, help="total particles of the --mode strong body (BASELINE configs[4])")
    ap.add_argument("--n", "--particles", dest="n", type=int, default=0,
                    help="(use --particles under torchrun, whose own parser finds --n ambiguous) particles per GPU for --mode weak (1.25 M) / per scene "
                         "for --mode batch (10 k) / of the configs[1] scene (100 k) / of --mode rebuild (1 M)")
    ap.add_argument("--scenes", type=int, default=8, help="--mode batch: scenes per GPU (8 x 8 GPUs = the 64 scenes of configs[3])")
    ap.add_argument("--halo", default="auto", choices=["auto", "p2p", "nccl"], help="slab modes: ghost exchange mechanism")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per cluster (0 = library default)")
    ap.add_argument("--cluster", type=int, default=0, help="particles per cluster (0 = library default)")
    ap.add_argument("--no-obstacle", action="store_true", help="ground-plane contact only")
    ap.add_argument("--obstacle", action="store_true", help="--mode batch: give every scene the DeepSDF obstacle too (one MLP chain per scene and step)")
    ap.add_argument("--no-configs1", action="store_true", help="N = 1: skip the configs[1] sub-record")
    ap.add_argument("--contact-cost", type=float, default=100_000.0,
                    help="N > 1: per-step cost of the obstacle query in particle-equivalents (0.2 ms at 1.9 ns per particle-step); the ranks under "
                         "the obstacle own that many particles fewer (0 = equal counts)")
    ap.add_argument("--balance", default="pairs", choices=["pairs", "count"],
                    help="N > 1: slab cuts equalise the estimated pair work (27-cell occupancy per particle) or the particle count")
    ap.add_argument("--parity-n", type=int, default=200_000, help="N > 1: particles of the slab-vs-single-domain parity check")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=60.0, help="--impl reference: seconds of CPU work the sample is sized for")
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
