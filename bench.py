#!/usr/bin/env python
"""bench.py -- particle-steps/s of the meshless inflatable soft-body step on B200.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path of this repo
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU arm (oracle port on the host cores)

One bench "step" = one simulation step (loop body sim.py:352-358: part_1, compute_A_pq, compute_nabla_u,
compute_elastic_forces, part_2) over all particles, plus the per-step obstacle-contact query of the workload.

Workloads
  N = 1 : BASELINE.json configs[1] -- ~100k-particle inflatable body (dense jittered sphere, reference defaults) dropped on
          a DeepSDF-encoded obstacle (the reference's 9 x 1024 architecture, deepsdf.py:12-38, with analytic octahedron
          weights) standing on the ground plane; contact = ground penalty (sim.py:238-244) + SDF penalty (extension).
  N > 1 : one scene of N x (--n) particles (default 1.25 M per GPU: N = 8 is the 10M-particle scene of configs[4]; an
          elongated body, long axis x) slab-partitioned across the N GPUs with a per-step
          halo exchange of the ghost particles' new positions (slab.py: fused P2P push over NVLink peer memory, or NCCL
          send/recv with --halo nccl); per-GPU work is fixed => "scaling": "weak".
          --mode batch runs independent scenes, one per GPU (configs[3]); --mode strong fixes the total particle count.

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
R_OCT = 0.012                # obstacle "radius" (m): |x|+|y|+|z| = R_OCT


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
    """~n-particle dense sphere whose lowest particle starts 0.4 mm above the tip of the octahedron obstacle."""
    from meshless_inflatable_softbody_b200 import scenes
    x0, out_num = scenes.jittered_sphere(n, seed=seed)
    x0[:, 1] += (R_OCT + 0.0004) - x0[:, 1].min()
    return x0.astype(np.float32)


def beam_scene(n_total, seed, world):
    """One elongated body (ellipsoid, long axis x, ~1.6 radii of length per GPU) in free fall, 0.3 m above the ground plane: no
    ground contact within a few thousand steps.  (At 1.25 M particles per GPU the body is 0.4 m thick; dropped onto the plane at the
    reference defaults it diverges ~800 steps after the impact -- on one GPU as well, scripts/beam_stability.py -- because the bottom
    layer alone carries the penalty force of the whole column.  The step's work does not depend on the state: the neighbour lists are
    static and no kernel exits early.)"""
    from meshless_inflatable_softbody_b200 import scenes
    x0 = scenes.jittered_ellipsoid(n_total, seed=seed, aspect=(1.6 * max(world, 1), 1.0, 1.0), low_drop=True)
    x0[:, 1] += 0.3
    return x0.astype(np.float32)


def obstacle_state():
    """Octahedron SDF (|x|+|y|+|z| - r)/sqrt(3) in the reference's architecture: 9 weight-normalised Linear layers,
    width 1024 (deepsdf.py:7,12-38).  Built here (not imported from oracle/) -- it is workload data, not a checker."""
    hidden, n_linear = 1024, 9
    st = {}
    dims = [3] + [hidden] * (n_linear - 1) + [1]
    for l in range(n_linear):
        o, i = dims[l + 1], dims[l]
        v = np.zeros((o, i), np.float32); b = np.zeros(o, np.float32)
        if l == 0:
            for a in range(3):
                v[2 * a, a] = 1.0; v[2 * a + 1, a] = -1.0
            v[6:, 0] = 1.0
            g = np.zeros((o, 1), np.float32); g[:6] = 1.0
        elif l < n_linear - 1:
            for u in range(6):
                v[u, u] = 1.0
            v[6:, 0] = 1.0
            g = np.zeros((o, 1), np.float32); g[:6] = 1.0
        else:
            v[0, :6] = 1.0
            g = np.full((1, 1), np.sqrt(6.0) / np.sqrt(3.0), np.float32)
            b[0] = -R_OCT / np.sqrt(3.0)
        st[f"network.{3 * l}.parametrizations.weight.original0"] = g
        st[f"network.{3 * l}.parametrizations.weight.original1"] = v
        st[f"network.{3 * l}.bias"] = b
    return st


def obstacle_bbox(cfg):
    m = cfg.collision_range * np.sqrt(3.0) + 5e-4            # band + a margin the body cannot cross in one step
    return [-R_OCT - m] * 3 + [R_OCT + m] * 3


def workload_config(args, n_total, mean_k, world, mode, extra=None):
    if world == 1:
        wl = ("BASELINE configs[1]: ~100k-particle inflatable body with DeepSDF obstacle contact on 1 B200 (dense sphere, "
              "reference defaults, 9x1024 DeepSDF octahedron obstacle + ground plane sim.py:238-244)")
    elif mode == "batch":
        wl = "BASELINE configs[3] shape: independent ~%d-particle scenes, one per GPU, no communication" % args.n
    elif mode == "strong":
        wl = "BASELINE configs[4] shape: one %d-particle elongated body slab-partitioned across %d GPUs, halo exchange of ghost positions every step" % (n_total, world)
    else:
        wl = ("one %d-particle elongated body (%d per GPU) in free fall, slab-partitioned across %d GPUs, halo exchange of ghost positions "
              "every step (BASELINE configs[4] mechanism at fixed per-GPU size)" % (n_total, args.n, world))
    c = {"workload": wl, "n_particles": int(n_total), "mean_neighbors": mean_k, "spacing_h": 0.5, "h": 0.007, "dt": 5e-5,
         "scene": "reference defaults E=1.5e5 nu=0.4 m=1e-4 x=-1, v0=(0,-0.4,0)", "mode": mode,
         "l2": "flushed between timed steps (256 MiB device write outside the event pairs); steady_state = un-flushed chained steps"}
    if extra:
        c.update(extra)
    if world > 1 and mode == "slab":
        # the N = 1 bench line is configs[1] (100k particles + obstacle), not this workload: point at the measured single-GPU
        # throughput of ONE GPU's share of this scene (same per-GPU particle count, no partition) as the weak-scaling denominator
        try:
            for l in open(os.path.join(ROOT, "profiles", "r01_scaling.jsonl")):
                d = json.loads(l)
                if d.get("run") == "n1_1250k" and abs(d["config"]["n_particles"] - args.n) < 0.01 * args.n:
                    c["single_gpu_same_share"] = {"value": d["value"], "unit": UNIT, "n_particles": d["config"]["n_particles"],
                                                  "source": "profiles/r01_scaling.jsonl run n1_1250k (bench.py --n 1250000 --no-obstacle)"}
        except Exception:
            pass
    return c


# ---------------------------------------------------------------------------------------------------- reference arm
def oracle_for(x0, cfg, threads=0):
    from oracle import c_oracle as co           # bench.py's cpu_baseline / reference arm: allowed importer
    o = co.Oracle(x0, h=cfg.h, dt=cfg.time_step, damping=cfg.damping, k_col=cfg.collision_penalty_stiffness,
                  col_range=cfg.collision_range, threads=threads)
    o.set_all_external_force(cfg.external_force); o.set_youngs_modulus(cfg.youngs_modulus)
    o.set_poisson_ratio(cfg.poisson_ratio); o.set_mass(cfg.mass); o.set_design(cfg.design_x)
    return o, co


def cpu_faithful_rate(cfg, n_sample, steps, warmup, seed=0, with_obstacle=True):
    """Reference CPU path = oracle in FAITHFUL mode (27-cell walk, per-candidate svd3 + stress, 5 passes per step as
    sim.py:353-358), OpenMP over particles on all host threads; the obstacle query is the oracle's fp32 MLP
    (torch CPU, all threads: what `sdf(points)` of sim.py:100 costs on a CPU) on the same fraction of particles
    the GPU run sends through the broad phase."""
    from meshless_inflatable_softbody_b200 import scenes
    x0, _ = scenes.jittered_sphere(n_sample, seed=seed, low_drop=True)
    o, co = oracle_for(x0, cfg)
    cores = co.max_threads()
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
        m = do.reference_like_module()
        m.load_state_dict({k: torch.as_tensor(v) for k, v in obstacle_state().items()})
        n_q = max(64, int(0.01 * len(x0)))                   # ~1 % of the particles pass the broad phase in the GPU run
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
    o, co = oracle_for(x0, cfg)
    o.startup(cfg.initial_velocity)
    t = time.perf_counter(); o.step(2); dt = time.perf_counter() - t
    out["hoisted_all_threads"] = {"value": 2 * len(x0) / dt, "unit": UNIT, "sample": f"2 steps, {len(x0)} particles, {co.max_threads()} threads, "
                                  "oracle CACHED mode (per-particle R_j, S_j; candidates with q >= 2 skipped)"}
    return out


def size_cpu_sample(cfg, n_full, total_steps, budget_s):
    rate, dt, cores, n0, _ = cpu_faithful_rate(cfg, 4000, 1, 0, with_obstacle=False)
    n_fit = int(rate * budget_s / max(1, total_steps))
    return max(2000, min(n_full, n_fit)), rate


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    n_s, _ = size_cpu_sample(cfg, args.n, args.steps + args.warmup, budget_s=args.ref_budget)
    rate, dt, cores, n_used, dt_mlp = cpu_faithful_rate(cfg, n_s, args.steps, args.warmup)
    sample = (f"{args.steps} steps (+{args.warmup} warm-up) of a {n_used}-particle sphere, same spacing/params as the GPU workload, "
              f"oracle FAITHFUL mode (27-cell walk, per-candidate svd3) on {cores} OpenMP threads + the 9x1024 MLP on 1% of the "
              f"particles per step (torch CPU fp32, {dt_mlp:.2f} s of the {dt:.2f} s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_used, None, 1, "single"),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference has no CPU implementation (device='cuda' hard-coded, Warp/Taichi absent, SURVEY 8c): this arm "
                "times our CPU restatement of sim.py:133-258,341-358 and deepsdf.py:9-41 on the host cores",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- our arm
def run_ours(args, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from meshless_inflatable_softbody_b200 import Simulator, DeepSDF
    from meshless_inflatable_softbody_b200.slab import SlabSimulator

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    mode = args.mode if world > 1 else "single"
    net = None
    if mode == "single" or mode == "batch":
        x0 = sphere_on_obstacle(args.n, seed=rank)
        sim = Simulator(x0, cfg, device=str(dev), lanes_per_particle=args.lanes, cluster_size=args.cluster)
        if not args.no_obstacle:
            net = DeepSDF(obstacle_state(), device=str(dev))
            sim.set_sdf_obstacle(net, bbox_model=obstacle_bbox(cfg), fd_eps=1e-4)
        n_local = n_total_local = len(x0)
        n_total = n_local * world
        core, stepper = sim, sim
        extra = {}
    else:
        n_total = args.n * world if mode == "slab" else args.n_total
        x0 = beam_scene(n_total, seed=0, world=world)
        n_total = len(x0)
        stepper = SlabSimulator(x0, cfg, rank=rank, world_size=world, device=str(dev), halo=args.halo,
                                lanes_per_particle=args.lanes, cluster_size=args.cluster)
        core = stepper.sim
        n_local, n_total_local = stepper.n_owned, core.n
        extra = {"halo": ("fused P2P push over NVLink peer memory from the force kernel's epilogue + epoch flags, inside the step graph"
                          if stepper.halo == "p2p" else "NCCL send/recv after every step"),
                 "owned_per_gpu": n_local, "ghosts_per_gpu": n_total_local - n_local,
                 "halo_bytes_per_step_per_gpu": (16 if stepper.halo == "p2p" else 12) * int(sum(len(v) for v in stepper.plan.send.values()))}
    info = core.neighbor_info()
    mean_k = info.total_pairs / core.n
    stepper.startup()
    stepper.step(args.warmup)
    core.synchronize()

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
    cur = torch.cuda.current_stream(dev)

    # ---- timed region A (`value`): K steps, L2 flushed between steps, one CUDA-event pair per step.
    # Events are recorded on the current stream, which waits for / is waited on by the library's stream around each step.
    sampler = ClockSampler(local_rank)
    launches0 = core.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    sampler.start()
    for a, b in evs:
        flush.fill_(1)                     # not timed: sits between the previous end event and this start event
        core.stream.wait_stream(cur)
        with torch.cuda.stream(core.stream):
            a.record()
            stepper.step(1)
            b.record()
        cur.wait_stream(core.stream)
    barrier()
    ms_flushed = sum(a.elapsed_time(b) for a, b in evs)
    launches = core.launch_count - launches0      # includes the MLP chain launches enqueued by the step
    counts = core.contact_counts() if net else (0, 0)
    # ---- timed region B (steady state): K chained steps, L2 warm
    stepper.step(64)                       # untimed: the 32-step graph chunks are captured and instantiated on first use
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(core.stream):
        e0.record(); stepper.step(K); e1.record()
    barrier()
    ms_steady = e0.elapsed_time(e1)
    clocks = sampler.stop()
    ms_flushed, ms_steady = max_over_ranks(ms_flushed), max_over_ranks(ms_steady)

    # ---- e2e: the public API with host buffers; per step H2D of the external-force field (the per-step input of the
    # reference API, sim.py:94,279-283) and D2H of position + velocity (sim.py:334,368-369), host sync every step
    n_io = core.n
    fext_host = torch.empty((n_io, 3), dtype=torch.float32).pin_memory()
    fext_host[:] = torch.tensor(cfg.external_force)
    x_host = [torch.empty((n_io, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    v_host = [torch.empty((n_io, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    Ke = max(1, min(K, args.e2e_steps))

    def timed(fn, finish):
        for k in range(3):
            fn(k)
        finish()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(core.stream):
            g0.record()
        for k in range(Ke):
            fn(k)
        finish()                           # every result is on the host
        with torch.cuda.stream(core.stream):
            g1.record()
        barrier()
        return max_over_ranks(max(g0.elapsed_time(g1), 0.0))

    # (a) streaming: the host hands over the force field of step k and receives the state of step k - 1 while step k runs
    # (uploads / downloads on the library's copy stream, double-buffered; the host still reads every step's result)
    def e2e_stream(k):
        stepper.set_external_forces_host(fext_host)
        stepper.step(1)
        core.get_state_host_async(x_host[k & 1], v_host[k & 1])
        core.wait_state_host(1)            # the state of step k - 1 has landed in x_host / v_host [(k - 1) & 1]
    ms_e2e = timed(e2e_stream, lambda: core.wait_state_host(0))
    # (b) lock-step: upload, step, download, host sync -- nothing overlaps
    def e2e_sync(k):
        stepper.set_external_forces_host(fext_host)
        stepper.step(1)
        core.get_state_host(x_host[0], v_host[0])
        core.synchronize()
    ms_e2e_sync = timed(e2e_sync, core.synchronize)
    # the step's work does not depend on the state, so a scene that diverges physically (possible at the reference defaults for very
    # long runs) still times the same kernels; say so in the line instead of losing it
    state_finite = bool(torch.isfinite(x_host[0]).all()) and bool(torch.isfinite(x_host[1]).all())
    if not state_finite:
        print("warning: the scene's state is no longer finite at the end of the run", file=sys.stderr, flush=True)
    if mode in ("slab", "strong"):
        assert stepper.halo_ok(), "a halo flag wait timed out"

    # ---- roofline of the dominant kernel: CUDA events around every launch (library stream), L2 warm
    peaks, peak_src = measured_peaks()
    hbm_peak = float(peaks["hbm_gbs"])
    Kp = min(K, 50)
    if net:
        core.set_sdf_obstacle(None, None)          # per-kernel timing of the two gather kernels alone
    ms_def, ms_for = core.profile_step(Kp)
    ms_def, ms_for = ms_def / Kp, ms_for / Kp
    nloc = core.n
    kern = {
        "k_force_c": {"ms": ms_for, "bytes_per_particle": BYTES_FORCE, "gbs": nloc * BYTES_FORCE / (ms_for * 1e-3) / 1e9,
                      "pairs_per_s": info.total_pairs / (ms_for * 1e-3),
                      "fp32_tflops_algorithmic": FLOP_PER_PAIR_FORCE * info.total_pairs / (ms_for * 1e-3) / 1e12},
        "k_deform_c": {"ms": ms_def, "bytes_per_particle": BYTES_DEFORM, "gbs": nloc * BYTES_DEFORM / (ms_def * 1e-3) / 1e9,
                       "pairs_per_s": 2 * info.total_pairs / (ms_def * 1e-3),
                       "fp32_tflops_algorithmic": FLOP_PER_PAIR_DEFORM * info.total_pairs / (ms_def * 1e-3) / 1e12},
    }
    sdf_obj = None
    if net:
        m_rows = 16384
        ms_gemm = net.profile_gemm(m_rows, reps=20)
        flop = 2.0 * m_rows * 1024 * 1024
        tf32_peak = float(peaks.get("bf16_tflops", 1500.0)) / 2.0
        sdf_obj = {"kernel": "k_sdf_gemm (tcgen05.mma kind::tf32, 3 products per fp32-accurate product)", "rows": m_rows,
                   "ms_per_layer": ms_gemm, "tflops_fp32_equivalent": flop / ms_gemm / 1e9, "tflops_tf32_issued": 3 * flop / ms_gemm / 1e9,
                   "tf32_peak_tflops": tf32_peak, "tensor_frac": 3 * flop / ms_gemm / 1e9 / tf32_peak,
                   "peak_note": "TF32 dense peak taken as half the measured bf16 burst peak (" + peak_src + ")",
                   "broad_phase_candidates_last_step": counts[0], "in_contact_band_last_step": counts[1]}
    dom = "k_force_c" if ms_for >= ms_def else "k_deform_c"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom, {}).get("bytes_per_launch_n100k")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kern[dom]["gbs"] / hbm_peak, "frac_of_spec_8000": kern[dom]["gbs"] / 8000.0, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": nloc * kern[dom]["bytes_per_particle"], "launch_ms": kern[dom]["ms"],
                "note": "at ~240 neighbours/particle the gather kernels are FP32-pipe / L1 bound, not HBM bound (SURVEY 8d): "
                        "fp32_fraction (algorithmic flops / measured FFMA peak) is the binding figure",
                "fp32_peak_tflops": FP32_PEAK_TFLOPS, "fp32_fraction": kern[dom]["fp32_tflops_algorithmic"] / FP32_PEAK_TFLOPS,
                "kernels": kern}
    # BASELINE.json's metric names the force kernel's HBM fraction explicitly: always present, whichever kernel dominates
    roofline["force_kernel"] = {"kernel": "k_force_c", "achieved": kern["k_force_c"]["gbs"], "unit": "GB/s",
                                "frac": kern["k_force_c"]["gbs"] / hbm_peak, "frac_of_spec_8000": kern["k_force_c"]["gbs"] / 8000.0,
                                "algorithmic_bytes_per_particle": BYTES_FORCE}
    if sdf_obj:
        roofline["sdf_mlp"] = sdf_obj

    value = n_total * K / (ms_flushed * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_flushed / K, "higher_is_better": True, "scaling": "strong" if mode == "strong" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_total, mean_k, world, mode, extra),
        "pairs_per_sec": value * mean_k,
        "steady_state": {"value": n_total * K / (ms_steady * 1e-3), "ms_per_step": ms_steady / K,
                         "note": "K chained steps (CUDA-graph chunks when no exchange intervenes), no L2 flush"},
        "e2e": {"value": n_total * Ke / (ms_e2e * 1e-3), "unit": UNIT, "steps": Ke,
                "h2d_bytes_per_step": n_io * 12 * world, "d2h_bytes_per_step": n_io * 24 * world,
                "what": "per step: mis_set_ext_force_host (pinned H2D of the force field) + step(1) + mis_get_state_host_async (x, v D2H to pinned "
                        "memory); transfers run on the library's copy stream and overlap the next step; the host waits for and owns the state "
                        "of step k-1 before it submits step k+1",
                "lock_step": {"value": n_total * Ke / (ms_e2e_sync * 1e-3), "what": "same traffic, host sync after every step (no overlap)"}},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "state_finite": state_finite,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        n_s, _ = size_cpu_sample(cfg, n_total, 1, budget_s=args.cpu_budget)
        rate, dt, cores, n_used, dt_mlp = cpu_faithful_rate(cfg, n_s, 1, 0)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 step of a {n_used}-particle sphere (same spacing/params; {dt:.1f} s, of which {dt_mlp:.2f} s the 9x1024 MLP on 1% "
                      f"of the particles), oracle FAITHFUL mode (27-cell walk, per-candidate svd3, sim.py:353-358), {cores} OpenMP threads"}
        try:
            line["cpu_baseline"]["other_modes"] = cpu_side_rates(cfg)
        except Exception as e:                       # a reported extra, never a reason to lose the bench line
            line["cpu_baseline"]["other_modes"] = {"error": str(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    stepper.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--particles", dest="n", type=int, default=0, help="(use --particles under torchrun, whose own parser finds --n ambiguous) " "particles per GPU (default: 100 000 at N = 1 = configs[1]; 1 250 000 at N > 1, so that "
                                                     "N = 8 is the 10M-particle scene of configs[4])")
    ap.add_argument("--n-total", type=int, default=10_000_000, help="total particles for --mode strong")
    ap.add_argument("--mode", default="slab", choices=["slab", "batch", "strong"], help="multi-GPU workload (N > 1)")
    ap.add_argument("--halo", default="auto", choices=["auto", "p2p", "nccl"], help="slab modes: ghost exchange mechanism")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per cluster (0 = library default)")
    ap.add_argument("--cluster", type=int, default=0, help="particles per cluster (0 = library default)")
    ap.add_argument("--no-obstacle", action="store_true", help="N = 1: ground-plane contact only")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=60.0, help="--impl reference: seconds of CPU work the sample is sized for")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.n <= 0:
        args.n = 100_000 if int(os.environ.get("WORLD_SIZE", "1")) == 1 else 1_250_000

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
