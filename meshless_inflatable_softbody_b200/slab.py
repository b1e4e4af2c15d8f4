"""Static slab partition of one large scene across the GPUs of a box, with halo particle exchange.

The reference is single-GPU (no NCCL/MPI call site exists in it).  Because it is Total-Lagrangian --
every neighbour query is centred on the reference position x0 (sim.py:161,178,203,224) -- the partition, the ghost
sets and the send/receive index lists are STATIC for the whole run:

  * slabs along the longest axis of the x0 bounding box, cut at quantiles of that coordinate: owned particle counts
    are equal to the particle;
  * a rank keeps two ghost layers of depth 2h (the support radius, sim.py:137-141) on each side of its slab: layer 1
    holds every neighbour of an owned particle; layer 2 holds every neighbour of a layer-1 particle, so the rank can recompute
    R_j, S_j of its layer-1 ghosts locally (compute_A_pq / compute_nabla_u, sim.py:170-209) and needs ONE exchange
    per step: the new positions (12 B / ghost) written by part_1 (sim.py:247-251);
  * ghosts are pinned locally (free_points = 0, sim.py:285-286) and overwritten by the exchange.

There is no global reduction in the physics (the reference has no pressure / enclosed-volume term).

Two exchange mechanisms:
  * halo="p2p" (default when every rank is on one NVLink box): the fused halo push of include/mis.h -- the force
    kernel's part_1 epilogue stores a boundary particle's new position straight into the peers' ghost slots over
    NVLink peer memory, a one-block kernel publishes / awaits an epoch flag, and step(n) is CUDA-graph chunks with no
    host, NCCL or copy kernel in the loop;
  * halo="nccl": gather -> batch_isend_irecv -> scatter after every step (works across nodes; the baseline).
Volumes V_i = m_i / rho_i (compute_v_i, sim.py:154-167) are static and need a particle's whole neighbourhood, which an
outer ghost lacks locally: owners send them once after set_mass.

Host logic only (numpy + torch.distributed point-to-point); the particle arithmetic stays in the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np


@dataclass
class RankPlan:
    rank: int
    owned: np.ndarray                     # global ids owned by this rank (ascending)
    ghosts: np.ndarray                    # global ids of ghost particles (ascending); local order = owned then ghosts
    ghost_layer: np.ndarray               # 1 or 2 per ghost
    send: Dict[int, np.ndarray] = field(default_factory=dict)   # peer -> LOCAL ids (owned) to send, ordered by global id
    recv: Dict[int, np.ndarray] = field(default_factory=dict)   # peer -> LOCAL ids (ghosts) to fill, ordered by global id

    @property
    def local_ids(self) -> np.ndarray:
        return np.concatenate([self.owned, self.ghosts])

    @property
    def n_owned(self) -> int:
        return int(len(self.owned))


@dataclass
class SlabPartition:
    axis: int
    cuts: np.ndarray                      # world_size + 1 coordinates along `axis`: rank r owns particles with cuts[r] <= x0[axis] < cuts[r+1]
    plans: List[RankPlan]

    @classmethod
    def build(cls, x0: np.ndarray, h: float, world_size: int, axis: Optional[int] = None, ghost_layers: int = 2,
              extra_cost: Optional[np.ndarray] = None, weights: Optional[np.ndarray] = None) -> "SlabPartition":
        """Cut planes at quantiles of the x0 coordinate along `axis` (equal owned counts, to the particle); ghost layer L of a
        rank = foreign particles within L * 2h of its slab along the axis.  Every neighbour (|x0_i - x0_j| < 2h, sim.py:137-141)
        of an owned particle is owned or layer 1; every neighbour of a layer-1 ghost is local.

        extra_cost[r]: fixed per-step work of rank r that is not proportional to its particle count (e.g. the obstacle's MLP
        query, which only the ranks under the obstacle run), in particle-equivalents: the cuts then equalise
        owned_r + extra_cost[r] instead of owned_r.  The partition is static, so this is a one-off load balance.

        weights[i]: relative per-step work of particle i (default 1).  The gather kernels' work is proportional to a particle's
        neighbour count, which is lower near the surface: `neighbour_weights(x0, h)` estimates it from the 27-cell occupancy, so
        that the tapered end slabs of a body own more particles than the full-section slabs in the middle.  extra_cost is then in
        units of the mean weight."""
        x0 = np.asarray(x0, np.float32).reshape(-1, 3)
        n = len(x0)
        if axis is None:
            axis = int(np.argmax(x0.max(0) - x0.min(0)))
        c = x0[:, axis].astype(np.float64)
        order = np.argsort(c, kind="stable")
        cs = c[order]
        reach = 2.0 * float(np.float32(h)) * (1.0 + 1e-4)          # support radius plus a margin far above fp32 rounding of the distance test
        extra = np.zeros(world_size) if extra_cost is None else np.asarray(extra_cost, np.float64).reshape(world_size)
        share = (n + extra.sum()) / world_size                       # cost every rank should carry
        owned_target = np.maximum(share - extra, 0.0)
        owned_target *= n / owned_target.sum()
        bounds = np.concatenate([[0.0], np.cumsum(owned_target)])
        if weights is not None:                                      # cumulative work along the axis instead of particle counts
            w = np.asarray(weights, np.float64).reshape(n)[order]
            cw = np.cumsum(w) * (n / w.sum())                        # same scale as a particle count
        cuts = [-np.inf]
        for r in range(1, world_size):
            k = int(round(bounds[r])) if weights is None else int(np.searchsorted(cw, bounds[r]))
            k = min(max(k, 1), n - 1)
            cuts.append(0.5 * (cs[k - 1] + cs[k]) if cs[k] > cs[k - 1] else cs[k])
        cuts.append(np.inf)
        cuts = np.asarray(cuts, np.float64)
        inner = np.diff(cuts[1:-1]) if world_size > 2 else np.zeros(0)
        # an inner slab must contain its neighbours' whole ghost depth (ghosts then come from adjacent ranks only, and a particle
        # is mirrored on at most two peers); the two end slabs have nobody beyond them
        if world_size > 1 and (np.any(inner < ghost_layers * reach) or np.any(np.diff(cuts) <= 0)):
            raise ValueError("slabs thinner than the ghost depth: too many ranks for this scene along its longest axis")

        def in_range(lo, hi):              # global ids (ascending) with lo <= coordinate < hi
            a, b = np.searchsorted(cs, lo, "left"), np.searchsorted(cs, hi, "left")
            return np.sort(order[a:b])
        plans = []
        for r in range(world_size):
            lo, hi = cuts[r], cuts[r + 1]
            owned = in_range(lo, hi)
            g_ids, g_layer = [], []
            for layer in range(1, ghost_layers + 1):
                for a, b in ((lo - layer * reach, lo - (layer - 1) * reach), (hi + (layer - 1) * reach, hi + layer * reach)):
                    if np.isfinite(a) or np.isfinite(b):
                        ids = in_range(a, b)
                        g_ids.append(ids); g_layer.append(np.full(len(ids), layer, np.int32))
            ghosts = np.concatenate(g_ids) if g_ids else np.zeros(0, np.int64)
            layers = np.concatenate(g_layer) if g_layer else np.zeros(0, np.int32)
            o = np.argsort(ghosts, kind="stable")
            plans.append(RankPlan(rank=r, owned=owned, ghosts=ghosts[o], ghost_layer=layers[o]))
        # send / receive lists: a ghost of rank r owned by rank q is sent q -> r; both sides order by global id
        owner = np.empty(n, np.int64)
        for p in plans:
            owner[p.owned] = p.rank
        for p in plans:
            ghost_owner = owner[p.ghosts]
            for q in np.unique(ghost_owner):
                q = int(q)
                sel = np.nonzero(ghost_owner == q)[0]
                p.recv[q] = (p.n_owned + sel).astype(np.int64)                    # local ids of those ghosts
                gl = p.ghosts[sel]                                                # their global ids (ascending)
                plans[q].send[p.rank] = np.searchsorted(plans[q].owned, gl).astype(np.int64)   # local (= owned index) on the sender
        return cls(axis=axis, cuts=cuts, plans=plans)


def neighbour_weights(x0: np.ndarray, h: float) -> np.ndarray:
    """Per-particle work estimate for SlabPartition.build(weights=...): the number of particles in the 27 cells (width 2h) around
    a particle's cell -- proportional, on average, to its neighbour count (support radius 2h), hence to its share of the gather
    kernels' pair evaluations.  Host numpy, one pass over the particles."""
    x0 = np.asarray(x0, np.float32).reshape(-1, 3)
    cw = 2.0 * float(np.float32(h))
    ijk = np.floor(x0.astype(np.float64) / cw).astype(np.int64)
    ijk -= ijk.min(0)
    dims = ijk.max(0) + 3                                            # one empty cell of padding on every side
    lin = ((ijk[:, 2] + 1) * dims[1] + (ijk[:, 1] + 1)) * dims[0] + (ijk[:, 0] + 1)
    occ = np.bincount(lin, minlength=int(dims.prod())).astype(np.float64).reshape(dims[2], dims[1], dims[0])
    nb = np.zeros_like(occ)
    for dz in (-1, 0, 1):                                            # 27-cell box sum (separable would do; the grid is small)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                nb += np.roll(occ, (dz, dy, dx), axis=(0, 1, 2))
    return nb.reshape(-1)[lin]


def exchange_halo(plan: RankPlan, gather, scatter, dist=None, group=None, device=None):
    """One halo exchange: send the new positions of owned boundary particles to every peer, receive the ghosts'.

    gather(local_ids_tensor) -> (k,3) fp32 tensor on `device`; scatter(local_ids_tensor, (k,3) tensor).
    Point-to-point only (each rank talks to at most two slab neighbours); no collective."""
    import torch
    if not plan.send and not plan.recv:
        return
    ops, recv_bufs = [], {}
    send_keep = []
    for q, ids in sorted(plan.recv.items()):
        buf = torch.empty((len(ids), 3), dtype=torch.float32, device=device)
        recv_bufs[q] = buf
        ops.append(dist.P2POp(dist.irecv, buf, q, group=group))
    for q, ids in sorted(plan.send.items()):
        buf = gather(ids)
        send_keep.append(buf)
        ops.append(dist.P2POp(dist.isend, buf, q, group=group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for q, ids in sorted(plan.recv.items()):
        scatter(ids, recv_bufs[q])


def plan_push(plan: RankPlan, peers: List[int], peer_recv_slots: Dict[int, np.ndarray]):
    """(ids, peer index, remote slot) triples of the fused halo push.  The k-th entry of plan.send[q] and the k-th entry
    of rank q's recv[plan.rank] are the same particle (both ordered by global id); peer_recv_slots[q][k] is the slot rank q
    keeps that ghost in.  `peers` fixes the peer indices (position in the list)."""
    ids, pidx, slots = [], [], []
    for q, local in sorted(plan.send.items()):
        sl = np.asarray(peer_recv_slots[q]).reshape(-1)
        if len(sl) != len(local):
            raise ValueError(f"rank {plan.rank} sends {len(local)} particles to {q}, which expects {len(sl)}")
        ids.append(np.asarray(local)); pidx.append(np.full(len(local), peers.index(q))); slots.append(sl)
    cat = lambda v: np.concatenate(v).astype(np.int32) if v else np.zeros(0, np.int32)
    return cat(ids), cat(pidx), cat(slots)


def peers_of(plan: RankPlan) -> List[int]:
    return sorted(set(plan.send) | set(plan.recv))


def exchange_volumes_in_process(sims):
    """Owners' volumes -> ghosts, all ranks in one process."""
    vols = {s.rank: s.sim.volumes() for s in sims}
    for s in sims:
        ids, vals = [], []
        for q, rid in s._recv.items():
            ids.append(rid); vals.append(vols[q][sims[q]._send[s.rank].long()].to(s.device))
        if ids:
            import torch
            s.sim.set_volumes(torch.cat(ids), torch.cat(vals))


def connect_in_process(sims):
    """Fused halo push between ranks that live in ONE process on one GPU (tests): raw device pointers instead of IPC handles,
    and NO in-kernel flag waits -- kernels of different ranks on one GPU are not guaranteed to run at the same time, so the
    host orders the ranks (step_in_process synchronises between steps).  The push tables, the P2P-store epilogue and the epoch
    counters are the ones the multi-GPU path uses."""
    ptrs = {s.rank: s.sim.halo_local_ptrs() for s in sims}
    recv_slots = {s.rank: {q: s.sim.slots_of(ids).cpu().numpy() for q, ids in s._recv.items()} for s in sims}
    peer_lists = {s.rank: peers_of(s.plan) for s in sims}
    for s in sims:
        peers = peer_lists[s.rank]
        ids, pidx, slots = plan_push(s.plan, peers, {q: recv_slots[q][s.rank] for q in s.plan.send})
        flag = [ptrs[q][2] + 4 * peer_lists[q].index(s.rank) for q in peers]
        s.sim.halo_connect([ptrs[q][0] for q in peers], [ptrs[q][1] for q in peers], flag, ids, pidx, slots,
                           np.arange(s.n_owned, s.sim.n, dtype=np.int32), s.plan.ghost_layer)
        s.sim.halo_set_wait(False)
        s.halo = "p2p"
    for s in sims:
        s.sim.synchronize()


def step_in_process(sims, n_steps: int = 1):
    """All ranks of a partition driven from ONE process (tests, single-GPU debugging): the same per-step sequence as
    SlabSimulator.step, with the point-to-point exchange replaced by device-to-device copies."""
    if sims and sims[0].halo == "p2p":
        # the pushes are inside each rank's step; one step at a time, host-ordered (see connect_in_process)
        for s in sims:
            s.sim.synchronize()                  # the peers' earlier pushes (priming) have landed
        for _ in range(int(n_steps)):
            for s in sims:
                s.sim.step(1)
            for s in sims:
                s.sim.synchronize()
        for s in sims:
            s.frame += int(n_steps)
        return

    def exchange():
        bufs = {}
        for s in sims:
            for q, ids in s._send.items():
                bufs[(s.rank, q)] = s.sim.gather_next_positions(ids)
        for s in sims:
            for q, ids in s._recv.items():
                s.sim.scatter_next_positions(ids, bufs[(q, s.rank)].to(s.device))
            s.exchanges += 1
    if n_steps == 0:
        exchange()
        return
    for _ in range(int(n_steps)):
        for s in sims:
            s.sim.step(1)
        exchange()
    for s in sims:
        s.frame += int(n_steps)


class SlabSimulator:
    """One rank's share of a slab-partitioned scene: a local Simulator over owned + ghost particles plus the
    per-step halo exchange.  Method names follow Simulator / the reference's control functions."""

    def __init__(self, x0_global, config=None, rank: int = 0, world_size: int = 1, device: str = "cuda:0",
                 group=None, partition: Optional[SlabPartition] = None, in_process: bool = False, halo: str = "auto",
                 extra_cost=None, weights=None, **sim_kw):
        import torch
        import torch.distributed as dist
        from .config import SceneConfig
        from .simulator import Simulator
        self.cfg = config or SceneConfig()
        self.rank, self.world = rank, world_size
        self.in_process = in_process          # True: the caller drives the exchange (step_in_process)
        self.dist, self.group = dist, group
        x0_global = np.asarray(x0_global, np.float32).reshape(-1, 3)
        self.n_global = len(x0_global)
        self.partition = partition or SlabPartition.build(x0_global, self.cfg.h, world_size, extra_cost=extra_cost, weights=weights)
        self.plan = self.partition.plans[rank]
        self.device = torch.device(device)
        local = self.plan.local_ids
        self.sim = Simulator(x0_global[local], self.cfg, device=device, **sim_kw)
        self.n_owned = self.plan.n_owned
        self._x0_owned = x0_global[self.plan.owned]
        if len(self.plan.ghosts):
            ghost_local = torch.arange(self.n_owned, len(local), device=self.device)
            self.sim.set_dirichlet(ghost_local, [0.0, 0.0, 0.0])        # ghosts move only through the exchange
        self._send = {q: torch.as_tensor(ids, dtype=torch.int32, device=self.device) for q, ids in self.plan.send.items()}
        self._recv = {q: torch.as_tensor(ids, dtype=torch.int32, device=self.device) for q, ids in self.plan.recv.items()}
        self.frame = 0
        self.exchanges = 0
        self.halo = "nccl"
        self._mapped = []
        if world_size > 1 and not in_process:
            self._exchange_volumes()
            requested = halo
            if halo == "auto":
                import os
                local = int(os.environ.get("LOCAL_WORLD_SIZE", world_size))
                halo = "p2p" if local == world_size else "nccl"
            if halo not in ("p2p", "nccl"):
                raise ValueError("halo must be 'auto', 'p2p' or 'nccl'")
            if halo == "p2p":
                err = self._connect_p2p()
                if err is not None:                  # agreed by all ranks: nobody is connected
                    if requested == "p2p":
                        raise RuntimeError("fused halo push unavailable: " + err)
                    self.halo = "nccl"               # 'auto' falls back to the NCCL exchange

    # halo plumbing -----------------------------------------------------------------------------------------
    def _exchange_volumes(self):
        """Owners' V_i -> ghosts (once per set_mass): point-to-point, same lists as the position exchange."""
        import torch
        vol = self.sim.volumes()
        ops, bufs = [], {}
        for q, ids in sorted(self._recv.items()):
            bufs[q] = torch.empty(len(ids), dtype=torch.float32, device=self.device)
            ops.append(self.dist.P2POp(self.dist.irecv, bufs[q], q, group=self.group))
        keep = []
        for q, ids in sorted(self._send.items()):
            keep.append(vol[ids.long()].contiguous())
            ops.append(self.dist.P2POp(self.dist.isend, keep[-1], q, group=self.group))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
            torch.cuda.synchronize(self.device)
        if bufs:
            qs = sorted(bufs)
            self.sim.set_volumes(torch.cat([self._recv[q] for q in qs]), torch.cat([bufs[q] for q in qs]))

    def _connect_p2p(self):
        """Exchange CUDA IPC handles of the position buffers / flag arrays and the ghosts' slots, map the peers' memory,
        and hand the push table to the library (mis_halo_connect).  Collective: every rank takes part in both exchanges even
        if a local step failed, and all ranks agree on the outcome.  Returns None on success, else the first error message
        (no rank is connected then)."""
        import ctypes as C
        import torch
        from . import native
        sim = self.sim
        peers = peers_of(self.plan)
        mine = {"peers": peers, "error": None}
        try:
            mine["handles"] = sim.halo_ipc_handles()
            mine["recv_slots"] = {q: sim.slots_of(ids).cpu().numpy() for q, ids in self._recv.items()}
        except Exception as e:                       # e.g. CUDA IPC not permitted in this container
            mine["error"] = f"rank {self.rank}: {e}"
        everyone = [None] * self.world
        self.dist.all_gather_object(everyone, mine, group=self.group)
        error = next((m["error"] for m in everyone if m["error"]), None)
        if error is None:
            try:
                xv0, xv1, flag = [], [], []
                for q in peers:
                    h = everyone[q]["handles"]
                    ptr = []
                    for k in range(3):
                        p = C.c_void_p()
                        native.check(sim.L.mis_ipc_open(h[64 * k: 64 * k + 64], C.byref(p)), "mis_ipc_open")
                        ptr.append(int(p.value)); self._mapped.append(int(p.value))
                    xv0.append(ptr[0]); xv1.append(ptr[1])
                    flag.append(ptr[2] + 4 * everyone[q]["peers"].index(self.rank))
                ids, pidx, slots = plan_push(self.plan, peers, {q: everyone[q]["recv_slots"][self.rank] for q in self.plan.send})
                sim.halo_connect(xv0, xv1, flag, ids, pidx, slots, np.arange(self.n_owned, sim.n, dtype=np.int32), self.plan.ghost_layer)
            except Exception as e:                   # e.g. no peer access between two of the GPUs
                error = f"rank {self.rank}: {e}"
        ok = torch.tensor([0 if error else 1], dtype=torch.int32, device=self.device)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            self.halo = "p2p"
            return None
        errors = [None] * self.world
        self.dist.all_gather_object(errors, error, group=self.group)
        sim.halo_disconnect()
        for p in self._mapped:
            sim.L.mis_ipc_close(p)
        self._mapped = []
        return next((e for e in errors if e), "unknown error")

    def _exchange(self):
        if self.world == 1 or self.in_process or self.halo == "p2p":
            return
        import torch
        sim = self.sim
        plan = RankPlan(self.rank, self.plan.owned, self.plan.ghosts, self.plan.ghost_layer, self._send, self._recv)
        exchange_halo(plan, sim.gather_next_positions, sim.scatter_next_positions, dist=self.dist, group=self.group,
                      device=self.device)
        self.exchanges += 1

    # control functions ---------------------------------------------------------------------------------------
    def set_mass(self, m):
        self.sim.set_mass(m)
        if self.world > 1 and not self.in_process:
            self._exchange_volumes()

    def set_sdf_obstacle(self, sdf, bbox_model, xform=None, fd_eps: float = 1e-3):
        """Per-step obstacle contact for this rank's particles (Simulator.set_sdf_obstacle).  Contact is local to a particle
        (no neighbour sum), so no exchange is added; the force of a ghost is never used (its owner integrates it)."""
        self.sim.set_sdf_obstacle(sdf, bbox_model, xform=xform, fd_eps=fd_eps)

    def obstacle_nearby(self, bbox_world, margin: float) -> bool:
        """True if any OWNED particle's reference position lies within `margin` of the world-space box (min xyz, max xyz).
        A rank for which this is False, with `margin` above the distance the body travels during the run, can skip the
        per-step obstacle query altogether (slab-level broad phase; the partition is static)."""
        bb = np.asarray(bbox_world, np.float64).reshape(2, 3)
        x0 = self._x0_owned
        return bool(np.any(np.all((x0 >= bb[0] - margin) & (x0 <= bb[1] + margin), axis=1)))

    def startup(self, v0=None):
        if self.halo == "p2p" and self.world > 1 and not self.in_process:
            self.sim.synchronize()
            self.dist.barrier(group=self.group)       # every rank's set-up kernels are done before the first push
        self.sim.startup(v0)
        self.sim.step(0)          # frame-0 force evaluation + part_1 (sim.py:349-353) so that x(1) exists
        self._exchange()
        self.frame = 0

    def step(self, n_steps: int = 1):
        if self.halo == "p2p":
            self.sim.step(int(n_steps))               # pushes + flag waits are inside the step graph
            self.exchanges += int(n_steps)
            self.frame += int(n_steps)
            self._steps_since_check = getattr(self, "_steps_since_check", 0) + int(n_steps)
            if self._steps_since_check >= 4096:       # periodic (synchronising) look at the device-side timeout flag
                self._steps_since_check = 0
                self._check_halo()
            return
        for _ in range(int(n_steps)):
            self.sim.step(1)
            self._exchange()
        self.frame += int(n_steps)

    def set_external_forces_host(self, f_host):
        """Per-step input path for the local particles (owned + ghosts): part_1 is redone with the new force and the
        ghosts' new positions are exchanged again."""
        self.sim.set_external_forces_host(f_host)
        self.sim.step(0)
        self._exchange()

    def _check_halo(self):
        """A flag wait that timed out (a peer died, or the ranks' call sequences diverged -- e.g. eval_forces / a re-prime on one
        rank only) is sticky on the device and every later step runs on stale ghosts: never hand such a state out silently."""
        if self.halo == "p2p" and self.world > 1 and not self.in_process:
            timed_out, _ = self.sim.halo_status()
            if timed_out:
                raise RuntimeError(f"rank {self.rank}: a halo flag wait timed out (MIS_HALO_TIMEOUT_MS); the state after that step used stale ghost positions")

    def position_velocity(self):
        """(x, v) of the OWNED particles, in the order of plan.owned (ascending global id)."""
        self._check_halo()
        x, v = self.sim.position_velocity()
        return x[: self.n_owned], v[: self.n_owned]

    def gather_global(self):
        """position / velocity of the whole scene in global particle order, on every rank (export path, sim.py:368-369)."""
        import torch
        x, v = self.position_velocity()
        if self.world == 1:
            return x, v
        counts = [p.n_owned for p in self.partition.plans]
        cap = max(counts)                                   # equal-size buffers: NCCL all_gather needs them
        mine = torch.zeros((cap, 6), dtype=torch.float32, device=self.device)
        mine[: self.n_owned, :3] = x; mine[: self.n_owned, 3:] = v
        parts = [torch.empty_like(mine) for _ in counts]
        self.dist.all_gather(parts, mine, group=self.group)
        X = torch.empty((self.n_global, 3), dtype=torch.float32, device=self.device)
        V = torch.empty_like(X)
        for p, c, buf in zip(self.partition.plans, counts, parts):
            idx = torch.as_tensor(p.owned, device=self.device)
            X[idx] = buf[:c, :3]; V[idx] = buf[:c, 3:]
        return X, V

    def halo_ok(self) -> bool:
        """False if a flag wait of the fused push timed out (a peer died or the ranks' call sequences diverged)."""
        return not self.sim.halo_status()[0]

    def close(self):
        if self._mapped:
            try:
                self._check_halo()
            except RuntimeError as e:          # closing must still unmap; say what happened
                import warnings
                warnings.warn(str(e))
            self.sim.synchronize()
            self.dist.barrier(group=self.group)       # nobody is still pushing into memory about to be unmapped
            self.sim.halo_disconnect()
            for p in self._mapped:
                self.sim.L.mis_ipc_close(p)
            self._mapped = []
        self.sim.close()
