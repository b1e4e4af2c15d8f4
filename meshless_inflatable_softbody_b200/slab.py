"""Static slab partition of one large scene across the GPUs of a box, with halo particle exchange.

The reference is single-GPU (no NCCL/MPI call site exists in it).  Because it is Total-Lagrangian --
every neighbour query is centred on the reference position x0 (sim.py:161,178,203,224) -- the partition, the ghost
sets and the send/receive index lists are STATIC for the whole run:

  * slabs along the longest axis of the x0 bounding box, cut at hash-grid cell boundaries (cell width 2h,
    sim.py:127) so that owned particle counts are balanced;
  * a rank keeps two ghost cell layers on each side of its slab: layer 1 (cells adjacent to the slab) holds every
    neighbour of an owned particle; layer 2 holds every neighbour of a layer-1 particle, so the rank can recompute
    R_j, S_j of its layer-1 ghosts locally (compute_A_pq / compute_nabla_u, sim.py:170-209) and needs ONE exchange
    per step: the new positions (12 B / ghost) written by part_1 (sim.py:247-251);
  * ghosts are pinned locally (free_points = 0, sim.py:285-286) and overwritten by the exchange.

There is no global reduction in the physics (the reference has no pressure / enclosed-volume term).

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
    cuts: np.ndarray                      # world_size + 1 cell coordinates along `axis`: rank r owns cells [cuts[r], cuts[r+1])
    plans: List[RankPlan]

    @staticmethod
    def cell_coord(x0: np.ndarray, h: float, axis: int) -> np.ndarray:
        """int(p * (1 / (2h))) with fp32 arithmetic and truncation, as the hash grid bins x0 (sim.py:127)."""
        inv = np.float32(1.0) / (np.float32(2.0) * np.float32(h))
        return np.trunc(np.asarray(x0, np.float32)[:, axis] * inv).astype(np.int64)

    @classmethod
    def build(cls, x0: np.ndarray, h: float, world_size: int, axis: Optional[int] = None, ghost_cells: int = 2) -> "SlabPartition":
        x0 = np.asarray(x0, np.float32).reshape(-1, 3)
        n = len(x0)
        if axis is None:
            axis = int(np.argmax(x0.max(0) - x0.min(0)))
        c = cls.cell_coord(x0, h, axis)
        cmin, cmax = int(c.min()), int(c.max())
        counts = np.bincount(c - cmin, minlength=cmax - cmin + 1)
        cum = np.concatenate([[0], np.cumsum(counts)])
        # cut r at the cell boundary whose cumulative count is closest to r * n / world_size (monotone, non-empty slabs)
        cuts = [cmin]
        for r in range(1, world_size):
            target = r * n / world_size
            k = int(np.argmin(np.abs(cum - target)))
            k = max(k, cuts[-1] - cmin + 1)
            k = min(k, len(counts) - (world_size - r))
            cuts.append(cmin + k)
        cuts.append(cmax + 1)
        cuts = np.asarray(cuts, np.int64)
        if np.any(np.diff(cuts) < ghost_cells) and world_size > 1:
            raise ValueError("slabs thinner than the ghost depth: too many ranks for this scene along its longest axis")
        order = np.argsort(c, kind="stable")
        c_sorted = c[order]
        def in_cells(lo, hi):              # global ids (ascending) with lo <= cell < hi
            a, b = np.searchsorted(c_sorted, lo, "left"), np.searchsorted(c_sorted, hi, "left")
            return np.sort(order[a:b])
        plans = []
        for r in range(world_size):
            lo, hi = int(cuts[r]), int(cuts[r + 1])
            owned = in_cells(lo, hi)
            g_ids, g_layer = [], []
            for layer in range(1, ghost_cells + 1):
                for a, b in ((lo - layer, lo - layer + 1), (hi + layer - 1, hi + layer)):
                    ids = in_cells(a, b)
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


def step_in_process(sims, n_steps: int = 1):
    """All ranks of a partition driven from ONE process (tests, single-GPU debugging): the same per-step sequence as
    SlabSimulator.step, with the point-to-point exchange replaced by device-to-device copies."""
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
                 group=None, partition: Optional[SlabPartition] = None, in_process: bool = False, **sim_kw):
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
        self.partition = partition or SlabPartition.build(x0_global, self.cfg.h, world_size)
        self.plan = self.partition.plans[rank]
        self.device = torch.device(device)
        local = self.plan.local_ids
        self.sim = Simulator(x0_global[local], self.cfg, device=device, **sim_kw)
        self.n_owned = self.plan.n_owned
        if len(self.plan.ghosts):
            ghost_local = torch.arange(self.n_owned, len(local), device=self.device)
            self.sim.set_dirichlet(ghost_local, [0.0, 0.0, 0.0])        # ghosts move only through the exchange
        self._send = {q: torch.as_tensor(ids, dtype=torch.int32, device=self.device) for q, ids in self.plan.send.items()}
        self._recv = {q: torch.as_tensor(ids, dtype=torch.int32, device=self.device) for q, ids in self.plan.recv.items()}
        self.frame = 0
        self.exchanges = 0

    # halo plumbing -----------------------------------------------------------------------------------------
    def _exchange(self):
        if self.world == 1 or self.in_process:
            return
        import torch
        sim = self.sim
        plan = RankPlan(self.rank, self.plan.owned, self.plan.ghosts, self.plan.ghost_layer, self._send, self._recv)
        exchange_halo(plan, sim.gather_next_positions, sim.scatter_next_positions, dist=self.dist, group=self.group,
                      device=self.device)
        self.exchanges += 1

    # control functions ---------------------------------------------------------------------------------------
    def startup(self, v0=None):
        self.sim.startup(v0)
        self.sim.step(0)          # frame-0 force evaluation + part_1 (sim.py:349-353) so that x(1) exists
        self._exchange()
        self.frame = 0

    def step(self, n_steps: int = 1):
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

    def position_velocity(self):
        """(x, v) of the OWNED particles, in the order of plan.owned (ascending global id)."""
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

    def close(self):
        self.sim.close()
