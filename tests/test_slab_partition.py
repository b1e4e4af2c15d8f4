"""Host logic of the slab partition + halo exchange (slab.py), on CPU: invariants of the static plan and a
world_size-2/3 gloo run of a stand-in engine with the SAME dependency stencil as the step (new position of an owned
particle depends on neighbours of neighbours), compared with the serial result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from meshless_inflatable_softbody_b200 import scenes
from meshless_inflatable_softbody_b200.slab import SlabPartition, RankPlan, exchange_halo

H = 0.007


def _beam(n=6000, seed=0):
    return scenes.jittered_beam(n, h=H, seed=seed, aspect=(6.0, 1.0, 1.0))


def _neighbours(x0):
    t = cKDTree(x0.astype(np.float64))
    return t.query_ball_point(x0.astype(np.float64), 2.0 * H * (1 - 1e-7))


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_plan_invariants(world):
    x0 = _beam()
    part = SlabPartition.build(x0, H, world)
    n = len(x0)
    owned_all = np.concatenate([p.owned for p in part.plans])
    assert np.array_equal(np.sort(owned_all), np.arange(n))                 # a partition of the particles
    counts = np.array([p.n_owned for p in part.plans])
    assert counts.min() > 0.6 * n / world and counts.max() < 1.5 * n / world
    nb = _neighbours(x0)
    for p in part.plans:
        local = set(p.local_ids.tolist())
        layer1 = set(p.ghosts[p.ghost_layer == 1].tolist())
        owned_set = set(p.owned.tolist())
        for i in p.owned[:: max(1, len(p.owned) // 200)]:                   # neighbours of owned: owned or layer 1
            assert all((j in owned_set) or (j in layer1) for j in nb[i])
        for i in list(layer1)[:: max(1, len(layer1) // 200)]:               # neighbours of layer 1: local
            assert all(j in local for j in nb[i])
        for q, ids in p.recv.items():                                       # both sides list the same particles, same order
            assert np.array_equal(p.local_ids[ids], part.plans[q].owned[part.plans[q].send[p.rank]])
            assert abs(q - p.rank) == 1                                     # only slab neighbours talk
    if world == 1:
        assert len(part.plans[0].ghosts) == 0 and not part.plans[0].send


def test_too_many_ranks_is_an_error():
    x0, _ = scenes.jittered_sphere(500, seed=0)
    with pytest.raises(ValueError):
        SlabPartition.build(x0, H, 16)


# ---------------------------------------------------------------- stand-in engine with the step's stencil
def _two_hop(x, nbr_idx, nbr_ptr):
    """y_i = mean_j x_j ; x'_i = x_i + 0.25 (mean_j y_j - y_i): depends on neighbours of neighbours."""
    def mean_nb(a):
        s = np.add.reduceat(a[nbr_idx], nbr_ptr[:-1], axis=0)
        return s / np.diff(nbr_ptr)[:, None]
    y = mean_nb(x)
    return x + 0.25 * (mean_nb(y) - y)


def _csr(nb_lists):
    ptr = np.concatenate([[0], np.cumsum([len(l) for l in nb_lists])])
    return np.concatenate([np.asarray(l, np.int64) for l in nb_lists]), ptr


def _worker(rank, world, x0, steps, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = SlabPartition.build(x0, H, world)
    plan = part.plans[rank]
    local = plan.local_ids
    xl = x0[local].astype(np.float64)
    t = cKDTree(xl)
    idx, ptr = _csr(t.query_ball_point(xl, 2.0 * H * (1 - 1e-7)))          # local neighbour lists (ghost lists are incomplete)
    state = {"x": (xl * np.array([1.0, 1.3, 0.7])).copy(), "next": None}
    send = {q: torch.as_tensor(v) for q, v in plan.send.items()}
    recv = {q: torch.as_tensor(v) for q, v in plan.recv.items()}
    tplan = RankPlan(rank, plan.owned, plan.ghosts, plan.ghost_layer, send, recv)
    for _ in range(steps):
        nxt = _two_hop(state["x"], idx, ptr)
        nxt[plan.n_owned:] = state["x"][plan.n_owned:]                     # ghosts pinned, then overwritten by the exchange
        state["next"] = nxt
        def gather(ids):
            return torch.as_tensor(state["next"][ids.numpy()], dtype=torch.float32)
        def scatter(ids, buf):
            state["next"][ids.numpy()] = buf.numpy().astype(np.float64)
        exchange_halo(tplan, gather, scatter, dist=dist, device="cpu")
        state["x"] = state["next"]
    np.save(os.path.join(out_dir, f"x_{rank}.npy"), state["x"][: plan.n_owned])
    np.save(os.path.join(out_dir, f"ids_{rank}.npy"), plan.owned)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_matches_serial(world, tmp_path):
    x0 = _beam(4000, seed=1)
    steps = 4
    xs = (x0.astype(np.float64) * np.array([1.0, 1.3, 0.7])).copy()
    idx, ptr = _csr(_neighbours(x0))
    for _ in range(steps):
        xs = _two_hop(xs, idx, ptr).astype(np.float32).astype(np.float64)   # the exchange carries fp32
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, x0, steps, port, str(tmp_path)), nprocs=world, join=True)
    got = np.zeros_like(xs)
    for r in range(world):
        got[np.load(tmp_path / f"ids_{r}.npy")] = np.load(tmp_path / f"x_{r}.npy")
    # owned particles see exactly the serial data flow; only ghost values are rounded to fp32 in transit
    assert np.abs(got - xs).max() < 5e-8


# ---------------------------------------------------------------- fused halo push: the static push table
def _push_worker(rank, world, x0, port, out_dir):
    """Every rank keeps its local particles in a private, permuted slot order (a stand-in for the library's cell sort),
    publishes the slots of its ghosts, builds the push triples with plan_push and 'pushes' by mailing (slot, value) pairs:
    afterwards every ghost slot must hold the owner's value of exactly that particle."""
    from meshless_inflatable_softbody_b200.slab import plan_push, peers_of
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = SlabPartition.build(x0, H, world)
    plan = part.plans[rank]
    local = plan.local_ids
    rng = np.random.default_rng(100 + rank)
    slot_of = rng.permutation(len(local))                    # local id -> slot
    peers = peers_of(plan)
    mine = {"peers": peers, "recv_slots": {q: slot_of[ids] for q, ids in plan.recv.items()}}
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    ids, pidx, slots = plan_push(plan, peers, {q: everyone[q]["recv_slots"][rank] for q in plan.send})
    assert len(ids) == sum(len(v) for v in plan.send.values())
    assert np.all(ids < plan.n_owned)                        # only owned particles are pushed
    assert np.bincount(ids, minlength=1).max() <= 2          # at most two mirrors per particle (the kernel's table holds two)
    value = local.astype(np.float64) * 10.0 + 1.0            # owner's value of a particle = f(global id)
    mail = {q: [] for q in range(world)}
    for i, p, s in zip(ids, pidx, slots):
        mail[peers[p]].append((int(s), float(value[i])))
    boxes = [None] * world
    dist.all_gather_object(boxes, mail)
    store = np.full(len(local), -1.0)
    writes = np.zeros(len(local), np.int64)
    for src in range(world):
        for s, v in boxes[src][rank]:
            store[s] = v; writes[s] += 1
    ghost_slots = slot_of[plan.n_owned:]
    assert np.all(writes[ghost_slots] == 1)                  # every ghost slot has exactly one writer
    assert writes.sum() == len(plan.ghosts)                  # and nothing else is written
    assert np.array_equal(store[ghost_slots], plan.ghosts.astype(np.float64) * 10.0 + 1.0)
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([1]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_push_plan_fills_every_ghost_slot_once(world, tmp_path):
    x0 = _beam(5000, seed=2)
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_push_worker, args=(world, x0, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok_{r}.npy").exists() for r in range(world))


def test_plan_invariants_on_a_regular_lattice_with_ties():
    """An un-jittered lattice puts whole planes of particles on one coordinate: a cut falls on a tie.  Ownership stays a partition,
    ghost layers still cover the neighbourhoods, and particles exactly 2h from a cut (not neighbours of anything across it: the
    support test is strict, sim.py:139-141) may or may not be ghosts without harm."""
    s = 0.5 * H
    ax = np.arange(40) * s
    ay = np.arange(9) * s
    x0 = np.stack(np.meshgrid(ax, ay, ay, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) + np.float32([0.0, 0.05, 0.0])
    for world in (2, 3):
        part = SlabPartition.build(x0, H, world)
        owned_all = np.concatenate([p.owned for p in part.plans])
        assert np.array_equal(np.sort(owned_all), np.arange(len(x0)))
        nb = _neighbours(x0)
        for p in part.plans:
            local = set(p.local_ids.tolist())
            layer1 = set(p.ghosts[p.ghost_layer == 1].tolist())
            owned_set = set(p.owned.tolist())
            for i in p.owned[:: max(1, len(p.owned) // 300)]:
                assert all((j in owned_set) or (j in layer1) for j in nb[i])
            for i in list(layer1)[:: max(1, len(layer1) // 300)]:
                assert all(j in local for j in nb[i])
            for q, ids in p.recv.items():
                assert np.array_equal(p.local_ids[ids], part.plans[q].owned[part.plans[q].send[p.rank]])


def test_cost_weighted_cuts_shift_particles_away_from_loaded_ranks():
    """SlabPartition.build(extra_cost=...): owned_r + extra_cost_r is equalised; every particle still has exactly one owner and the
    ghost layers still cover the support radius."""
    rng = np.random.default_rng(3)
    x0 = rng.uniform(0.0, 1.0, (60000, 3)).astype(np.float32)
    x0[:, 0] *= 4.0
    base = SlabPartition.build(x0, 0.007, 4)
    w = SlabPartition.build(x0, 0.007, 4, extra_cost=[0, 4000, 4000, 0])
    assert [p.n_owned for p in base.plans] == [15000] * 4
    assert [p.n_owned for p in w.plans] == [17000, 13000, 13000, 17000]
    owner = np.full(len(x0), -1)
    for p in w.plans:
        assert (owner[p.owned] == -1).all()
        owner[p.owned] = p.rank
    assert (owner >= 0).all()
    for p in w.plans:                         # every send list has a matching receive list of the same length
        for q, ids in p.send.items():
            assert len(ids) == len(w.plans[q].recv[p.rank])


def test_pair_work_weights_move_the_cuts_towards_the_full_sections():
    """neighbour_weights: 27-cell occupancy per particle; with it the tapered end slabs of an ellipsoid own more particles and every
    rank carries the same estimated pair work."""
    from meshless_inflatable_softbody_b200 import scenes
    from meshless_inflatable_softbody_b200.slab import neighbour_weights
    x0 = scenes.jittered_ellipsoid(120_000, seed=1, aspect=(6.0, 1.0, 1.0)).astype(np.float32)
    w = neighbour_weights(x0, 0.007)
    assert w.shape == (len(x0),) and w.min() >= 1
    # brute-force check of the estimate on a few particles
    cw = 2 * float(np.float32(0.007))
    cells = np.floor(x0.astype(np.float64) / cw).astype(np.int64)
    for i in (0, 1234, len(x0) - 1):
        assert w[i] == np.sum(np.all(np.abs(cells - cells[i]) <= 1, axis=1))
    a = SlabPartition.build(x0, 0.007, 4)
    b = SlabPartition.build(x0, 0.007, 4, weights=w)
    na, nb = [p.n_owned for p in a.plans], [p.n_owned for p in b.plans]
    assert max(na) - min(na) <= 1
    assert nb[0] > nb[1] and nb[3] > nb[2] and sum(nb) == len(x0)
    work = np.array([w[p.owned].sum() for p in b.plans])
    assert work.max() / work.min() < 1.002
