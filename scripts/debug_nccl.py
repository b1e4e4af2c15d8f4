import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from meshless_inflatable_softbody_b200 import SceneConfig, scenes, Simulator
from meshless_inflatable_softbody_b200.slab import SlabSimulator

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
x0 = scenes.jittered_beam(8000, seed=0, aspect=(6.0, 1.0, 1.0), centre=(0.0, 0.012, 0.0)); x0[:, 1] += 0.0006 - x0[:, 1].min()
x0 = x0.astype(np.float32)
# single-domain on this device
one = Simulator(x0, SceneConfig(), device=f"cuda:{rank}")
one.startup(); one.step(20)
x1, v1 = one.position_velocity()
print(rank, "single-domain finite", torch.isfinite(x1).all().item(), flush=True)
sim = SlabSimulator(x0, SceneConfig(), rank=rank, world_size=world, device=f"cuda:{rank}")
print(rank, "owned", sim.n_owned, "local", sim.sim.n, {q: len(v) for q, v in sim._send.items()}, {q: len(v) for q, v in sim._recv.items()}, flush=True)
sim.sim.startup(); sim.sim.step(0)
for q, ids in sim._recv.items():
    before = sim.sim.gather_next_positions(ids).clone()
sim._exchange()
torch.cuda.synchronize()
for q, ids in sim._recv.items():
    after = sim.sim.gather_next_positions(ids)
    # expected: x0 + dt*v0 + ... close to x0 of those particles
    gl = sim.plan.local_ids[ids.cpu().numpy()]
    exp = torch.as_tensor(x0[gl], device=after.device)
    print(rank, "ghost next - x0: max", (after - exp).abs().max().item(), " before:", (before - exp).abs().max().item(), "finite", torch.isfinite(after).all().item(), flush=True)
for k in range(5):
    sim.step(1)
    x, v = sim.position_velocity()
    print(rank, "step", k, "finite", torch.isfinite(x).all().item(), "dv", (v - v.mean(0)).abs().max().item(), flush=True)
dist.barrier()
dist.destroy_process_group()
