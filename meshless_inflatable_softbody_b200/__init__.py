"""B200-native per-step meshless particle update (the sim.py / sim_taichi.py hot path).

Host side is Python/PyTorch (device memory, streams, torch.distributed); all particle
arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI of include/mis.h.
There is no CPU fallback: importing the simulator without the built library fails loudly.
"""
from .config import SceneConfig  # noqa: F401
from . import scenes  # noqa: F401


def __getattr__(name):
    if name in ("Simulator", "native"):
        from . import simulator, native
        return {"Simulator": simulator.Simulator, "native": native}[name]
    raise AttributeError(name)
