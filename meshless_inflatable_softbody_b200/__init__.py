"""B200-native per-step meshless particle update (the sim.py / sim_taichi.py hot path).

Host side is Python/PyTorch (device memory, streams, torch.distributed); all particle
arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI of include/mis.h.
There is no CPU fallback: constructing a Simulator without a CUDA device or without the
built library fails loudly.
"""
import importlib

from .config import SceneConfig  # noqa: F401
from . import scenes  # noqa: F401

__all__ = ["SceneConfig", "scenes", "Simulator", "DeepSDF", "native"]


def __getattr__(name):
    if name == "native":
        return importlib.import_module(".native", __name__)
    if name == "Simulator":
        return importlib.import_module(".simulator", __name__).Simulator
    if name == "DeepSDF":
        return importlib.import_module(".deepsdf", __name__).DeepSDF
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
