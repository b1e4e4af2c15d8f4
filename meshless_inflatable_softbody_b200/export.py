"""State export consumed by the reference's renderer / target writer.

The reference reads `position[f].numpy()` of every 50th frame into `visualize` (sim.py:334,393-395: the first `out_num`
particles are the vertices of the surface mesh, `r.add_triangle_mesh(vertices=v[:out_num], elements=faces, ...)`) and of every
30th frame into `position_{i}.npy` / `velocity_{i}.npy` (sim.py:363-369).  It can do that because it keeps all 3 001 frames on
the device (sim.py:84-95).  This engine keeps one frame, so the consumers are fed while the rollout runs:

  * `record_frames` streams every k-th frame to pinned host memory through the library's double-buffered export
    (`mis_get_state_host_async`: the un-permute runs on the step stream, the copy on a copy stream, the next steps overlap it);
  * `Simulator.export_targets` writes the `.npy` targets in the reference's naming;
  * `trianglemesh_block` formats one frame as the `Shape "trianglemesh"` block `PbrtRenderer.render` writes
    (pbrt_renderer.py:145-171,252-262), so a scene file can be assembled without the renderer class.  pbrt / ffmpeg
    themselves are not part of this path (and not installed here).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import numpy as np
import torch


def record_frames(sim, frames: int, every: int = 50, out_num: Optional[int] = None, with_velocity: bool = False):
    """Run `frames` steps from the current frame and return the state of frames 0, every, 2*every, ... (< frames), as
    `visualize` samples them (sim.py:393-395): a float32 array (n_frames, m, 3) of positions with m = out_num or n
    (and the velocities, if asked).  Frame f is exported while frames f+1.. are being computed."""
    n = sim.n
    m = n if out_num is None else int(out_num)
    marks = list(range(0, int(frames), int(every)))
    xs = np.empty((len(marks), m, 3), np.float32)
    vs = np.empty((len(marks), m, 3), np.float32) if with_velocity else None
    bufs = [(torch.empty((n, 3), dtype=torch.float32).pin_memory(), torch.empty((n, 3), dtype=torch.float32).pin_memory())
            for _ in range(2)]
    sim.step(0)                                   # frame 0 of the rollout is primed (sim.py:349-351)
    done = 0
    pending: List[int] = []

    def drain(keep: int):
        while len(pending) > keep:
            k = pending.pop(0)
            sim.wait_state_host(len(pending))      # at most `len(pending)` younger exports may still be in flight
            xs[k] = bufs[k & 1][0].numpy()[:m]
            if vs is not None:
                vs[k] = bufs[k & 1][1].numpy()[:m]

    for k, f in enumerate(marks):
        if f > done:
            sim.step(f - done); done = f
        drain(1)                                   # the buffer about to be reused (k - 2) has been consumed
        sim.get_state_host_async(bufs[k & 1][0], bufs[k & 1][1])
        pending.append(k)
    if frames > done:
        sim.step(int(frames) - done)
    drain(0)
    return (xs, vs) if with_velocity else xs


def trianglemesh_block(vertices, elements, texture_coords=None, alpha: float = 1.0, indent: str = "   ") -> str:
    """The `Shape "trianglemesh"` lines of one mesh exactly as PbrtRenderer.render formats them (pbrt_renderer.py:252-262 with
    the properties of add_triangle_mesh, :145-171): integer arrays as ints, real arrays as Python floats, one property per line."""
    v = np.asarray(vertices, dtype=np.float64).ravel()
    e = np.asarray(elements, dtype=np.int64).ravel()
    props = [("integer indices", e), ("point3 P", v)]
    if texture_coords is not None:
        props.append(("point2 uv", np.asarray(texture_coords, dtype=np.float64).ravel()))
    lines = [f"{indent}Shape \"trianglemesh\"\n"]
    for name, arr in props:
        lines.append(f"{indent}    \"{name}\" [" + " ".join(str(x) for x in arr.tolist()) + "]\n")
    lines.append(f"{indent}    \"float alpha\" {float(alpha)}\n")
    return "".join(lines)
