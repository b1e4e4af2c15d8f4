"""Diagnostic: reverse-pass gradient vs central differences under several scene variants (fp64)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from meshless_inflatable_softbody_b200 import Simulator, SceneConfig, scenes
from test_gpu_fp64_and_grad import _grad_scene, _targets

x0, design = _grad_scene()
for name, cfg, frames in (("default", SceneConfig(), 30), ("no ground", SceneConfig(ground_contact=False), 30),
                          ("identity R", SceneConfig(identity_rotation=True), 30), ("symmetric pair", SceneConfig(symmetric_pair=True), 30),
                          ("1 frame", SceneConfig(), 1), ("2 frames", SceneConfig(), 2), ("12 frames", SceneConfig(), 12)):
    nt = 1 if frames < 3 else 3
    targets = _targets(x0, max(frames, 3), nt) if frames >= 3 else _targets(x0, 3, 1)
    sim = Simulator(x0, cfg, precision="f64")
    sim.set_design(design)
    loss, grad = sim.rollout_grad(targets, frames=frames)
    grad = grad.cpu().numpy()
    order = np.argsort(-np.abs(grad))
    errs = []
    for i in [int(order[0]), int(order[3]), int(order[40]), int(order[150])]:
        eps = 1e-5
        d = design.copy(); d[i] += eps; sim.set_design(d); lp, _ = sim.rollout_grad(targets, frames=frames)
        d[i] -= 2 * eps; sim.set_design(d); lm, _ = sim.rollout_grad(targets, frames=frames)
        num = (lp - lm) / (2 * eps)
        errs.append((i, grad[i], num, abs(num - grad[i]) / max(abs(num), 1e-300)))
    print(name, "loss", loss, "|grad|max", np.abs(grad).max())
    for e in errs:
        print("   i=%d adj=%.6e fd=%.6e rel=%.2e" % e)
