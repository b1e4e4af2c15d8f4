"""Generate tests/golden/deepsdf_seed0.npz from the REFERENCE's own deepsdf.py (run in the build container,
where /root/reference exists; the fixture travels, the reference does not).

    python tests/golden/make_deepsdf_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
import deepsdf as ref_deepsdf   # noqa: E402  (the reference module, deepsdf.py:1-41)

torch.manual_seed(0)
torch.set_num_threads(1)
model = ref_deepsdf.DeepSDFWithCode().to("cpu")
state = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
rng = np.random.default_rng(0)
pts = rng.uniform(-0.06, 0.06, size=(256, 3)).astype(np.float32)     # model-space metres (points * 0.01, sim.py:47-48)
pts[128:] = rng.uniform(-1.0, 1.0, size=(128, 3)).astype(np.float32)  # wide inputs: exercise the ReLU masks
with torch.no_grad():
    out = model(torch.from_numpy(pts)).numpy().astype(np.float32)     # [256,1], sim.py:100
    out64 = model.double()(torch.from_numpy(pts).double()).numpy()
keys = sorted(state)
np.savez_compressed(
    os.path.join(os.path.dirname(os.path.abspath(__file__)), "deepsdf_seed0.npz"),
    points=pts, sdf=out, sdf64=out64,
    keys=np.array(keys),
    shapes=np.array([str(tuple(state[k].shape)) for k in keys]),
    sums=np.array([state[k].astype(np.float64).sum() for k in keys]),
    abs_sums=np.array([np.abs(state[k].astype(np.float64)).sum() for k in keys]),
    first_row=state["network.3.parametrizations.weight.original1"][0].astype(np.float32),
    torch_version=np.array(torch.__version__),
)
print("wrote deepsdf_seed0.npz:", out[:4, 0], "keys", len(keys))
