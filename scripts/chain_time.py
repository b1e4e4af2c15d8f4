"""Device time of one DeepSDF forward (layer 0 + hidden layers + last layer) by row count and hidden-layer kernel path. Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ctypes as C
import bench
from meshless_inflatable_softbody_b200 import native
from meshless_inflatable_softbody_b200 import DeepSDF
net = DeepSDF(bench.obstacle_state())
rng = np.random.default_rng(0)
for path in (1, 3):
    net.set_gemm_path(path)
    for rows in (64, 128, 256, 384, 512, 1024):
        p = torch.as_tensor(rng.uniform(-0.02, 0.02, size=(rows, 3)).astype(np.float32), device="cuda")
        out = torch.empty(rows, device="cuda", dtype=torch.float32)
        st = C.c_void_p(net.stream.cuda_stream)
        def fwd():
            native.check(net.L.mis_sdf_query(net._h, p.data_ptr(), rows, None, out.data_ptr(), None, 0.0, st), "mis_sdf_query")
        torch.cuda.synchronize()
        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(net.stream):
            e0.record()
            for _ in range(20):
                fwd()
            e1.record()
        net.stream.synchronize()
        print("path %d rows %4d: %.1f us per forward" % (path, rows, 1e3 * e0.elapsed_time(e1) / 20), flush=True)
