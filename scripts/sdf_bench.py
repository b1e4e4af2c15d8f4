"""Throughput of the DeepSDF tcgen05 chain (run under gpurun): per-layer GEMM time and the full query."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meshless_inflatable_softbody_b200 import DeepSDF
from oracle import deepsdf_oracle as do

st = do.seeded_state(0)
net = DeepSDF(st)
for m in (4096, 16384, 65536):
    ms = net.profile_gemm(m, reps=20)
    flop = 2.0 * m * 1024 * 1024
    print("gemm m=%6d: %.3f ms/layer  %.1f TFLOP/s fp32-equivalent  (%.1f TF/s tf32 issued, 3 products)" % (m, ms, flop / ms / 1e9, 3 * flop / ms / 1e9), flush=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
p = torch.rand((n, 3), device="cuda") * 2 - 1
net.query(p)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(net.stream)
s = net.query(p)
e1.record(net.stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("query n=%d: %.3f ms  %.2e points/s  %.1f TFLOP/s fp32-equivalent" % (n, ms, n / ms * 1e3, 14.7e6 * n / ms / 1e9))
# torch fp32 reference on the same device for scale (library SGEMM, TF32 disabled)
torch.backends.cuda.matmul.allow_tf32 = False
m = do.reference_like_module()
m.load_state_dict({k: torch.as_tensor(v) for k, v in st.items()})
m = m.cuda()
with torch.no_grad():
    m(p); torch.cuda.synchronize()
    t = time.perf_counter(); y = m(p); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("torch fp32 (cuBLAS SGEMM) n=%d: %.3f ms ; max |ours - torch| = %.3e (scale %.3e)" % (n, dt * 1e3, (s - y[:, 0]).abs().max().item(), y.abs().max().item()))
