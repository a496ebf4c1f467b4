"""The batched realtime-reconstruction GEMM alone (for an ncu capture of gemm_dmma_kernel<128,128,...>)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ceit_b200 import _lib
n_det, m, F = 4050, 240, 100000
g = torch.Generator(device="cuda").manual_seed(0)
inv = torch.rand((n_det, m), dtype=torch.float64, device="cuda", generator=g)
dV = torch.rand((m, F), dtype=torch.float64, device="cuda", generator=g)
out = torch.empty((n_det, F), dtype=torch.float64, device="cuda")
for _ in range(3):
    _lib.reconstruct(inv, dV, out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); _lib.reconstruct(inv, dV, out); b.record(); torch.cuda.synchronize()
print("ms", a.elapsed_time(b), "TFLOP/s", 2.0 * n_det * m * F / a.elapsed_time(b) / 1e9)
