"""The dense kernels of the inverse-model / realtime legs alone (for ncu captures):
  * batched realtime reconstruction 4050 x 100000 x 240 (gemm_dmma_kernel<64,64>, opA = N, opB = N),
  * the 400x400-mesh Gram matrix 992 x 992 x 257762 (split-K, lower-only, opB = T) and inverse-model product,
  * the single-frame GEMV (gemv_rows_kernel)."""
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
print("reconstruct ms", a.elapsed_time(b), "TFLOP/s", 2.0 * n_det * m * F / a.elapsed_time(b) / 1e9)
one = dV[:, 0].contiguous(); o1 = torch.empty(n_det, dtype=torch.float64, device="cuda")
for _ in range(3):
    _lib.reconstruct(inv, one, o1)
torch.cuda.synchronize()
del dV, out
if "--big" in sys.argv:
    m2, nd2 = 992, 257762
    J = 1.0 - 1e-8 * torch.rand((m2, nd2), dtype=torch.float64, device="cuda", generator=g)
    det = torch.arange(nd2, dtype=torch.int64, device="cuda")
    for _ in range(2):
        r = _lib.inverse_model(J, det, 289.0)
    torch.cuda.synchronize()
    a.record(); r = _lib.inverse_model(J, det, 289.0); b.record(); torch.cuda.synchronize()
    print("inverse model 992 x 257762 ms", a.elapsed_time(b))
