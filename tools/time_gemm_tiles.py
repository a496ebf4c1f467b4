"""Time the DMMA GEMM tile variants (gemm_tile option) on the path's large products:
   reconstruct 4050 x 100000 x 240 (N,N), Y_T 160000 x 544 x 544 (N,N), inverse-model product 257762 x 992 x 992 (T,N),
   Gram 992 x 992 x 257762 (N,T, via ceit_dgemm: no split-K)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ceit_b200 import _lib
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return torch.rand(s, dtype=torch.float64, device="cuda", generator=g)
cases = {
    "reconstruct NN 4050x100000x240": (rnd(4050, 240), rnd(240, 100000), False, False),
    "Y_T NN 160000x544x544": (rnd(160000, 544), rnd(544, 544), False, False),
    "inv TN 257762x992x992": (rnd(992, 257762), rnd(992, 992), True, False),
    "ZW NT 160000x800x544": (rnd(160000, 544), rnd(800, 544), False, True),
}
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, (A, B, ta, tb) in cases.items():
    M = A.shape[1] if ta else A.shape[0]; K = A.shape[0] if ta else A.shape[1]; N = B.shape[0] if tb else B.shape[1]
    C = torch.empty((M, N), dtype=torch.float64, device="cuda")
    ref = None
    for tile in (0, 1, 2, 3):
        _lib.set_option("gemm_tile", tile)
        for _ in range(2): _lib.dgemm(A, B, ta, tb, C=C)
        torch.cuda.synchronize(); a.record()
        for _ in range(5): _lib.dgemm(A, B, ta, tb, C=C)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        if ref is None: ref = C.clone()
        print(f"{name:34s} tile={tile} {ms:8.3f} ms {2.0*M*N*K/ms/1e9:6.2f} TFLOP/s  maxdiff {float((C-ref).abs().max()):.1e}")
    _lib.set_option("gemm_tile", 0)
    Ad = (A.T if ta else A); Bd = (B.T if tb else B)
    for _ in range(2): torch.matmul(Ad, Bd, out=C)
    torch.cuda.synchronize(); a.record()
    for _ in range(5): torch.matmul(Ad, Bd, out=C)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"{name:34s} cuBLAS  {ms:8.3f} ms {2.0*M*N*K/ms/1e9:6.2f} TFLOP/s")
    del A, B, C, ref
