"""Time a few GEMM shapes of the path with the library's DMMA kernel vs torch (cuBLAS)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ceit_b200 import _lib
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
g = torch.Generator(device="cuda").manual_seed(0)
shapes = [("gram 992 x 257762 (A A^T)", 992, 992, 257762, False, True), ("inv  257762 x 992 x 992 (A^T B)", 257762, 992, 992, True, False),
          ("YT   160000 x 544 x 544", 160000, 544, 544, False, False), ("rec  4050 x 100000 x 240", 4050, 100000, 240, False, False)]
for name, M, N, K, ta, tb in shapes:
    A = torch.randn((K, M) if ta else (M, K), dtype=torch.float64, device="cuda", generator=g)
    B = A if (name.startswith("gram")) else torch.randn((N, K) if tb else (K, N), dtype=torch.float64, device="cuda", generator=g)
    C = torch.empty((M, N), dtype=torch.float64, device="cuda")
    for tile in (0, 1, 2, 3):
        _lib.set_option("gemm_tile", tile)
        ms = t(lambda: _lib.dgemm(A, B, ta, tb, C=C))
        print(f"{name:36s} tile={tile} {ms:8.3f} ms {2.0*M*N*K/ms/1e9:7.2f} TFLOP/s")
    _lib.set_option("gemm_tile", 0)
    ms = t(lambda: torch.matmul(A.T if ta else A, B.T if tb else B, out=C))
    print(f"{name:36s} cuBLAS {ms:8.3f} ms {2.0*M*N*K/ms/1e9:7.2f} TFLOP/s")
