"""Phase stamps of the persistent Woodbury kernel (CTA 7): consumer group 0 and the producer, per item.

    CEIT_TILE_CLOCKS=1 python -m ceit_b200.build --force && CEIT_TILE_CLOCKS=1 python tools/wbq_clocks.py
"""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from ceit_b200 import _lib
import bench
L_ = _lib.lib()
mesh = bench.load_mesh('50_50'); N, E, L = len(mesh['nodes']), len(mesh['elem']), 16
omega = 2 * np.pi * 2e4
plan = _lib.Plan(mesh['nodes'], mesh['elem'], mesh['electrodes'])
perm = torch.tensor(mesh['elem_perm'], device='cuda'); var = torch.zeros(E, dtype=torch.float64, device='cuda')
J = torch.zeros((240, E), dtype=torch.float64, device='cuda')
plan.assemble(perm, var, omega); plan.factor(False)
for it in range(2):
    plan.jacobian_block(0, L, 0, E, 1e-3, omega, J, E); torch.cuda.synchronize()
out = np.zeros(512, dtype=np.int64); L_.ceit_debug_wbq_clocks(ctypes.c_void_p(out.ctypes.data))
c = out.reshape(32, 16)
t0 = c[c > 0].min()
print("consumer group 0 (items 0,2,4,...): start, topo, wait_full, corr, solve, main  | producer: begin, empty_wait, issued")
for it in range(22):
    a = c[it]
    cons = (a[:6] - t0).tolist() if a[0] else None
    prod = (a[8:11] - t0).tolist() if a[8] else None
    print(it, "cons", cons, "d", np.diff(a[:6]).tolist() if a[0] else None, "prod", prod)
