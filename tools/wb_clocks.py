"""Phase cycle counts of the Woodbury kernels for two sample CTAs (needs a build with CEIT_TILE_CLOCKS=1).

    CEIT_TILE_CLOCKS=1 python -m ceit_b200.build --force && python tools/wb_clocks.py [patch]
"""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from ceit_b200 import _lib
import bench
L_ = _lib.lib()
mesh = bench.load_mesh('50_50'); N, E, L = len(mesh['nodes']), len(mesh['elem']), 16
omega = 2 * np.pi * 2e4
if len(sys.argv) > 1 and sys.argv[1] == 'patch':
    L_.ceit_set_option(b"no_lowrank_kernel", 1)
plan = _lib.Plan(mesh['nodes'], mesh['elem'], mesh['electrodes'])
perm = torch.tensor(mesh['elem_perm'], device='cuda'); var = torch.zeros(E, dtype=torch.float64, device='cuda')
J = torch.zeros((240, E), dtype=torch.float64, device='cuda')
plan.assemble(perm, var, omega); plan.factor(False)
for it in range(3):
    plan.jacobian_block(0, L, 0, E, 1e-3, omega, J, E); torch.cuda.synchronize()
    out = np.zeros(16, dtype=np.int64); L_.ceit_debug_wb_clocks(ctypes.c_void_p(out.ctypes.data))
    for base, name in ((0, 'CTA 81'), (8, 'CTA 1500')):
        d = np.diff(out[base:base + 8])
        print(name, 'phase cycles', d.tolist(), 'total', int(out[base + 7] - out[base]), plan.info()[14:16])
