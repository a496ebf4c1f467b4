"""Phase cycle counts of woodbury_patch_kernel (needs a build with CEIT_TILE_CLOCKS=1)."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0,'.')
from ceit_b200 import _lib
import bench
L_=_lib.lib()
mesh=bench.load_mesh('50_50'); N,E,L=len(mesh['nodes']),len(mesh['elem']),16
omega=2*np.pi*2e4
plan=_lib.Plan(mesh['nodes'],mesh['elem'],mesh['electrodes'])
perm=torch.tensor(mesh['elem_perm'],device='cuda'); var=torch.zeros(E,dtype=torch.float64,device='cuda')
J=torch.zeros((240,E),dtype=torch.float64,device='cuda')
plan.assemble(perm,var,omega); plan.factor(False)
for it in range(3):
    plan.jacobian_block(0,L,0,E,1e-3,omega,J,E); torch.cuda.synchronize()
    out=np.zeros(8,dtype=np.int64); L_.ceit_debug_wb_clocks(ctypes.c_void_p(out.ctypes.data))
    d=np.diff(out[:7]); print('cycles: issue-copies',d[0],'wait-copies',d[1],'pass1',d[2],'solve(w0)+sync',d[3],'sync',d[4],'pass2',d[5],'total',out[6]-out[0], plan.info()[14:16])
