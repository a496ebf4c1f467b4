"""Phase cycle counts of potrf_trtri_tile_kernel (needs a build with -DCEIT_TILE_CLOCKS)."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0,'.')
from ceit_b200 import _lib
import bench
L=_lib.lib()
n=int(sys.argv[1]) if len(sys.argv)>1 else 240
# SPD matrix -> inverse model path uses potrf_trtri_blocked with n=240 (4 tiles of 64,64,64,48)
g=torch.Generator(device='cuda').manual_seed(0)
J=1+1e-8*torch.rand((n,500),dtype=torch.float64,device='cuda',generator=g)
det=torch.arange(500,device='cuda')
for mode in (0,):
  for it in range(3):
      inv=_lib.inverse_model(J,det,289.0,form=2); torch.cuda.synchronize()
      out=np.zeros(8,dtype=np.int64); L.ceit_debug_tile_clocks(ctypes.c_void_p(out.ctypes.data))
      d=np.diff(out[:7]); print('phases cycles: load',d[0],'elim',d[1],'scale',d[2],'inv16',d[3],'recur',d[4],'store',d[5],'total',out[6]-out[0])
