"""Where the end-to-end (class API) step spends its time: cProfile of K e2e steps on MESH_50_50."""
import sys, os, time, tempfile, cProfile, pstats
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ceit_b200 import meshgen
mesh = bench.load_mesh("50_50")
td = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
cfg = dict(meshgen.REFERENCE_CONFIG)
os.environ["CEIT_B200_CONFIG"] = meshgen.make_workspace(td, mesh["nodes"], mesh["elem"], cfg)
from ceit_b200.EJAC import EJAC
jac = EJAC()
def step():
    jac.fwd_FEM.invalidate()
    jac.JAC_calculation()
    jac.save_inv_matrix(289)
for _ in range(4): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 50 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
