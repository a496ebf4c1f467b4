"""Experiment: capture one full hot-path step (assemble, factor, Jacobian, inverse model) into a CUDA graph and
compare replay time with eager stream launches.  usage: python tools/graph_step.py [50_50|21_21]"""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ceit_b200 import _lib
import bench
tag = sys.argv[1] if len(sys.argv) > 1 else "50_50"
mesh = bench.load_mesh(tag)
N, E, L = len(mesh["nodes"]), len(mesh["elem"]), len(mesh["electrodes"])
omega = 2 * np.pi * 2e4
plan = _lib.Plan(mesh["nodes"], mesh["elem"], mesh["electrodes"])
perm = torch.tensor(mesh["elem_perm"], device="cuda"); var = torch.zeros(E, dtype=torch.float64, device="cuda")
det = torch.tensor(mesh["detection_index"], device="cuda")
J = torch.zeros((L * (L - 1), E), dtype=torch.float64, device="cuda"); base = torch.zeros(L * (L - 1), dtype=torch.float64, device="cuda")
def step():
    plan.assemble(perm, var, omega); plan.factor(False)
    plan.jacobian_block(0, L, 0, E, 1e-3, omega, J, E, base)
    return _lib.inverse_model(J, det, 289)
def timeit(fn, reps=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
inv_eager = step().clone()
print("eager ms/step", timeit(step))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    inv_g = step()
torch.cuda.synchronize()
print("graph ms/step", timeit(g.replay))
print("max |inv_graph - inv_eager|", float((inv_g - inv_eager).abs().max()))
