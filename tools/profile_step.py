"""One warm step + N profiled steps of the hot path (for ncu launch lists): tools/profile_step.py [50_50|21_21|200|400] [steps]"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ceit_b200 import _lib
import bench
tag = sys.argv[1] if len(sys.argv) > 1 else "50_50"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mesh = bench.load_workload(tag) if tag in bench.CONFIGS else bench.load_mesh(tag)
N, E, L = len(mesh["nodes"]), len(mesh["elem"]), len(mesh["electrodes"])
omega = 2 * np.pi * 2e4
plan = _lib.Plan(mesh["nodes"], mesh["elem"], mesh["electrodes"])
perm = torch.tensor(mesh["elem_perm"], device="cuda"); var = torch.zeros(E, dtype=torch.float64, device="cuda")
det = torch.tensor(mesh["detection_index"], device="cuda")
J = torch.zeros((L * (L - 1), E), dtype=torch.float64, device="cuda"); base = torch.zeros(L * (L - 1), dtype=torch.float64, device="cuda")
for s in range(1 + steps):
    plan.assemble(perm, var, omega); plan.factor(False)
    plan.jacobian_block(0, L, 0, E, 1e-3, omega, J, E, base)
    inv = _lib.inverse_model(J, det, 289)
    torch.cuda.synchronize()
    print("step", s, "launches", _lib.launch_count())
