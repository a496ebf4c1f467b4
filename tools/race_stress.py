"""Determinism stress: repeat the step on one plan and compare every result bitwise with the first one.
The algorithm is deterministic (no atomics), so any difference is a race between streams or inside a kernel.
usage: python tools/race_stress.py [tag] [target] [nd] [iters]"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ceit_b200 import _lib
import bench
tag = sys.argv[1] if len(sys.argv) > 1 else "50_50"
target = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nd = int(sys.argv[3]) if len(sys.argv) > 3 else 64
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 200
mesh = bench.load_mesh(tag)
N, E, L = len(mesh["nodes"]), len(mesh["elem"]), len(mesh["electrodes"])
omega = 2 * np.pi * 2e4
_lib.set_option("nd_interior", nd)
plan = _lib.Plan(mesh["nodes"], mesh["elem"], mesh["electrodes"], target_block=target)
_lib.set_option("nd_interior", 64)
perm = torch.tensor(mesh["elem_perm"], device="cuda"); var = torch.zeros(E, dtype=torch.float64, device="cuda")
det = torch.tensor(mesh["detection_index"], device="cuda")
J = torch.zeros((L * (L - 1), E), dtype=torch.float64, device="cuda"); base = torch.zeros(L * (L - 1), dtype=torch.float64, device="cuda")
ref = None; bad = 0
for it in range(iters):
    J.zero_(); base.zero_()
    plan.assemble(perm, var, omega); plan.factor(False)
    plan.jacobian_block(0, L, 0, E, 1e-3, omega, J, E, base)
    inv = _lib.inverse_model(J, det, 289)
    torch.cuda.synchronize()
    cur = (J.clone(), base.clone(), inv.clone())
    if ref is None:
        ref = cur; continue
    d = [(a != b) for a, b in zip(cur, ref)]
    if any(x.any().item() for x in d):
        bad += 1
        dj = d[0]
        rows = dj.any(dim=1).nonzero().flatten().tolist(); cols = dj.any(dim=0).nonzero().flatten().tolist()
        print(f"iter {it}: J differs in {int(dj.sum())} entries, rows {rows[:8]}..({len(rows)}), cols {cols[:8]}..({len(cols)}), "
              f"max |dJ| {float((cur[0] - ref[0]).abs().max()):.3e}; base differs {int(d[1].sum())}; inv differs {int(d[2].sum())}")
print(f"{tag} target={target} nd={nd}: {bad} of {iters - 1} repeats differ from the first step")
