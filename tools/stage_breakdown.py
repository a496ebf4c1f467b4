"""Stage timing (CUDA events inside the library) of the hot path on a shipped or synthetic mesh.
usage: python tools/stage_breakdown.py [50_50|21_21|<n> [per_side]] [--target B] [--complex] [--opt name=value ...]"""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from ceit_b200 import _lib, meshgen
import bench
opts = [sys.argv[k + 1] for k, a in enumerate(sys.argv[:-1]) if a == "--opt"]     # --opt name=value (ceit_set_option)
for o_ in opts:
    _lib.set_option(o_.split("=")[0], float(o_.split("=")[1]))
args = [a for k, a in enumerate(sys.argv[1:], 1) if not a.startswith("--") and a not in opts and sys.argv[k - 1] != "--target"]
cplx = False
target = int(sys.argv[sys.argv.index("--target") + 1]) if "--target" in sys.argv else 0
tag = args[0] if args else "50_50"
if tag in ("50_50", "21_21"):
    mesh = bench.load_mesh(tag)
    nodes, elem, els, perm_h = mesh["nodes"], mesh["elem"], mesh["electrodes"], mesh["elem_perm"]
    det_h = mesh["detection_index"]
else:
    import ceit_oracle as o
    n = int(tag); per_side = int(args[1]) if len(args) > 1 else 5
    cfg = meshgen.synthetic_config(n, per_side)
    nodes, elem = meshgen.structured_mesh(n)
    param = o.elem_param(nodes, elem)
    els = o.electrode_elements(param, np.array(cfg["electrode_centers"]) / 1000, cfg["electrode_radius"] / 1000)
    perm_h = np.ones(len(elem)); det_h = o.detection_index(nodes, elem, cfg["detection_bound"] / 1000)
N, E, L = len(nodes), len(elem), len(els)
omega = 2 * np.pi * 2e4
t0 = time.time(); plan = _lib.Plan(nodes, elem, els, target_block=target); print("plan s", time.time() - t0, plan.info()[:12])
perm = torch.tensor(perm_h, device="cuda"); var = torch.zeros(E, dtype=torch.float64, device="cuda")
det = torch.tensor(det_h, device="cuda")
J = torch.zeros((L * (L - 1), E), dtype=torch.float64, device="cuda"); base = torch.zeros(L * (L - 1), dtype=torch.float64, device="cuda")
def step(inv=True):
    plan.assemble(perm, var, omega); plan.factor(cplx)
    plan.jacobian_block(0, L, 0, E, 1e-3, omega, J, E, base)
    if inv: return _lib.inverse_model(J, det, 289)
cplx = "--complex" in sys.argv
if cplx: var += 3e-8
big = len(det_h) * L * (L - 1) * 8 > 20e9
step(not big); torch.cuda.synchronize()
_lib.set_option("stage_timers", 1); _lib.stage_times(); l0 = _lib.launch_count()
reps = int(os.environ.get("CEIT_REPS", "3"))
t0 = time.time()
for _ in range(reps): step(not big)
torch.cuda.synchronize(); wall = (time.time() - t0) / reps
st = _lib.stage_times()
print("N", N, "E", E, "L", L, "cols", L * E, "wall ms/step", wall * 1e3, "launches/step", (_lib.launch_count() - l0) / reps)
for k, (ms, n_) in st.items():
    if n_: print(f"  {k:14s} {ms / reps:10.3f} ms/step  ({n_ // reps} timed regions)")
print("  columns/s", L * E / wall, " J-1 range", float((J - 1).min()), float((J - 1).max()), "flag", plan.numeric_flag())
