"""Error levels of the CUDA path against the committed reference outputs (tests/golden): max |J - J_ref| and the
relative error of the signal J - 1, MESH_21_21 (full matrix) and MESH_50_50 (sampled columns)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ceit_b200 import _lib
for o_ in [sys.argv[k + 1] for k, a in enumerate(sys.argv[:-1]) if a == "--opt"]:
    _lib.set_option(o_.split("=")[0], float(o_.split("=")[1]))
G = os.path.join(ROOT, "tests", "golden")
omega = 2 * np.pi * 2e4
for tag in ("21_21", "50_50"):
    m = dict(np.load(os.path.join(G, f"mesh_{tag}.npz"))); r = dict(np.load(os.path.join(G, f"ref_{tag}.npz")))
    ptr, ee = m["electrode_ptr"], m["electrode_elems"]
    els = [ee[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)]
    E, L = len(m["elem"]), len(els)
    plan = _lib.Plan(m["nodes"], m["elem"], els)
    perm = torch.tensor(m["elem_perm"], device="cuda"); var = torch.zeros(E, dtype=torch.float64, device="cuda")
    J = torch.zeros((L * (L - 1), E), dtype=torch.float64, device="cuda")
    plan.assemble(perm, var, omega); plan.factor(False)
    plan.jacobian_block(0, L, 0, E, 1e-3, omega, J, E)
    Jh = J.cpu().numpy()
    if tag == "21_21":
        ref, got = r["JAC"], Jh
    else:
        rows = lambda i: np.arange(i * (L - 1), (i + 1) * (L - 1))
        ref = np.stack([r["sample_cols"][k] for k in range(len(r["sample_pairs"]))])
        got = np.stack([Jh[rows(i), j] for i, j in r["sample_pairs"]])
    err = np.abs(got - ref)
    print(tag, "max|J-Jref| %.3e" % err.max(), " ||dJ||/||Jref-1|| %.3e" % (np.linalg.norm(got - ref) / np.linalg.norm(ref - 1)),
          " max(J-1) %.3e" % (Jh - 1).max())
    plan.close()
