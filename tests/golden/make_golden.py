#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference; the GPU box has neither).

    python tests/golden/make_golden.py [--ref /root/reference] [--jobs 8]

What it does (level "L0 = true reference"):
  * copies the reference tree to a scratch directory (the reference writes caches next to
    its package, and /root/reference is read-only),
  * injects stub modules for matplotlib / progressbar (not installed here) and restores the
    numpy aliases np.float / np.int the reference still uses (EFEM.py:41, mesh.py:246),
  * for MESH_21_21 and MESH_50_50: init_mesh() -> MeshObj -> saves the mesh data contract
    (nodes, elem, elem_param, electrode_mesh, detection_index),
  * MESH_21_21: the full EJAC.JAC_calculation() (16 x 882 forward solves, sharded over
    processes by calc_from/calc_end exactly like the reference's manual distribution),
    the 16 baseline solves, Solver inv_mat for lambda=289, solve() on a seeded frame,
  * MESH_50_50: dense K for the zero base, baseline solves, 96 sampled (i, j) Jacobian columns
    (electrode-interior, electrode-adjacent, measuring-electrode and bulk elements),
  * one Example_02-style forward solve with a non-zero object (complex K) per mesh,
  * a non-zero `overall_origin_variable` Jacobian sample on MESH_21_21.
Outputs: tests/golden/mesh_21_21.npz, mesh_50_50.npz, ref_21_21.npz, ref_50_50.npz
"""
import argparse
import json
import multiprocessing as mp
import os
import shutil
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def _install_shims():
    import numpy as np
    np.float = float
    np.int = int
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.font_manager"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    mpl.patches = sys.modules["matplotlib.patches"]
    mpl.font_manager = sys.modules["matplotlib.font_manager"]
    pb = types.ModuleType("progressbar")
    pb.progressbar = lambda it, *a, **k: it
    sys.modules["progressbar"] = pb


def _make_tree(ref, folder, overrides=None):
    """Private writable copy of the reference with config.json folder_name switched."""
    d = tempfile.mkdtemp(prefix="ceit_ref_")
    for item in ("CEIT", "config.json", folder):
        s = os.path.join(ref, item)
        t = os.path.join(d, item)
        if os.path.isdir(s):
            shutil.copytree(s, t)
        else:
            shutil.copy(s, t)
    for root, _, files in os.walk(d):
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    with open(os.path.join(d, "config.json")) as f:
        cfg = json.load(f)
    cfg["folder_name"] = folder
    cfg.update(overrides or {})
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    return d


def _import_ref(tree):
    _install_shims()
    for k in [k for k in sys.modules if k == "CEIT" or k.startswith("CEIT.")]:
        del sys.modules[k]
    sys.path.insert(0, tree)
    os.chdir(tree)
    import CEIT  # noqa: F401
    return tree


def _jac_rows_worker(args):
    """One process = a [calc_from, calc_end) block of excitations (EJAC.py:107-119)."""
    ref, folder, cache_src, lo, hi, overrides = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import numpy as np
    ov = dict(overrides or {})
    ov.update({"calc_from": lo, "calc_end": hi})
    tree = _make_tree(ref, folder, ov)
    for f in os.listdir(cache_src):
        if f.startswith("cache_Mesh") or f.startswith("Mesh_Cache"):
            shutil.copy(os.path.join(cache_src, f), os.path.join(tree, folder, f))
    _import_ref(tree)
    from CEIT.EJAC import EJAC
    ej = EJAC()
    if lo > 0:
        ej.save_JAC_np()  # JAC_calculation loads the cache first when calc_from > 0 (EJAC.py:111-112)
    t0 = time.time()
    J = ej.JAC_calculation()
    dt = time.time() - t0
    L = ej.electrode_num
    out = np.array(J[lo * (L - 1): hi * (L - 1)])
    shutil.rmtree(tree, ignore_errors=True)
    return lo, hi, out, dt


def _cols_worker(args):
    """Sampled Jacobian columns: perturb element j, EFEM.calculation(i) (EJAC.py:123-133)."""
    ref, folder, cache_src, pairs, c, overrides = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import numpy as np
    tree = _make_tree(ref, folder, overrides)
    for f in os.listdir(cache_src):
        if f.startswith("cache_Mesh") or f.startswith("Mesh_Cache"):
            shutil.copy(os.path.join(cache_src, f), os.path.join(tree, folder, f))
    _import_ref(tree)
    from CEIT.EFEM import EFEM
    from CEIT.models.mesh import MeshObj
    from CEIT.util.utilities import get_config
    cfg = get_config()
    fem = EFEM(MeshObj())
    fem.reset_variable(overall_variable=cfg["overall_origin_variable"])
    L = fem.mesh.electrode_num
    res = []
    for (i, j) in pairs:
        fem.elem_variable[j] += c
        _, _, ep = fem.calculation(i)
        fem.elem_variable[j] -= c
        res.append(np.array([np.abs(ep[m]) for m in range(L) if m != i]))
    shutil.rmtree(tree, ignore_errors=True)
    return pairs, np.array(res)


def build_mesh_fixture(ref, folder, out_npz):
    import numpy as np
    tree = _make_tree(ref, folder)
    _import_ref(tree)
    from CEIT.readmesh import init_mesh
    from CEIT.models.mesh import MeshObj
    t0 = time.time()
    init_mesh()
    t_init = time.time() - t0
    mesh = MeshObj()
    L = mesh.electrode_num
    el_ptr = np.zeros(L + 1, dtype=np.int64)
    el_idx = []
    for i in range(L):
        el_idx += list(mesh.electrode_mesh[i])
        el_ptr[i + 1] = len(el_idx)
    np.savez_compressed(
        out_npz,
        nodes=np.asarray(mesh.nodes, dtype=np.float64),
        elem=np.asarray(mesh.elem, dtype=np.int64),
        elem_perm=np.asarray(mesh.elem_perm, dtype=np.float64),
        elem_param=np.asarray(mesh.elem_param, dtype=np.float64),
        electrode_ptr=el_ptr, electrode_elems=np.asarray(el_idx, dtype=np.int64),
        detection_index=np.asarray(mesh.detection_index, dtype=np.int64),
        detection_elem=np.asarray(mesh.detection_elem, dtype=np.int64),
        electrode_centers=np.asarray(mesh.electrode_center_list, dtype=np.float64),
        electrode_radius=np.float64(mesh.electrode_radius),
        init_mesh_seconds=np.float64(t_init),
    )
    return tree, mesh


def sample_pairs(mesh, n_bulk, seed):
    """(i, j) pairs covering the edge cases SURVEY.md §8c lists."""
    import numpy as np
    rng = np.random.default_rng(seed)
    L = mesh.electrode_num
    E = len(mesh.elem)
    el_nodes = [set(int(n) for e in mesh.electrode_mesh[i] for n in mesh.elem[e]) for i in range(L)]
    pairs = []
    for i in (0, 2, 5, 11):
        inside = list(mesh.electrode_mesh[i])
        pairs += [(i, int(inside[0])), (i, int(inside[-1]))]            # wholly Dirichlet: column == baseline
        adj = [e for e in range(E) if e not in inside and any(int(n) in el_nodes[i] for n in mesh.elem[e])]
        one = [e for e in adj if sum(int(n) in el_nodes[i] for n in mesh.elem[e]) == 1]
        two = [e for e in adj if sum(int(n) in el_nodes[i] for n in mesh.elem[e]) == 2]
        pairs += [(i, int(one[0]))] if one else []
        pairs += [(i, int(two[0]))] if two else []
        m = (i + 3) % L
        pairs += [(i, int(mesh.electrode_mesh[m][0])), (i, int(mesh.electrode_mesh[m][-1]))]  # inside a measuring electrode
    for _ in range(n_bulk):
        pairs.append((int(rng.integers(L)), int(rng.integers(E))))
    return pairs


def main():
    import numpy as np
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    ctx = mp.get_context("spawn")

    for folder, tag in (("MESH_21_21", "21_21"), ("MESH_50_50", "50_50")):
        if a.only and a.only != tag:
            continue
        t_all = time.time()
        tree, mesh = build_mesh_fixture(a.ref, folder, os.path.join(HERE, f"mesh_{tag}.npz"))
        cache_src = os.path.join(tree, folder)
        from CEIT.EFEM import EFEM
        from CEIT.util.utilities import get_config
        cfg = get_config()
        L = mesh.electrode_num
        E = len(mesh.elem)
        c = cfg["variable_change_for_JAC"]
        out = {}
        # --- baseline solves, dense K (zero base) -------------------------------------------
        fem = EFEM(mesh)
        fem.reset_variable(0)
        base_node, base_elec, t_fwd = [], [], []
        for i in range(L):
            t0 = time.time()
            npot, _, ep = fem.calculation(i)
            t_fwd.append(time.time() - t0)
            base_node.append(np.array(npot, dtype=np.float64))
            base_elec.append(np.array(ep))
        out["base_node_potential"] = np.array(base_node)
        out["base_electrode_potential"] = np.array(base_elec)
        out["forward_seconds"] = np.array(t_fwd)
        fem.calc_init()
        fem.construct_sparse_matrix()
        import scipy.sparse as sp
        K = sp.csr_matrix(fem.K_sparse)
        K.sort_indices()
        out["K_indptr"], out["K_indices"], out["K_data"] = K.indptr.astype(np.int64), K.indices.astype(np.int64), K.data
        # --- Example_02-style complex forward solve ------------------------------------------
        fem.reset_variable(0)
        fem.change_add_variable_geometry([-20 / 1000, -10 / 1000], 10 / 1000, 1e-7, "square")
        out["obj_variable"] = np.array(fem.elem_variable)
        npot, epot, ep = fem.calculation(2)
        out["obj_node_potential"] = np.array(npot, dtype=np.float64)
        out["obj_element_potential"] = np.array(epot)
        out["obj_electrode_potential"] = np.array(ep)
        fem.reset_variable(0)
        # --- sampled columns (both meshes) ---------------------------------------------------
        pairs = sample_pairs(mesh, 60 if tag == "50_50" else 40, seed=7)
        chunks = [pairs[k::a.jobs] for k in range(a.jobs)]
        with ctx.Pool(a.jobs) as pool:
            res = pool.map(_cols_worker, [(a.ref, folder, cache_src, ch, c, None) for ch in chunks if ch])
        sp_pairs, sp_vals = [], []
        for p, v in res:
            sp_pairs += p
            sp_vals.append(v)
        out["sample_pairs"] = np.array(sp_pairs, dtype=np.int64)
        out["sample_cols"] = np.concatenate(sp_vals, axis=0)
        if tag == "21_21":
            # --- full Jacobian, sharded by excitation like the reference's manual scheme ------
            blocks = [(a.ref, folder, cache_src, lo, min(lo + 2, L), None) for lo in range(0, L, 2)]
            with ctx.Pool(a.jobs) as pool:
                res = pool.map(_jac_rows_worker, blocks)
            J = np.zeros((L * (L - 1), E))
            tot = 0.0
            for lo, hi, rows, dt in res:
                J[lo * (L - 1): hi * (L - 1)] = rows
                tot += dt
            out["JAC"] = J
            out["JAC_cpu_seconds_1thread_sum"] = np.float64(tot)
            # --- Solver: inv_mat for lambda=289 and a seeded frame ---------------------------
            np.save(os.path.join(cache_src, "JAC_cache.npy"), J)
            from CEIT.Solver import Solver
            if os.path.exists(os.path.join(cache_src, "inv_mat.npy")):
                os.remove(os.path.join(cache_src, "inv_mat.npy"))
            s = Solver(289)
            out["inv_mat_289"] = np.array(s.inv_mat)
            dV = np.random.default_rng(0).random(L * (L - 1))
            out["solve_dV"] = dV
            out["solve_out"] = np.array(s.solve(dV))
            # small lambda (ill-conditioned regime)
            s.get_inv_matrix(2e-5)
            out["inv_mat_2e-5"] = np.array(s.inv_mat)
            # --- non-zero base variable: complex K, phi0 != 1 ---------------------------------
            ov = {"overall_origin_variable": 3e-8}
            pairs_nz = sample_pairs(mesh, 16, seed=11)
            chunks = [pairs_nz[k::a.jobs] for k in range(a.jobs)]
            with ctx.Pool(a.jobs) as pool:
                res = pool.map(_cols_worker, [(a.ref, folder, cache_src, ch, c, ov) for ch in chunks if ch])
            p2, v2 = [], []
            for p, v in res:
                p2 += p
                v2.append(v)
            out["nz_base_variable"] = np.float64(3e-8)
            out["nz_sample_pairs"] = np.array(p2, dtype=np.int64)
            out["nz_sample_cols"] = np.concatenate(v2, axis=0)
        out["variable_change_for_JAC"] = np.float64(c)
        out["signal_frequency"] = np.float64(cfg["signal_frequency"])
        out["resistance"] = np.float64(cfg["resistance"])
        out["detection_bound"] = np.float64(cfg["detection_bound"])
        np.savez_compressed(os.path.join(HERE, f"ref_{tag}.npz"), **out)
        print(f"[{tag}] done in {time.time() - t_all:.1f}s", flush=True)
        sys.path.remove(tree)


if __name__ == "__main__":
    main()
