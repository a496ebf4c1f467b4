"""Round-2 parity additions (VERDICT r01 'weak' list): branches of the product path that no test reached, full-matrix
checks where only samples existed, numerics of every one-shot Gauss-Newton variant through the class API."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ceit_oracle as o  # noqa: E402

from ceit_b200 import _lib, meshgen  # noqa: E402

pytestmark = pytest.mark.gpu
RAW_TOL = 1e-9
SIGNAL_TOL = 1e-4
MODEL_TOL = 1e-9


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    return torch


@pytest.fixture()
def workspace(tmp_path, golden, monkeypatch):
    def make(tag, **overrides):
        m, _ = golden(tag)
        cfg = dict(meshgen.REFERENCE_CONFIG)
        cfg["folder_name"] = "MESH_" + tag
        cfg.update(overrides)
        path = meshgen.make_workspace(str(tmp_path / tag), m["nodes"], m["elem"], cfg)
        monkeypatch.setenv("CEIT_B200_CONFIG", path)
        monkeypatch.chdir(os.path.dirname(path))
        return path
    return make


def _rows(i, L):
    return slice(i * (L - 1), (i + 1) * (L - 1))


def test_full_mesh_50_50_jacobian_against_woodbury_oracle(torch_cuda, golden):
    """ALL 240 x 5000 entries of the MESH_50_50 Jacobian against the oracle's exact rank-3 restatement
    (oracle.jacobian_woodbury, itself pinned on the reference's full MESH_21_21 Jacobian and on 84 reference columns
    of this mesh in tests/test_oracle_golden.py)."""
    t = torch_cuda
    m, r = golden("50_50")
    N, E, L = len(m["nodes"]), len(m["elem"]), 16
    omega = 2 * np.pi * float(r["signal_frequency"])
    c = float(r["variable_change_for_JAC"])
    p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"])
    p.assemble(t.tensor(m["elem_perm"], device="cuda"), t.zeros(E, dtype=t.float64, device="cuda"), omega)
    p.factor(False)
    J = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
    p.jacobian_block(0, L, 0, E, c, omega, J, E)
    assert p.numeric_flag() == 0
    Jh = J.cpu().numpy()
    param = o.elem_param(m["nodes"], m["elem"])
    Jo = o.jacobian_woodbury(m["elem"], param, m["elem_perm"], np.zeros(E), omega, m["electrodes"], c, N)
    assert Jo.shape == Jh.shape
    assert np.abs(Jh - Jo).max() < RAW_TOL
    assert np.linalg.norm(Jh - Jo) / np.linalg.norm(Jo - 1) < SIGNAL_TOL
    p.close()


def test_odd_number_of_electrode_nodes_takes_the_cp_async_staging_branch(torch_cuda):
    """woodbury_lr_kernel stages the Yh rows with TMA bulk copies when nU is even (16-byte aligned rows) and with
    8-byte cp.async otherwise; every shipped / synthetic layout has an even nU, so this mesh drops one electrode
    element to make nU odd."""
    t = torch_cuda
    n = 33
    cfg = meshgen.synthetic_config(n, 5)
    nodes, elem = meshgen.structured_mesh(n)
    param = o.elem_param(nodes, elem)
    els = o.electrode_elements(param, np.array(cfg["electrode_centers"]) / 1000, cfg["electrode_radius"] / 1000)
    els = [np.asarray(e) for e in els]
    p = None
    for drop in range(len(els[3])):                        # find an element whose removal removes exactly one node
        trial = list(els)
        trial[3] = np.delete(els[3], drop)
        nU = len(set(np.concatenate([elem[e].reshape(-1) for e in trial]).tolist()))
        if nU % 2 == 1:
            els = trial
            break
    else:
        pytest.fail("could not build an odd electrode-node count")
    L, E = len(els), len(elem)
    omega = 2 * np.pi * cfg["signal_frequency"]
    p = _lib.Plan(nodes, elem, els)
    assert p.nU % 2 == 1
    p.assemble(t.ones(E, dtype=t.float64, device="cuda"), t.zeros(E, dtype=t.float64, device="cuda"), omega)
    p.factor(False)
    J = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
    p.jacobian_block(0, L, 0, E, 1e-3, omega, J, E)
    assert p.numeric_flag() == 0
    rng = np.random.default_rng(11)
    pairs = [(int(rng.integers(L)), int(rng.integers(E))) for _ in range(10)] + [(3, int(els[3][0])), (2, int(els[3][1]))]
    cols = o.jacobian_columns_bruteforce(elem, param, np.ones(E), np.zeros(E), omega, els, pairs, 1e-3, len(nodes))
    Jh = J.cpu().numpy()
    for k, (i, j) in enumerate(pairs):
        assert np.abs(Jh[_rows(i, L), j] - cols[k]).max() < RAW_TOL
        sig = np.abs(cols[k] - 1).max()
        if sig > 1e-10:
            assert np.abs(Jh[_rows(i, L), j] - cols[k]).max() < 1e-3 * sig
    p.close()


def test_inverse_model_400x400_dual_splitk_against_numpy(torch_cuda):
    """BASELINE config C4 shape: m = 992, n_det = 257 762 (dual form, split-K Gram matrix).  A seeded J of that shape
    with the signal level of the real one; a sample of detection rows of inv_mat against numpy on the same J."""
    t = torch_cuda
    m_rows, E = 992, 318402
    g = t.Generator(device="cuda").manual_seed(4)
    J = 1.0 - 1e-8 * t.rand((m_rows, E), dtype=t.float64, device="cuda", generator=g)
    det_h = np.sort(np.random.default_rng(4).choice(E, size=257762, replace=False)).astype(np.int64)
    det = t.tensor(det_h, device="cuda")
    for lam in (289.0, 3e-6):
        flag = t.zeros(1, dtype=t.int32, device="cuda")
        inv = _lib.inverse_model(J, det, lam, flag=flag)
        assert int(flag.item()) == 0
        Jt = J.index_select(1, det) - 1.0
        G = (Jt @ Jt.T + lam * lam * t.eye(m_rows, dtype=t.float64, device="cuda")).cpu().numpy()
        Gi = np.linalg.inv(G)
        rows = np.random.default_rng(5).choice(len(det_h), size=400, replace=False)
        ref = Jt[:, t.tensor(rows, device="cuda")].cpu().numpy().T @ Gi
        got = inv[t.tensor(rows, device="cuda")].cpu().numpy()
        assert np.abs(got - ref).max() <= (MODEL_TOL if lam > 1 else 1e-6) * np.abs(ref).max()
        del inv, Jt


def test_singular_normal_matrix_raises_like_numpy(torch_cuda, workspace, golden):
    """np.linalg.inv raises LinAlgError('Singular matrix') on a singular normal matrix (Solver.py:121); the device
    path reports the non-positive Cholesky pivot through the flag and the classes raise before caching inv_mat.npy"""
    t = torch_cuda
    path = workspace("21_21")
    m, r = golden("21_21")
    folder = os.path.join(os.path.dirname(path), "MESH_21_21")
    J = t.ones((240, 882), dtype=t.float64, device="cuda")            # J - 1 == 0: Gram matrix 0, lambda = 0
    det = t.tensor(m["detection_index"], device="cuda")
    flag = t.zeros(1, dtype=t.int32, device="cuda")
    _lib.inverse_model(J, det, 0.0, flag=flag)
    assert int(flag.item()) == 1
    with pytest.raises(np.linalg.LinAlgError, match="Singular matrix"):
        _lib.check_flag(flag)
    _lib.inverse_model(J, det, 1.0, flag=flag)                        # the flag is reset by every call
    assert int(flag.item()) == 0
    Jn = np.ones((240, 882))
    Jn[0, 0] = np.nan
    np.save(os.path.join(folder, "JAC_cache.npy"), Jn)
    from ceit_b200.Solver import Solver
    with pytest.raises(np.linalg.LinAlgError):
        Solver(289)
    assert not os.path.exists(os.path.join(folder, "inv_mat.npy"))    # nothing cached


def test_real_complex_real_on_one_engine_with_repeats(torch_cuda, workspace, golden):
    """alternating material type on ONE engine with repeated (graph-replayed) calls.  The library
    re-allocates its workspaces on a real <-> complex switch; graphs captured before hold stale pointers and must be
    dropped (allocation generation), and nothing may be allocated inside a capture."""
    workspace("21_21")
    m, r = golden("21_21")
    from ceit_b200.EFEM import EFEM
    from ceit_b200.EJAC import EJAC
    fwd = EFEM()
    eng = fwd._engine
    for rep in range(3):
        fwd.reset_variable(0)
        for _ in range(3):                                            # eager, captured, replayed
            node_u, _, ep = fwd.calculation(7)
            assert np.abs(node_u - r["base_node_potential"][7]).max() < RAW_TOL
            assert np.abs(ep - r["base_electrode_potential"][7]).max() < RAW_TOL
        fwd.change_add_variable_geometry([-20 / 1000, -10 / 1000], 10 / 1000, 1e-7, "square")
        for _ in range(3):
            node_u, elem_u, ep = fwd.calculation(2)
            assert np.abs(node_u - r["obj_node_potential"]).max() < RAW_TOL
            assert np.abs(ep - r["obj_electrode_potential"]).max() < RAW_TOL
    gen = eng.plan.generation()
    assert gen >= 4                                                   # the workspaces were re-allocated on the switches
    jac = EJAC(fwd.mesh)                                              # same mesh -> same engine
    assert jac._eng is eng
    for _ in range(3):
        jac.fwd_FEM.invalidate()
        J = jac.JAC_calculation()
        assert np.abs(J - r["JAC"]).max() < RAW_TOL
    fwd.change_add_variable_geometry([-20 / 1000, -10 / 1000], 10 / 1000, 1e-7, "square")
    node_u, _, _ = fwd.calculation(2)                                 # complex again after the real Jacobian graphs
    assert np.abs(node_u - r["obj_node_potential"]).max() < RAW_TOL
    for _ in range(2):
        jac.fwd_FEM.invalidate()
        assert np.abs(jac.JAC_calculation() - r["JAC"]).max() < RAW_TOL
    # a captured sequence that would need a larger workspace is refused, not silently re-allocated under capture
    t = torch_cuda
    E, L = len(m["elem"]), 16
    p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"])
    perm = t.tensor(m["elem_perm"], device="cuda")
    var = t.zeros(E, dtype=t.float64, device="cuda")
    Jd = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
    p.assemble(perm, var, 2 * np.pi * 2e4)
    p.factor(False)
    p.jacobian_block(0, 1, 0, E, 1e-3, 2 * np.pi * 2e4, Jd, E)      # workspace for ONE excitation
    with pytest.raises(Exception):
        _lib.StepGraph(lambda: p.jacobian_block(0, L, 0, E, 1e-3, 2 * np.pi * 2e4, Jd, E))   # needs 16: refused
    t.cuda.synchronize()
    p.jacobian_block(0, L, 0, E, 1e-3, 2 * np.pi * 2e4, Jd, E)      # eager: allocates, correct
    assert np.abs(Jd.cpu().numpy() - r["JAC"]).max() < RAW_TOL
    p.close()


def test_every_one_shot_gauss_newton_variant_numerically(torch_cuda, workspace, golden):
    """EJAC.eit_solve_direct, the 4- and 8-electrode wrappers in both modes and eit_solve_on_some_electrodes
    (mode='direct') against numpy restatements of CEIT/EJAC.py:184-229,252-259 on the reference's own J."""
    path = workspace("21_21", is_first_JAC_calculation=False)
    m, r = golden("21_21")
    folder = os.path.join(os.path.dirname(path), "MESH_21_21")
    np.save(os.path.join(folder, "JAC_cache.npy"), r["JAC"])
    from ceit_b200.EJAC import EJAC
    jac = EJAC()
    J = jac.JAC_calculation()
    assert np.array_equal(J, r["JAC"])
    det = m["detection_index"]
    L = 16
    rng = np.random.default_rng(8)
    V = 1.0 + 1e-3 * rng.random(240)
    base = jac.electrode_original_potential

    def ref_subset(electrodes, potential, lmbda, mode):
        sl = []
        for i in electrodes:
            for j in electrodes:
                if j < i:
                    sl.append(i * (L - 1) + j)
                if j > i:
                    sl.append(i * (L - 1) + j - 1)
        dV = (potential - base)[sl] if mode == "direct" else potential
        Jt = J[:, det][sl] - 1
        return np.linalg.inv(Jt.T @ Jt + lmbda ** 2 * np.eye(len(det))) @ Jt.T @ dV * 1e-4

    cases = [
        (jac.eit_solve_4electrodes(V), ref_subset([2, 6, 10, 14], V, 90, "direct")),
        (jac.eit_solve_4electrodes_delta_V(np.arange(12.0) * 1e-3), ref_subset([2, 6, 10, 14], np.arange(12.0) * 1e-3, 90, "diff")),
        (jac.eit_solve_8electrodes(V), ref_subset([0, 2, 4, 6, 8, 10, 12, 14], V, 265, "direct")),
        (jac.eit_solve_8electrodes_delta_V(np.arange(56.0) * 1e-3),
         ref_subset([0, 2, 4, 6, 8, 10, 12, 14], np.arange(56.0) * 1e-3, 265, "diff")),
        (jac.eit_solve_on_some_electrodes([1, 5, 9], V, 120, mode="direct"), ref_subset([1, 5, 9], V, 120, "direct")),
    ]
    for got, ref in cases:
        assert got.shape == ref.shape == (len(det),)
        assert np.abs(got - ref).max() <= MODEL_TOL * np.abs(ref).max()
    with pytest.raises(AssertionError, match="Please enter the right solver mode"):
        jac.eit_solve_on_some_electrodes([1, 5], V, 10, mode="nope")
    # eit_solve_direct: reads inv_mat.npy (EJAC.py:252-259)
    jac.save_inv_matrix(203)
    inv = o.inverse_model(J, det, 203)
    ref = inv @ (V - base)
    got = jac.eit_solve_direct(V)
    assert np.abs(got - ref).max() <= MODEL_TOL * np.abs(ref).max()
    # a Jacobian assigned by the caller after a calculation is picked up (no stale device copy)
    jac.JAC_matrix = np.array(J) * 1.0
    jac.JAC_matrix[:, det[5]] += 1e-9
    got2 = jac.eit_solve(V, 295)
    ref2 = o.reconstruct(o.inverse_model(jac.JAC_matrix, det, 295), V - base)
    assert np.abs(got2 - ref2).max() <= MODEL_TOL * np.abs(ref2).max()


@pytest.mark.parametrize("tile", [0, 1, 2, 3])
def test_dgemm_vector_staging_edges(torch_cuda, tile):
    """16-byte cp.async staging (even leading dimensions) with ODD logical sizes: the last pair of every row is half
    valid and must be zero-filled, in all four transpose cases; and 8-byte-misaligned bases fall back to 8-byte copies."""
    import ctypes
    t = torch_cuda
    g = t.Generator(device="cuda").manual_seed(7)
    L = _lib.lib()
    _lib.set_option("gemm_tile", tile)
    try:
        for (M, N, K) in ((65, 33, 47), (129, 71, 17), (31, 130, 99)):
            for ta in (0, 1):
                for tb in (0, 1):
                    for off in (0, 1):                                # off = 1: base pointer only 8-byte aligned
                        Ab = t.randn((K if ta else M, (M if ta else K) + 3 + off), dtype=t.float64, device="cuda", generator=g)
                        Bb = t.randn((N if tb else K, (K if tb else N) + 5 + off), dtype=t.float64, device="cuda", generator=g)
                        if Ab.stride(0) % 2:
                            Ab = t.cat([Ab, Ab[:, :1]], 1).contiguous()
                        if Bb.stride(0) % 2:
                            Bb = t.cat([Bb, Bb[:, :1]], 1).contiguous()
                        A = Ab[:, off:off + (M if ta else K)]
                        B = Bb[:, off:off + (K if tb else N)]
                        C = t.zeros((M, N + 1), dtype=t.float64, device="cuda")
                        rc = L.ceit_dgemm(ta, tb, M, N, K, 1.0, ctypes.c_void_p(A.data_ptr()), A.stride(0),
                                          ctypes.c_void_p(B.data_ptr()), B.stride(0), 0.0, ctypes.c_void_p(C.data_ptr()),
                                          C.stride(0), _lib.stream_ptr())
                        assert rc == 0
                        ref = (A.T if ta else A) @ (B.T if tb else B)
                        assert ((C[:, :N] - ref).abs().max() / ref.abs().max()).item() < 1e-13
                        assert C[:, N].abs().max().item() == 0.0
    finally:
        _lib.set_option("gemm_tile", 0)


@pytest.mark.parametrize("tag", ["21_21", "50_50"])
def test_front_step_variants_agree(torch_cuda, golden, tag):
    """The schedules of the multi-level base stage - 4-way dissection with the destination-centric pull (default), the
    fused front kernel on small fronts, children pushing their update matrices, plain bisection with pull / push -
    give the same Jacobian, and each matches the reference."""
    t = torch_cuda
    m, r = golden(tag)
    E, L = len(m["elem"]), 16
    omega = 2 * np.pi * float(r["signal_frequency"])
    perm = t.tensor(m["elem_perm"], device="cuda")
    var = t.zeros(E, dtype=t.float64, device="cuda")
    outs = []
    defaults = {"nd_pull": 1, "nd_ways": 4, "front_kernel": 0}
    for opts in ({}, {"front_kernel": 1}, {"nd_pull": 0}, {"nd_ways": 2}, {"nd_ways": 2, "nd_pull": 0}):
        for k, v in opts.items():
            _lib.set_option(k, v)
        try:
            p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"])
            assert p.tree_levels >= 3
            p.assemble(perm, var, omega)
            p.factor(False)
            J = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
            p.jacobian_block(0, L, 0, E, 1e-3, omega, J, E)
            assert p.numeric_flag() == 0
            outs.append(J.cpu().numpy())
            p.close()
        finally:
            for k in opts:
                _lib.set_option(k, defaults[k])
    for Jh in outs:
        for k, (i, j) in enumerate(r["sample_pairs"]):
            assert np.abs(Jh[_rows(i, L), j] - r["sample_cols"][k]).max() < RAW_TOL
        if "JAC" in r:
            assert np.abs(Jh - r["JAC"]).max() < RAW_TOL
    assert all(np.abs(outs[0] - o_).max() < 1e-12 for o_ in outs[1:])


def test_shared_inverse_product_by_second_sweep(torch_cuda, golden):
    """Y_T = Xh T as a second backward sweep over the elimination forest (the default on large meshes) against the
    N x nU x nU product (the default on small ones): same Jacobian, and the reference's."""
    t = torch_cuda
    m, r = golden("50_50")
    E, L = len(m["elem"]), 16
    omega = 2 * np.pi * float(r["signal_frequency"])
    perm = t.tensor(m["elem_perm"], device="cuda")
    var = t.zeros(E, dtype=t.float64, device="cuda")
    outs = []
    for sweep in (0, 1):
        _lib.set_option("yt_sweep", sweep)
        try:
            p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"])
            for _ in range(2):                                   # twice: buffers are reused across steps
                p.assemble(perm, var, omega)
                p.factor(False)
                J = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
                p.jacobian_block(0, L, 0, E, 1e-3, omega, J, E)
            assert p.numeric_flag() == 0
            outs.append(J.cpu().numpy())
            p.close()
        finally:
            _lib.set_option("yt_sweep", -1)
    for Jh in outs:
        for k, (i, j) in enumerate(r["sample_pairs"]):
            assert np.abs(Jh[_rows(i, L), j] - r["sample_cols"][k]).max() < RAW_TOL
    assert np.abs(outs[0] - outs[1]).max() < 1e-12


def test_mesh_model_on_device_matches_reference_and_host_model(torch_cuda, workspace, golden, tmp_path, monkeypatch):
    """MeshObj's elem_param / electrode element sets / detection set from ONE kernel (ceit_mesh_model):
    index sets bit-exact against the reference's MeshObj (goldens) on the shipped meshes and against the host
    expressions on the 400x400 / 32-electrode mesh; elem_param <= 1e-12 relative."""
    from ceit_b200.models.mesh import MeshObj
    for tag in ("21_21", "50_50"):
        workspace(tag)
        m, _ = golden(tag)
        mesh = MeshObj()
        assert mesh.on_device
        scale = np.abs(m["elem_param"]).max(axis=0)
        assert (np.abs(mesh.elem_param - m["elem_param"]) / scale).max() < 1e-12
        for i, e in enumerate(m["electrodes"]):
            assert mesh.electrode_mesh[i] == [int(x) for x in e]
        assert np.array_equal(mesh.detection_index, m["detection_index"])
        assert np.array_equal(mesh.detection_elem, m["detection_elem"])
    path = meshgen.make_synthetic_workspace(str(tmp_path / "syn400"), 400, 9)
    monkeypatch.setenv("CEIT_B200_CONFIG", path)
    dev_mesh = MeshObj()
    assert dev_mesh.on_device and dev_mesh.electrode_num == 32
    monkeypatch.setattr(_lib, "cuda_available", lambda: False)
    host_mesh = MeshObj()
    assert not host_mesh.on_device
    assert np.array_equal(dev_mesh.detection_index, host_mesh.detection_index)
    assert all(dev_mesh.electrode_mesh[i] == host_mesh.electrode_mesh[i] for i in range(32))
    scale = np.abs(host_mesh.elem_param).max(axis=0)
    assert (np.abs(dev_mesh.elem_param - host_mesh.elem_param) / scale).max() < 1e-12
    # an electrode square that contains no centroid raises the reference's exception text
    monkeypatch.setattr(_lib, "cuda_available", lambda: True)
    with pytest.raises(Exception, match="No element is selected"):
        MeshObj(dev_mesh.mesh_obj, 1, [[10.0, 10.0]], 1e-4)


@pytest.mark.parametrize("tag", ["21_21", "50_50"])
def test_element_parallel_assembly_equals_gather_form_and_reference(torch_cuda, golden, tag):
    """K1 (CEIT/EFEM.py:47-69): the element-parallel form (default: geometry once per element, slot-sorted contribution
    list, segmented sum) against the slot-gather form and the reference's K, real and complex material."""
    t = torch_cuda
    m, r = golden(tag)
    E = len(m["elem"])
    omega = 2 * np.pi * float(r["signal_frequency"])
    p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"])
    perm = t.tensor(m["elem_perm"], device="cuda")
    rng = np.random.default_rng(5)
    for var_h in (np.zeros(E), rng.random(E) * 1e-7):
        var = t.tensor(var_h, device="cuda")
        out = []
        for gather in (0, 1):
            _lib.set_option("assemble_gather", gather)
            try:
                vals = t.zeros((p.nnz, 2), dtype=t.float64, device="cuda")
                p.assemble(perm, var, omega, vals)
            finally:
                _lib.set_option("assemble_gather", 0)
            out.append(vals.cpu().numpy())
        scale = np.abs(out[1]).max(axis=0) + 1e-300
        assert (np.abs(out[0] - out[1]) / scale).max() < 4e-16
        if not var_h.any():
            v = out[0]
            assert np.abs(v[:, 0] + 1j * v[:, 1] - r["K_data"]).max() <= 1e-12 * np.abs(r["K_data"]).max()
    # symmetric bit for bit: K[a, b] == K[b, a]
    tb = p.tables()
    rowptr, colidx = tb["rowptr"], tb["colidx"]
    import scipy.sparse as sp
    K = sp.csr_matrix((out[0][:, 0] + 1j * out[0][:, 1], colidx, rowptr), shape=(len(m["nodes"]),) * 2)
    assert abs(K - K.T).max() == 0


@pytest.mark.parametrize("tag,parts", [("50_50", 2), ("50_50", 4), ("21_21", 4), ("50_50", 3), ("50_50", 8)])
def test_domain_decomposed_step_equals_single_plan(torch_cuda, golden, tag, parts):
    """The multi-GPU base stage (ceit_factor_phase) with its ranks run one after the other on ONE device: the
    exchange between the two halves is done here with plain copies / sums instead of NCCL.  Every rank's Jacobian
    columns (its own elements, all excitations) and its rows of the inverse model against the undecomposed plan and
    the reference's J."""
    from ceit_b200.sharding import _DevMem
    t = torch_cuda
    m, r = golden(tag)
    N, E, L = len(m["nodes"]), len(m["elem"]), 16
    omega = 2 * np.pi * float(r["signal_frequency"])
    c = float(r["variable_change_for_JAC"])
    perm = t.tensor(m["elem_perm"], device="cuda")
    var = t.zeros(E, dtype=t.float64, device="cuda")
    det = t.tensor(m["detection_index"], device="cuda")
    # the undecomposed plan
    p0 = _lib.Plan(m["nodes"], m["elem"], m["electrodes"])
    J0 = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
    b0 = t.zeros(L * (L - 1), dtype=t.float64, device="cuda")
    p0.assemble(perm, var, omega)
    p0.factor(False)
    p0.jacobian_block(0, L, 0, E, c, omega, J0, E, b0)
    inv0 = _lib.inverse_model(J0, det, 289.0, shift=1.0, form=2)
    with pytest.raises(_lib.CeitError):
        _lib.Plan(m["nodes"], m["elem"], m["electrodes"], parts=parts, part=0).factor(False)
    plans = [_lib.Plan(m["nodes"], m["elem"], m["electrodes"], parts=parts, part=q) for q in range(parts)]
    own = np.stack([p.own_elements() for p in plans])
    assert (own.sum(axis=0) == 1).all()                       # the elements are partitioned
    for p in plans:
        p.assemble(perm, var, omega)
        p.factor_phase(1)
    t.cuda.synchronize()
    bufs = []
    for p in plans:
        info, g, rr = p.exchange_info()
        assert info[0] == parts and info[4] < N               # a strict subset of the node rows per rank
        per, nred = int(info[2]), int(info[3])
        bufs.append((t.as_tensor(_DevMem(g, per * parts), device="cuda") if per else None,
                     t.as_tensor(_DevMem(rr, nred), device="cuda")))
    total = sum(b[1] for b in bufs)
    for q, (xg, red) in enumerate(bufs):
        red.copy_(total)
        if xg is not None:
            for s in range(parts):
                if s != q:
                    xg[s * per:(s + 1) * per] = bufs[s][0][s * per:(s + 1) * per]
    J = t.zeros_like(J0)
    for q, p in enumerate(plans):
        p.factor_phase(2)
        Jq = t.zeros_like(J0)
        bq = t.zeros_like(b0)
        p.jacobian_block(0, L, 0, E, c, omega, Jq, E, bq)
        assert p.numeric_flag() == 0
        mask = t.from_numpy(own[q]).to("cuda")
        assert float(Jq[:, ~mask].abs().max()) == 0.0         # nothing outside the rank's own columns
        assert float((bq - b0).abs().max()) < 1e-12
        J[:, mask] = Jq[:, mask]
        # this rank's rows of the inverse model from its own detection columns (Gram matrices summed by hand below)
    assert float((J - J0).abs().max()) < 1e-12
    assert float(t.linalg.norm(J - J0) / t.linalg.norm(J0 - 1)) < 1e-5     # a different summation order, not bit-equal
    if "JAC" in r:
        Jh = J.cpu().numpy()
        assert np.abs(Jh - r["JAC"]).max() < RAW_TOL
        assert np.linalg.norm(Jh - r["JAC"]) / np.linalg.norm(r["JAC"] - 1) < SIGNAL_TOL
    # inverse model sharded by the same element ownership: partial Gram matrices of the own detection columns
    mrows = L * (L - 1)
    G = t.zeros((mrows, mrows), dtype=t.float64, device="cuda")
    dets = []
    for q in range(parts):
        mask = t.from_numpy(own[q]).to("cuda")
        sel = mask[det]
        dets.append((t.nonzero(sel).flatten(), det[sel].contiguous()))
        Gq = t.empty_like(G)
        _lib.check(_lib.lib().ceit_gram_partial(_lib.tptr(J), E, mrows, _lib.tptr(dets[q][1]), dets[q][1].numel(),
                                                _lib.tptr(None), mrows, 1.0, _lib.tptr(Gq), _lib.stream_ptr()))
        G += Gq
    inv = t.zeros_like(inv0)
    for q in range(parts):
        posq, dq = dets[q]
        out = t.empty((dq.numel(), mrows), dtype=t.float64, device="cuda")
        _lib.check(_lib.lib().ceit_inverse_from_gram(_lib.tptr(J), E, mrows, _lib.tptr(dq), dq.numel(), _lib.tptr(None),
                                                     mrows, 1.0, 289.0, 1.0, _lib.tptr(G), _lib.tptr(out), _lib.tptr(None),
                                                     _lib.stream_ptr()))
        inv[posq] = out
    # given the same J the sharded model equals the undecomposed one; against the model of J0 only the rounding of
    # the signal J - 1 (1e-6 relative, see above) shows
    inv_same = _lib.inverse_model(J, det, 289.0, shift=1.0, form=2)
    assert float((inv - inv_same).abs().max() / inv_same.abs().max()) < MODEL_TOL
    assert float((inv - inv0).abs().max() / inv0.abs().max()) < 1e-4


def test_persistent_woodbury_kernel_equals_per_cta_kernel(torch_cuda, golden):
    """K2: the persistent ring kernel (wb_persist = 1, default) against the one-CTA-per-(patch, excitation) kernel on
    the same plan (same arithmetic per column; the 3x3 corrections are summed in a different order, far below one ulp
    of J), incl. a partial excitation / element range."""
    t = torch_cuda
    for tag in ("21_21", "50_50"):
        m, r = golden(tag)
        E, L = len(m["elem"]), 16
        omega = 2 * np.pi * float(r["signal_frequency"])
        c = float(r["variable_change_for_JAC"])
        perm = t.tensor(m["elem_perm"], device="cuda")
        var = t.zeros(E, dtype=t.float64, device="cuda")
        for mode in (1,):
            _lib.set_option("wb_persist", mode)
            try:
                p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"])
                p.assemble(perm, var, omega)
                p.factor(False)
                outs = []
                for persist in (mode, 0):
                    _lib.set_option("wb_persist", persist)
                    l0 = _lib.launch_count()
                    J = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
                    p.jacobian_block(0, L, 0, E, c, omega, J, E)
                    J2 = t.zeros_like(J)
                    p.jacobian_block(3, 7, E // 5, E // 2, c, omega, J2, E)
                    outs.append((J, J2))
            finally:
                _lib.set_option("wb_persist", 1)
            assert float((outs[0][0] - outs[1][0]).abs().max()) <= 4.5e-16
            assert float((outs[0][1] - outs[1][1]).abs().max()) <= 4.5e-16
            assert float((outs[0][0] - 1).abs().max()) > 1e-9                     # the kernels did write the signal
            assert float(outs[0][1][:3 * (L - 1)].abs().max()) == 0.0 and float(outs[0][1][:, :E // 5].abs().max()) == 0.0
            if "JAC" in r:
                assert np.abs(outs[0][0].cpu().numpy() - r["JAC"]).max() < RAW_TOL
            p.close()
