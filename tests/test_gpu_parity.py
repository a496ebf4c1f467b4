"""Parity of the CUDA path (through the C ABI and the drop-in classes) with the reference goldens
and the CPU oracle.  Needs a B200: run with `pytest -m gpu`.

Tolerances: raw J entries (which are ~1) <= 1e-9 relative - measured ~1e-13;
the second-order signal J-1 (~1e-8..1e-7) <= 1e-4 relative in Frobenius norm against the reference,
whose own np.linalg.inv noise (~1e-13 absolute) is the floor; inverse models and maps <= 1e-9
relative given the same J; CSR pattern and every index bit-exact.
"""
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


# ---------------------------------------------------------------------------------------------------
# C ABI level
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tile", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(64, 64, 16), (1, 1, 1), (100, 37, 53), (273, 264, 202), (515, 129, 70)])
def test_dgemm_matches_torch_fp64(torch_cuda, shape, tile):
    """every CTA-tile configuration of the DMMA kernel (0 = automatic choice) x every transpose case"""
    t = torch_cuda
    M, N, K = shape
    g = t.Generator(device="cuda").manual_seed(1)
    _lib.set_option("gemm_tile", tile)
    try:
        _dgemm_cases(t, M, N, K, g)
    finally:
        _lib.set_option("gemm_tile", 0)


def _dgemm_cases(t, M, N, K, g):
    for ta in (False, True):
        for tb in (False, True):
            A = t.randn((K, M) if ta else (M, K), dtype=t.float64, device="cuda", generator=g)
            B = t.randn((N, K) if tb else (K, N), dtype=t.float64, device="cuda", generator=g)
            C0 = t.randn((M, N), dtype=t.float64, device="cuda", generator=g)
            C = C0.clone()
            _lib.dgemm(A, B, ta, tb, alpha=0.7, beta=-0.3, C=C)
            ref = 0.7 * ((A.T if ta else A) @ (B.T if tb else B)) - 0.3 * C0
            assert ((C - ref).abs().max() / ref.abs().max()).item() < 1e-13


@pytest.mark.parametrize("tile", [0, 1, 2])
@pytest.mark.parametrize("shape", [(64, 64, 16), (1, 1, 1), (100, 37, 53), (273, 264, 202)])
def test_zgemm_matches_torch_complex128(torch_cuda, shape, tile):
    """complex DMMA kernel (split re/im planes, four DMMA chains) against torch.matmul in complex128"""
    t = torch_cuda
    M, N, K = shape
    g = t.Generator(device="cuda").manual_seed(2)

    def rnd(*s):
        return t.complex(t.randn(*s, dtype=t.float64, device="cuda", generator=g),
                         t.randn(*s, dtype=t.float64, device="cuda", generator=g))
    _lib.set_option("gemm_tile", tile)
    try:
        for ta in (False, True):
            for tb in (False, True):
                A = rnd(*((K, M) if ta else (M, K)))
                B = rnd(*((N, K) if tb else (K, N)))
                C0 = rnd(M, N)
                C = C0.clone()
                _lib.zgemm(A, B, ta, tb, alpha=0.7 - 0.2j, beta=-0.3 + 0.1j, C=C)
                ref = (0.7 - 0.2j) * ((A.T if ta else A) @ (B.T if tb else B)) + (-0.3 + 0.1j) * C0
                assert ((C - ref).abs().max() / ref.abs().max()).item() < 1e-13
    finally:
        _lib.set_option("gemm_tile", 0)


@pytest.mark.parametrize("target,nd", [(100, 40), (64, 0), (0, 64), (0, 16)])
def test_step_is_bitwise_reproducible(torch_cuda, golden, target, nd):
    """The path has no atomics, so repeated steps must agree bit for bit; a difference is a race between the
    streams of the base stage (the two arms of the block chain once accumulated into the middle block from two
    streams: 1 step in 10 was wrong for chains of >= 3 blocks).  tools/race_stress.py is the long version."""
    t = torch_cuda
    m, r = golden("50_50")
    E, L = len(m["elem"]), 16
    omega = 2 * np.pi * float(r["signal_frequency"])
    _lib.set_option("nd_interior", nd)
    try:
        p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"], target_block=target)
    finally:
        _lib.set_option("nd_interior", 64)
    perm = t.tensor(m["elem_perm"], device="cuda")
    var = t.zeros(E, dtype=t.float64, device="cuda")
    det = t.tensor(m["detection_index"], device="cuda")
    J = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
    first = None
    for _ in range(60):
        J.zero_()
        p.assemble(perm, var, omega)
        p.factor(False)
        p.jacobian_block(0, L, 0, E, 1e-3, omega, J, E)
        inv = _lib.inverse_model(J, det, 289)
        t.cuda.synchronize()
        if first is None:
            first = (J.clone(), inv.clone())
        else:
            assert t.equal(J, first[0]) and t.equal(inv, first[1])
    p.close()


@pytest.mark.parametrize("tag", ["21_21", "50_50"])
@pytest.mark.parametrize("target,nd", [(0, 64), (64, 0), (10 ** 6, 0), (0, 20), (100, 40), (0, 9), (192, 64)])
def test_cabi_pipeline_against_reference(torch_cuda, golden, tag, target, nd):
    """every ordering variant: nested-dissection interiors on/off x block-tridiagonal block size
    (target 10**6 with nd 0 = one dense block)"""
    t = torch_cuda
    m, r = golden(tag)
    N, E, L = len(m["nodes"]), len(m["elem"]), 16
    omega = 2 * np.pi * float(r["signal_frequency"])
    c = float(r["variable_change_for_JAC"])
    _lib.set_option("nd_interior", nd)
    try:
        p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"], target_block=target)
    finally:
        _lib.set_option("nd_interior", 64)
    assert (p.P > 0 or p.tree_levels > 0) == (nd > 0)
    assert (p.tree_levels > 0) == (target == 0 and nd > 0)     # target 0 = default ordering: multi-level dissection
    perm = t.tensor(m["elem_perm"], device="cuda")
    var = t.zeros(E, dtype=t.float64, device="cuda")
    vals = t.zeros((p.nnz, 2), dtype=t.float64, device="cuda")
    ep = t.zeros((E, 9), dtype=t.float64, device="cuda")
    p.assemble(perm, var, omega, vals, ep)
    v = vals.cpu().numpy()
    assert np.abs(v[:, 0] + 1j * v[:, 1] - r["K_data"]).max() <= 1e-12 * np.abs(r["K_data"]).max()
    scale = np.abs(m["elem_param"]).max(axis=0)
    assert (np.abs(ep.cpu().numpy() - m["elem_param"]) / scale).max() < 1e-12
    p.factor(False)
    J = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
    base = t.zeros(L * (L - 1), dtype=t.float64, device="cuda")
    p.jacobian_block(0, L, 0, E, c, omega, J, E, base)
    assert p.numeric_flag() == 0
    Jh = J.cpu().numpy()
    assert np.abs(base.cpu().numpy() - 1).max() < 1e-11
    for k, (i, j) in enumerate(r["sample_pairs"]):
        assert np.abs(Jh[_rows(i, L), j] - r["sample_cols"][k]).max() < RAW_TOL
    if "JAC" in r:
        assert np.abs(Jh - r["JAC"]).max() < RAW_TOL
        assert np.linalg.norm(Jh - r["JAC"]) / np.linalg.norm(r["JAC"] - 1) < SIGNAL_TOL
        # columns of elements wholly inside the excitation electrode equal the baseline
        for i in range(L):
            for e in m["electrodes"][i]:
                assert np.abs(Jh[_rows(i, L), e] - base.cpu().numpy()[_rows(i, L)]).max() < 1e-15
    # forward solve
    na = t.zeros(N, dtype=t.float64, device="cuda")
    epot = t.zeros(E, dtype=t.float64, device="cuda")
    el = t.zeros(L, dtype=t.float64, device="cuda")
    p.forward(3, na, epot, el)
    assert np.abs(na.cpu().numpy() - r["base_node_potential"][3]).max() < RAW_TOL
    assert np.abs(el.cpu().numpy() - r["base_electrode_potential"][3].real).max() < RAW_TOL
    # complex-symmetric path: Example_02 object
    p.assemble(perm, t.tensor(r["obj_variable"], device="cuda"), omega)
    p.factor(True)
    p.forward(2, na, epot, el)
    assert p.numeric_flag() == 0
    assert np.abs(na.cpu().numpy() - r["obj_node_potential"]).max() < RAW_TOL
    assert np.abs(epot.cpu().numpy() - r["obj_element_potential"].real).max() < RAW_TOL
    assert np.abs(el.cpu().numpy() - r["obj_electrode_potential"].real).max() < RAW_TOL
    if "nz_sample_pairs" in r:
        p.assemble(perm, t.full((E,), float(r["nz_base_variable"]), dtype=t.float64, device="cuda"), omega)
        p.factor(True)
        p.jacobian_block(0, L, 0, E, c, omega, J, E, base)
        Jh = J.cpu().numpy()
        for k, (i, j) in enumerate(r["nz_sample_pairs"]):
            assert np.abs(Jh[_rows(i, L), j] - r["nz_sample_cols"][k]).max() < RAW_TOL
    p.close()


def test_jacobian_block_ranges_compose(torch_cuda, golden):
    """Excitation ranges (calc_from/calc_end) and element ranges tile the same matrix."""
    t = torch_cuda
    m, r = golden("21_21")
    E, L = len(m["elem"]), 16
    omega = 2 * np.pi * 2e4
    p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"])
    p.assemble(t.tensor(m["elem_perm"], device="cuda"), t.zeros(E, dtype=t.float64, device="cuda"), omega)
    p.factor(False)
    full = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
    p.jacobian_block(0, L, 0, E, 1e-3, omega, full, E)
    parts = t.zeros_like(full)
    for (i0, i1) in ((0, 5), (5, 6), (6, 16)):
        for (e0, e1) in ((0, 1), (1, 400), (400, E)):
            p.jacobian_block(i0, i1, e0, e1, 1e-3, omega, parts, E)
    assert t.equal(full, parts)
    with pytest.raises(_lib.CeitError):
        p.jacobian_block(0, L + 1, 0, E, 1e-3, omega, full, E)
    p.close()


def test_inverse_model_and_reconstruct(torch_cuda, golden):
    t = torch_cuda
    m, r = golden("21_21")
    J = t.tensor(r["JAC"], device="cuda")
    det = t.tensor(m["detection_index"], device="cuda")
    for form in (1, 2):
        for lam, key in ((289, "inv_mat_289"), (2e-5, "inv_mat_2e-5")):
            inv = _lib.inverse_model(J, det, lam, form=form).cpu().numpy()
            assert np.abs(inv - r[key]).max() <= MODEL_TOL * np.abs(r[key]).max()
    inv = _lib.inverse_model(J, det, 289)
    out = _lib.reconstruct(inv, t.tensor(r["solve_dV"], device="cuda")).cpu().numpy()
    assert np.abs(out - r["solve_out"]).max() <= MODEL_TOL * np.abs(r["solve_out"]).max()
    g = t.Generator(device="cuda").manual_seed(0)
    dV = t.rand((240, 777), dtype=t.float64, device="cuda", generator=g)
    outb = _lib.reconstruct(inv, dV)
    assert ((outb - inv @ dV).abs().max() / outb.abs().max()).item() < 1e-13
    # batched == frame by frame
    one = _lib.reconstruct(inv, dV[:, 5].contiguous())
    assert ((one - outb[:, 5]).abs().max() / one.abs().max()).item() < 1e-13
    # sharded over detection elements (multi-GPU form; one rank here): slices of the same matrix
    full = _lib.inverse_model(J, det, 289, form=2)
    for lo, hi in ((0, 300), (300, 722)):
        part = _lib.inverse_model_sharded(J, det[lo:hi].contiguous(), 289)
        # one rank: the Gram matrix only covers the slice, so compare against the model of that slice
        ref_part = _lib.inverse_model(J, det[lo:hi].contiguous(), 289, form=2)
        assert ((part - ref_part).abs().max() / ref_part.abs().max()).item() < 1e-12
    # row subset + scale (eit_solve_on_some_electrodes)
    rows = [2 * 15 + 5, 6 * 15 + 2, 10 * 15 + 13, 14 * 15 + 2]
    sub = _lib.inverse_model(J, det, 90, rows=t.tensor(rows, device="cuda"), scale=1e-4).cpu().numpy()
    Jt = r["JAC"][rows][:, m["detection_index"]] - 1
    ref = np.linalg.inv(Jt.T @ Jt + 90 ** 2 * np.eye(Jt.shape[1])) @ Jt.T * 1e-4
    assert np.abs(sub - ref).max() <= MODEL_TOL * np.abs(ref).max()


# ---------------------------------------------------------------------------------------------------
# drop-in classes (the reference's own call sequences: Example_02 / 03 / 04, test_Solver)
# ---------------------------------------------------------------------------------------------------
def test_example_02_efem(torch_cuda, workspace, golden):
    workspace("21_21")
    m, r = golden("21_21")
    from ceit_b200.EFEM import EFEM
    fwd = EFEM()
    fwd.change_add_variable_geometry([-20 / 1000, -10 / 1000], 10 / 1000, 1e-7, "square")
    assert np.array_equal(fwd.elem_variable, r["obj_variable"])
    node_u, elem_u, electrode_potential = fwd.calculation(2)
    assert node_u.dtype == np.float64 and elem_u.dtype == np.complex128
    assert np.abs(node_u - r["obj_node_potential"]).max() < RAW_TOL
    assert np.abs(elem_u - r["obj_element_potential"]).max() < RAW_TOL
    assert np.abs(electrode_potential - r["obj_electrode_potential"]).max() < RAW_TOL
    fwd.reset_variable(0)
    for i in (0, 7):
        node_u, _, ep = fwd.calculation(i)
        assert np.abs(node_u - r["base_node_potential"][i]).max() < RAW_TOL
        assert np.abs(ep - r["base_electrode_potential"][i]).max() < RAW_TOL
    rowptr, colidx, vals = fwd.csr_K()
    assert np.array_equal(rowptr, r["K_indptr"]) and np.array_equal(colidx, r["K_indices"])
    assert np.abs(vals - r["K_data"]).max() <= 1e-12 * np.abs(r["K_data"]).max()
    K = fwd.K_sparse
    assert K.shape == (len(m["nodes"]),) * 2 and np.abs(K - K.T).max() == 0
    with pytest.raises(IndexError):
        fwd.calculation(16)


def test_example_03_ejac_with_resume(torch_cuda, workspace, golden):
    path = workspace("21_21", calc_from=0, calc_end=9)
    m, r = golden("21_21")
    from ceit_b200.EJAC import EJAC
    jac = EJAC()
    assert np.abs(jac.electrode_original_potential - 1).max() < 1e-11
    J1 = jac.JAC_calculation().copy()
    assert np.all(J1[9 * 15:] == 0) and np.abs(J1[:9 * 15] - r["JAC"][:9 * 15]).max() < RAW_TOL
    cache = os.path.join(os.path.dirname(path), "MESH_21_21", "JAC_cache.npy")
    assert np.array_equal(np.load(cache), J1)
    # second machine / resumed run: calc_from=9 loads the cache first (EJAC.py:111-112)
    workspace("21_21", calc_from=9, calc_end=16)
    np.save(cache, J1)
    jac2 = EJAC()
    J = jac2.JAC_calculation()
    assert J.shape == (240, 882) and J.dtype == np.float64 and J.flags["C_CONTIGUOUS"]
    assert np.abs(J - r["JAC"]).max() < RAW_TOL
    assert np.linalg.norm(J - r["JAC"]) / np.linalg.norm(r["JAC"] - 1) < SIGNAL_TOL
    # is_first_JAC_calculation = false -> pure load
    workspace("21_21", is_first_JAC_calculation=False)
    np.save(cache, J)
    jac3 = EJAC()
    assert np.array_equal(jac3.JAC_calculation(), J)
    # one-shot Gauss-Newton variants against the oracle fed the same J
    dV = np.random.default_rng(3).random(240)
    det = m["detection_index"]
    ref = o.reconstruct(o.inverse_model(J, det, 295), dV - jac3.electrode_original_potential)
    got = jac3.eit_solve(dV, 295)
    assert np.abs(got - ref).max() <= MODEL_TOL * np.abs(ref).max()
    Jd = J[:, det]
    ref = np.linalg.inv(Jd.T @ Jd + 295 ** 2 * np.eye(len(det))) @ Jd.T @ dV       # no "-1" (EJAC.py:177)
    got = jac3.eit_solve_delta_V(dV, 295)
    assert np.abs(got - ref).max() <= 1e-7 * np.abs(ref).max()
    got = jac3.eit_solve_4electrodes_delta_V(np.arange(12.0))
    assert got.shape == (len(det),)
    jac3.save_inv_matrix(203)
    ref = o.inverse_model(J, det, 203)
    assert np.abs(jac3.read_inv_matrix() - ref).max() <= MODEL_TOL * np.abs(ref).max()


def test_write_behind_cache_files(torch_cuda, workspace, golden, monkeypatch):
    """opt-in write-behind (CEIT_B200_WRITE_BEHIND=1): the cache files are written by a worker thread; every reader
    of this package joins first, and after wait_npy() the files hold the same bytes as the synchronous default"""
    path = workspace("21_21")
    m, r = golden("21_21")
    folder = os.path.join(os.path.dirname(path), "MESH_21_21")
    from ceit_b200.EJAC import EJAC
    from ceit_b200.Solver import Solver
    from ceit_b200.util.utilities import wait_npy
    jac = EJAC()
    J_sync = jac.JAC_calculation().copy()
    jac.save_inv_matrix(289)
    files_sync = [open(os.path.join(folder, f), "rb").read() for f in ("JAC_cache.npy", "inv_mat.npy")]
    monkeypatch.setenv("CEIT_B200_WRITE_BEHIND", "1")
    for _ in range(5):                                          # repeated steps reuse the pinned buffers
        jac.fwd_FEM.invalidate()
        J = jac.JAC_calculation()
        jac.save_inv_matrix(289)
    assert np.array_equal(jac.read_inv_matrix(), np.load(os.path.join(folder, "inv_mat.npy")))   # joins before reading
    s = Solver()                                                # reads both caches through load_npy
    assert np.array_equal(s.JAC_mat, J_sync) and np.array_equal(J, J_sync)
    wait_npy()
    files_async = [open(os.path.join(folder, f), "rb").read() for f in ("JAC_cache.npy", "inv_mat.npy")]
    assert files_async[0] == files_sync[0]
    ref = o.inverse_model(J_sync, m["detection_index"], 289)     # the model of THIS Jacobian (Solver() rewrote the file)
    assert np.abs(np.load(os.path.join(folder, "inv_mat.npy")) - ref).max() <= MODEL_TOL * np.abs(ref).max()


def test_example_04_solver_and_reinitialize(torch_cuda, workspace, golden):
    path = workspace("21_21")
    m, r = golden("21_21")
    folder = os.path.join(os.path.dirname(path), "MESH_21_21")
    from ceit_b200.Solver import Solver, reinitialize_solver
    with pytest.raises(AssertionError, match="The JAC matrix is not generated"):
        Solver()
    np.save(os.path.join(folder, "JAC_cache.npy"), r["JAC"])
    s = Solver()                                                # default lmbda=289
    assert os.path.exists(os.path.join(folder, "inv_mat.npy"))
    assert s.inv_mat.shape == (722, 240)
    assert np.abs(s.inv_mat - r["inv_mat_289"]).max() <= MODEL_TOL * np.abs(r["inv_mat_289"]).max()
    out = s.solve(r["solve_dV"])
    assert out.shape == (722,) and np.abs(out - r["solve_out"]).max() <= MODEL_TOL * np.abs(r["solve_out"]).max()
    frames = np.random.default_rng(0).random((240, 64))
    outb = s.solve(frames)
    assert outb.shape == (722, 64)
    assert np.abs(outb - r["inv_mat_289"] @ frames).max() <= MODEL_TOL * np.abs(outb).max()
    s2 = Solver(1.0)                                            # cached file wins over lmbda (Solver.py:44-51)
    assert np.array_equal(s2.inv_mat, s.inv_mat)
    s3 = reinitialize_solver(2e-5)
    assert np.abs(s3.inv_mat - r["inv_mat_2e-5"]).max() <= MODEL_TOL * np.abs(r["inv_mat_2e-5"]).max()
    with pytest.raises(AssertionError, match="positive lmbda"):
        reinitialize_solver(0)
    np.save(os.path.join(folder, "inv_mat.npy"), np.zeros((5, 240)))
    with pytest.raises(AssertionError, match="dimension not equal"):
        Solver()


def test_simulator_and_resistance(torch_cuda, workspace, golden):
    workspace("21_21", resistance=2.0)
    m, r = golden("21_21")
    from ceit_b200.Simulator import Simulator
    sim = Simulator()
    v = sim.calc_potential()
    assert v.shape == (240,) and np.abs(v - 1).max() < 1e-11           # real base: phi == 1 for any resistance
    E, N = len(m["elem"]), len(m["nodes"])
    omega = 2 * np.pi * 2e4
    sim.fwd_simulator.change_add_variable_geometry([0.01, 0.0], 0.012, 2e-7, "square")
    v = sim.calc_potential()
    p = o.elem_param(m["nodes"], m["elem"])
    K = o.assemble_csr(m["elem"], p, m["elem_perm"] / 2.0, sim.fwd_simulator.elem_variable, omega, N)
    for i in (0, 9):
        _, _, el = o.forward_sparse(K, m["elem"], m["electrodes"], i)
        assert np.abs(v[_rows(i, 16)] - o.jac_rows(el, i, 16)).max() < RAW_TOL


# ---------------------------------------------------------------------------------------------------
# synthetic structured meshes: oracle at small size, size-independent properties at full size
# ---------------------------------------------------------------------------------------------------
def _synthetic(t, n, per_side, c=1e-3, target=0):
    cfg = meshgen.synthetic_config(n, per_side)
    nodes, elem = meshgen.structured_mesh(n)
    param = o.elem_param(nodes, elem)
    centers = np.array(cfg["electrode_centers"]) / 1000
    els = o.electrode_elements(param, centers, cfg["electrode_radius"] / 1000)
    L, E = len(els), len(elem)
    omega = 2 * np.pi * cfg["signal_frequency"]
    p = _lib.Plan(nodes, elem, els, target_block=target)
    p.assemble(t.ones(E, dtype=t.float64, device="cuda"), t.zeros(E, dtype=t.float64, device="cuda"), omega)
    p.factor(False)
    J = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
    base = t.zeros(L * (L - 1), dtype=t.float64, device="cuda")
    p.jacobian_block(0, L, 0, E, c, omega, J, E, base)
    assert p.numeric_flag() == 0
    return dict(nodes=nodes, elem=elem, param=param, els=els, L=L, E=E, omega=omega, plan=p, J=J, base=base)


def test_synthetic_40x40_against_bruteforce_oracle(torch_cuda):
    s = _synthetic(torch_cuda, 40, 5)
    L, E = s["L"], s["E"]
    rng = np.random.default_rng(5)
    pairs = [(int(rng.integers(L)), int(rng.integers(E))) for _ in range(12)]
    pairs += [(0, int(s["els"][0][0])), (3, int(s["els"][4][1]))]
    cols = o.jacobian_columns_bruteforce(s["elem"], s["param"], np.ones(E), np.zeros(E), s["omega"], s["els"], pairs,
                                         1e-3, len(s["nodes"]))
    Jh = s["J"].cpu().numpy()
    for k, (i, j) in enumerate(pairs):
        assert np.abs(Jh[_rows(i, L), j] - cols[k]).max() < RAW_TOL
        sig = np.abs(cols[k] - 1).max()
        if sig > 1e-10:
            assert np.abs(Jh[_rows(i, L), j] - cols[k]).max() < 1e-3 * sig
    s["plan"].close()


@pytest.mark.parametrize("n,per_side", [(200, 5), (400, 9)])
def test_large_mesh_properties(torch_cuda, n, per_side):
    """BASELINE configs C3/C4 at full size: properties that do not need an oracle."""
    t = torch_cuda
    s = _synthetic(t, n, per_side)
    L, E, J, base = s["L"], s["E"], s["J"], s["base"]
    assert (base - 1).abs().max().item() < 1e-9                        # real base: phi0 == 1
    d = J - 1
    assert d.max().item() < 1e-11 and d.min().item() > -1e-5           # second-order signal is negative
    # wholly-Dirichlet elements reproduce the baseline exactly
    for i in (0, L // 2):
        idx = t.tensor(np.asarray(s["els"][i]), device="cuda")
        blk = J[i * (L - 1):(i + 1) * (L - 1)][:, idx]
        assert (blk - base[i * (L - 1):(i + 1) * (L - 1), None]).abs().max().item() < 1e-15
    # sampled columns against the sparse brute-force oracle (a few solves of the full mesh)
    rng = np.random.default_rng(n)
    pairs = [(int(rng.integers(L)), int(rng.integers(E))) for _ in range(3)]
    cols = o.jacobian_columns_bruteforce(s["elem"], s["param"], np.ones(E), np.zeros(E), s["omega"], s["els"], pairs,
                                         1e-3, len(s["nodes"]))
    Jh = J[:, [j for _, j in pairs]].cpu().numpy()
    for k, (i, j) in enumerate(pairs):
        assert np.abs(Jh[_rows(i, L), k] - cols[k]).max() < RAW_TOL
    # mirror symmetry of the structured sheet: excitation 0 sits on the x axis (y -> -y maps the
    # electrode ring onto itself: electrode m <-> L-m); triangles are not mirror images (fixed
    # diagonal) so compare the element-pair sums of every cell row instead of single elements
    s["plan"].close()


@pytest.mark.parametrize("is_complex", [False, True])
def test_rank_update_excitation_equals_direct_factorisation(torch_cuda, golden, is_complex):
    """Per-excitation Schur stage: rank-(b+1) update of one shared inverse vs a factorisation of every
    padded S_U (the fallback path) - same J, real and complex-symmetric base."""
    t = torch_cuda
    m, r = golden("50_50")
    E, L = len(m["elem"]), 16
    omega = 2 * np.pi * 2e4
    p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"])
    perm = t.tensor(m["elem_perm"], device="cuda")
    var = t.full((E,), 3e-8 if is_complex else 0.0, dtype=t.float64, device="cuda")
    out = []
    for direct in (0, 1):
        _lib.set_option("direct_excitation", direct)
        try:
            p.assemble(perm, var, omega)
            p.factor(is_complex)
            J = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
            base = t.zeros(L * (L - 1), dtype=t.float64, device="cuda")
            p.jacobian_block(0, L, 0, E, 1e-3, omega, J, E, base)
            assert p.numeric_flag() == 0
            out.append((J.clone(), base.clone()))
        finally:
            _lib.set_option("direct_excitation", 0)
    assert (out[0][0] - out[1][0]).abs().max().item() < 1e-11
    assert (out[0][1] - out[1][1]).abs().max().item() < 1e-11
    assert (out[0][1] - 1).abs().max().item() < (1e-9 if is_complex else 5e-13)   # phi0 == 1 on the real base
    d0, d1 = out[0][0] - 1, out[1][0] - 1
    big = d1.abs() > 1e-9
    assert ((d0[big] - d1[big]).abs() / d1[big].abs()).max().item() < 1e-3
    p.close()


def test_overlapping_electrodes_use_the_fallback_paths(torch_cuda):
    """Electrodes wide enough to share nodes: U ranges are no longer contiguous per electrode, so the
    rank-update excitation stage and the patch Woodbury kernel hand over to the direct factorisation and
    the warp-per-column kernel.  Checked against the brute-force oracle."""
    t = torch_cuda
    n = 30
    cfg = meshgen.synthetic_config(n, 5, radius_mm=7.6)
    nodes, elem = meshgen.structured_mesh(n)
    param = o.elem_param(nodes, elem)
    els = o.electrode_elements(param, np.array(cfg["electrode_centers"]) / 1000, cfg["electrode_radius"] / 1000)
    sets = [set(np.unique(elem[np.asarray(e)].reshape(-1)).tolist()) for e in els]
    assert any(sets[i] & sets[j] for i in range(len(sets)) for j in range(i + 1, len(sets)))   # they overlap
    L, E = len(els), len(elem)
    omega = 2 * np.pi * cfg["signal_frequency"]
    p = _lib.Plan(nodes, elem, els)
    p.assemble(t.ones(E, dtype=t.float64, device="cuda"), t.zeros(E, dtype=t.float64, device="cuda"), omega)
    p.factor(False)
    J = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
    p.jacobian_block(0, L, 0, E, 1e-3, omega, J, E)
    assert p.numeric_flag() == 0
    rng = np.random.default_rng(2)
    pairs = [(int(rng.integers(L)), int(rng.integers(E))) for _ in range(10)] + [(1, int(els[1][0])), (2, int(els[3][0]))]
    cols = o.jacobian_columns_bruteforce(elem, param, np.ones(E), np.zeros(E), omega, els, pairs, 1e-3, len(nodes))
    Jh = J.cpu().numpy()
    for k, (i, j) in enumerate(pairs):
        assert np.abs(Jh[_rows(i, L), j] - cols[k]).max() < RAW_TOL
    p.close()


def test_32_electrodes_small_mesh_against_oracle(torch_cuda):
    s = _synthetic(torch_cuda, 70, 9)
    L, E = s["L"], s["E"]
    assert L == 32
    rng = np.random.default_rng(9)
    pairs = [(int(rng.integers(L)), int(rng.integers(E))) for _ in range(10)]
    cols = o.jacobian_columns_bruteforce(s["elem"], s["param"], np.ones(E), np.zeros(E), s["omega"], s["els"], pairs,
                                         1e-3, len(s["nodes"]))
    Jh = s["J"].cpu().numpy()
    for k, (i, j) in enumerate(pairs):
        assert np.abs(Jh[_rows(i, L), j] - cols[k]).max() < RAW_TOL
        sig = np.abs(cols[k] - 1).max()
        if sig > 1e-10:
            assert np.abs(Jh[_rows(i, L), j] - cols[k]).max() < 1e-3 * sig
    s["plan"].close()


def test_step_graph_replay_is_bit_identical(torch_cuda, golden):
    """The captured CUDA graph of a full step replays the same kernels: bit-identical J and inverse model,
    also after the material changed in the static buffers (arguments are frozen, data is not)."""
    torch = torch_cuda
    m, _ = golden("21_21")
    nodes, elem, els = m["nodes"], m["elem"], m["electrodes"]
    E, L = len(elem), len(els)
    omega = 2 * np.pi * 2e4
    plan = _lib.Plan(nodes, elem, els)
    perm = torch.tensor(m["elem_perm"], device="cuda")
    var = torch.zeros(E, dtype=torch.float64, device="cuda")
    det = torch.tensor(m["detection_index"], device="cuda")
    J = torch.zeros((L * (L - 1), E), dtype=torch.float64, device="cuda")

    def step():
        plan.assemble(perm, var, omega)
        plan.factor(False)
        plan.jacobian_block(0, L, 0, E, 1e-3, omega, J, E)
        return _lib.inverse_model(J, det, 289.0)

    inv0 = step().clone()
    J0 = J.clone()
    n0 = _lib.launch_count()
    g = _lib.StepGraph(step)
    assert g.kernels > 10
    J.zero_()
    inv1 = g.replay()
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 == 2 * g.kernels           # capture + one replay
    assert torch.equal(J, J0) and torch.equal(inv1, inv0)
    perm.mul_(1.5)                                             # new material in the same buffer
    inv2 = g.replay().clone()
    J2 = J.clone()
    inv3 = step()
    torch.cuda.synchronize()
    assert torch.equal(J, J2) and torch.equal(inv3, inv2)
    assert not torch.equal(inv2, inv0)
    plan.close()


def test_class_api_graph_cache(torch_cuda, workspace, golden):
    """EJAC through the engine's graph cache: eager call, captured call and replayed calls agree bit for bit
    and follow material changes (Example_03 flow repeated)."""
    workspace("21_21")
    from ceit_b200.EJAC import EJAC
    jac = EJAC()
    runs = []
    for _ in range(4):
        jac.fwd_FEM.invalidate()
        runs.append(np.array(jac.JAC_calculation()))
    assert all(np.array_equal(runs[0], r) for r in runs[1:])
    eng = jac._eng
    assert any(not isinstance(v, str) for v in eng._graphs.values())       # something was captured
    jac.fwd_FEM.mesh.elem_perm = jac.fwd_FEM.mesh.elem_perm * 2.0
    J2 = np.array(jac.JAC_calculation())
    eng.use_graphs = False
    jac.fwd_FEM.invalidate()
    J3 = np.array(jac.JAC_calculation())
    assert np.array_equal(J2, J3) and not np.array_equal(J2, runs[0])
    jac.save_inv_matrix(289)
    inv_a = jac.read_inv_matrix()
    eng.use_graphs = True
    for _ in range(3):
        jac.save_inv_matrix(289)
    assert np.array_equal(jac.read_inv_matrix(), inv_a)
