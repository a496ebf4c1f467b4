"""Multi-level nested-dissection base stage on the CPU: the library's symbolic tables (plan.cpp::build_tree_plan) and
level schedule (csrc/factor_tree.h) executed by the host backend of oracle/nd_host_check.cpp and compared with a dense
numpy inverse of K00 = K[F0, F0].  This is the check that lets the CUDA backend be debugged before any GPU time:
the GPU path runs the same schedule with the library's batched kernels."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ceit_oracle as o  # noqa: E402

from ceit_b200 import meshgen  # noqa: E402

SO = os.path.join(ROOT, "oracle", "_build", "libnd_host_check.so")
SRCS = [os.path.join(ROOT, "oracle", "nd_host_check.cpp"), os.path.join(ROOT, "ceit_b200", "csrc", "plan.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "ceit_b200", "csrc", f) for f in ("plan.h", "factor_tree.h")]


@pytest.fixture(scope="module")
def checker():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(f) for f in DEPS):
        subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", SO] + SRCS, check=True)
    L = ctypes.CDLL(SO)
    L.nd_host_last_error.restype = ctypes.c_char_p
    L.nd_host_check.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int64] * 2 + [ctypes.c_void_p] * 2 + \
        [ctypes.c_int64] * 2 + [ctypes.c_void_p] * 7 + [ctypes.c_int] * 2
    L.nd_host_check_parts.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int64] * 2 + [ctypes.c_void_p] * 2 + \
        [ctypes.c_int64] * 2 + [ctypes.c_void_p] + [ctypes.c_int] * 2 + [ctypes.c_void_p] * 7
    return L


def _run(L, nodes, elem, electrodes, leaf_cap, K, mode=0, ways=4):
    nodes = np.ascontiguousarray(nodes, np.float64)
    elem = np.ascontiguousarray(elem, np.int64)
    N, E = len(nodes), len(elem)
    ptr = np.zeros(len(electrodes) + 1, np.int64)
    for i, e in enumerate(electrodes):
        ptr[i + 1] = ptr[i] + len(e)
    idx = np.ascontiguousarray(np.concatenate([np.asarray(e, np.int64) for e in electrodes]))
    P = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)  # noqa: E731
    info = np.zeros(8, np.int64)
    rc = L.nd_host_check(P(nodes), P(elem), N, E, P(ptr), P(idx), len(electrodes), leaf_cap, None, P(info), None, None,
                         None, None, None, mode, ways)
    assert rc == 0, L.nd_host_last_error()
    nU, N0 = int(info[0]), int(info[1])
    pos, U = np.zeros(N, np.int64), np.zeros(nU, np.int64)
    X, SU, g0 = np.zeros((N0, nU)), np.zeros((nU, nU)), np.zeros((E, 6))
    Kd = np.ascontiguousarray(K, np.float64)
    rc = L.nd_host_check(P(nodes), P(elem), N, E, P(ptr), P(idx), len(electrodes), leaf_cap, P(Kd), P(info), P(pos),
                         P(U), P(X), P(SU), P(g0), mode, ways)
    assert rc == 0, L.nd_host_last_error()
    return info, pos, U, X, SU, g0


def _dense_reference(K, elem, pos, U):
    N = K.shape[0]
    N0 = N - len(U)
    node_of = np.empty(N, np.int64)
    node_of[pos] = np.arange(N)
    F0 = node_of[:N0]                                   # F0 nodes in compact position order
    K00, K0U, KUU = K[np.ix_(F0, F0)], K[np.ix_(F0, U)], K[np.ix_(U, U)]
    G = np.linalg.inv(K00)
    X = G @ K0U
    SU = KUU - K0U.T @ X
    g0 = np.zeros((len(elem), 6))
    pa, pb = [0, 1, 2, 0, 1, 2], [0, 1, 2, 1, 2, 0]
    for q in range(6):
        a, b = pos[elem[:, pa[q]]], pos[elem[:, pb[q]]]
        ok = (a < N0) & (b < N0)
        g0[ok, q] = G[a[ok], b[ok]]
    return X, SU, g0


def _check(L, nodes, elem, electrodes, leaf_cap, perm=None):
    N, E = len(nodes), len(elem)
    param = o.elem_param(nodes, elem)
    K = o.assemble_csr(elem, param, np.ones(E) if perm is None else perm, np.zeros(E), 2 * np.pi * 2e4, N)
    K = np.asarray(K.todense()).real
    # the schedules of the plan family: destination-centric pull with the separate steps (default), with the fused
    # front step on small fronts, children pushing (no pull levels); 4-way and 2-way dissection
    for mode, ways in ((1, 4), (0, 4), (2, 4), (1, 2), (2, 2)):
        info, pos, U, X, SU, g0 = _run(L, nodes, elem, electrodes, leaf_cap, K, mode, ways)
        assert info[7] == 0                                  # every pivot positive
        assert (info[6] // 100000 > 0) == (mode != 2)        # pull levels exist unless disabled
        info[6] %= 100000
        Xr, SUr, g0r = _dense_reference(K, np.asarray(elem), pos, U)
        assert np.abs(X - Xr).max() <= 1e-11 * max(1.0, np.abs(Xr).max())
        assert np.abs(SU - SUr).max() <= 1e-11 * np.abs(SUr).max()
        assert np.abs(g0 - g0r).max() <= 1e-11 * np.abs(g0r).max()
    return info


@pytest.mark.parametrize("tag,leaf_cap", [("21_21", 64), ("21_21", 16), ("21_21", 9), ("50_50", 64), ("50_50", 40)])
def test_tree_base_stage_matches_dense_inverse_shipped_meshes(checker, golden, tag, leaf_cap):
    m, _ = golden(tag)
    info = _check(checker, m["nodes"], m["elem"], m["electrodes"], leaf_cap, m["elem_perm"])
    assert info[2] >= 3 and info[5] <= 64 + 7            # several levels; separators of these meshes fit one tile


@pytest.mark.parametrize("n,per_side,leaf_cap", [(30, 5, 64), (47, 5, 20), (40, 9, 32)])
def test_tree_base_stage_matches_dense_inverse_structured(checker, n, per_side, leaf_cap):
    cfg = meshgen.synthetic_config(n, per_side)
    nodes, elem = meshgen.structured_mesh(n)
    param = o.elem_param(nodes, elem)
    els = o.electrode_elements(param, np.array(cfg["electrode_centers"]) / 1000, cfg["electrode_radius"] / 1000)
    _check(checker, nodes, elem, els, leaf_cap)


def test_tree_handles_disconnected_and_tiny_subdomains(checker):
    """An L-shaped / pinched sheet: bisection produces empty separators and unbalanced sub-trees."""
    n = 24
    nodes, elem = meshgen.structured_mesh(n)
    cx, cy = nodes[elem].mean(axis=1).T
    keep = ~((np.abs(cx) < 0.004) & (cy > -0.015))        # cut a slot: two lobes joined by a thin bridge
    elem = elem[keep]
    used = np.unique(elem)
    remap = -np.ones(len(nodes), np.int64)
    remap[used] = np.arange(len(used))
    nodes, elem = nodes[used], remap[elem]
    param = o.elem_param(nodes, elem)
    centers = np.array([[-0.02, -0.02], [0.02, -0.02], [0.02, 0.02], [-0.02, 0.02]])
    els = o.electrode_elements(param, centers, 0.003)
    _check(checker, nodes, elem, els, 12)


def _check_parts(L, nodes, elem, electrodes, leaf_cap, parts, ways=4, perm=None):
    """Domain-decomposed base stage: `parts` simulated ranks, one exchange; every rank's rows of
    X = K00^-1 K0U, its copy of S_U and the selected inverse of its own elements against the dense inverse."""
    nodes = np.ascontiguousarray(nodes, np.float64)
    elem = np.ascontiguousarray(elem, np.int64)
    N, E = len(nodes), len(elem)
    param = o.elem_param(nodes, elem)
    K = o.assemble_csr(elem, param, np.ones(E) if perm is None else perm, np.zeros(E), 2 * np.pi * 2e4, N)
    K = np.ascontiguousarray(np.asarray(K.todense()).real)
    ptr = np.zeros(len(electrodes) + 1, np.int64)
    for i, e in enumerate(electrodes):
        ptr[i + 1] = ptr[i] + len(e)
    idx = np.ascontiguousarray(np.concatenate([np.asarray(e, np.int64) for e in electrodes]))
    P = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)  # noqa: E731
    info = np.zeros(16, np.int64)
    n0c = np.zeros(parts, np.int64)
    pos = np.zeros((parts, N), np.int64)
    owner = np.zeros(E, np.int64)
    rc = L.nd_host_check_parts(P(nodes), P(elem), N, E, P(ptr), P(idx), len(electrodes), leaf_cap, None, parts, ways,
                               P(info), P(n0c), P(pos), None, None, None, P(owner))
    assert rc == 0, L.nd_host_last_error()
    nU, N0 = int(info[0]), int(info[1])
    assert owner.min() >= 0                                  # every element has exactly one owner
    assert len(np.unique(owner)) == parts                    # and every rank has work
    assert n0c.max() < N0 and n0c.sum() >= N0                # each rank holds a strict subset; together they cover F0
    X = np.full((parts, N0, nU), np.nan)
    SU = np.zeros((parts, nU, nU))
    g0 = np.zeros((E, 6))
    rc = L.nd_host_check_parts(P(nodes), P(elem), N, E, P(ptr), P(idx), len(electrodes), leaf_cap, P(K), parts, ways,
                               P(info), P(n0c), P(pos), P(X), P(SU), P(g0), P(owner))
    assert rc == 0, L.nd_host_last_error()
    assert info[7] == 0
    # dense reference in node numbering
    is_u = np.zeros(N, bool)
    U = np.array([v for v in range(N) if False], np.int64)
    # U in the plan's order = positions n0c[r] .. n0c[r] + nU of rank 0
    inv_pos0 = {int(q): v for v, q in enumerate(pos[0]) if q >= 0}
    U = np.array([inv_pos0[int(n0c[0]) + u] for u in range(nU)], np.int64)
    is_u[U] = True
    F0 = np.flatnonzero(~is_u)
    G = np.linalg.inv(K[np.ix_(F0, F0)])
    Xr = G @ K[np.ix_(F0, U)]
    SUr = K[np.ix_(U, U)] - K[np.ix_(F0, U)].T @ Xr
    loc = -np.ones(N, np.int64)
    loc[F0] = np.arange(len(F0))
    covered = np.zeros(N, bool)
    for r in range(parts):
        mine = np.flatnonzero((pos[r] >= 0) & ~is_u)
        assert len(mine) == n0c[r]
        covered[mine] = True
        assert np.abs(X[r, pos[r, mine]] - Xr[loc[mine]]).max() <= 1e-11 * max(1.0, np.abs(Xr).max())
        assert np.abs(SU[r] - SUr).max() <= 1e-11 * np.abs(SUr).max()
        assert (pos[r, U] == n0c[r] + np.arange(nU)).all()
        # the nodes of the rank's elements all have rows on the rank
        assert (pos[r][elem[owner == r]] >= 0).all()
    assert covered[F0].all()
    g0r = np.zeros((E, 6))
    pa, pb = [0, 1, 2, 0, 1, 2], [0, 1, 2, 1, 2, 0]
    for q in range(6):
        a, b = loc[elem[:, pa[q]]], loc[elem[:, pb[q]]]
        ok = (a >= 0) & (b >= 0)
        g0r[ok, q] = G[a[ok], b[ok]]
    assert np.abs(g0 - g0r).max() <= 1e-11 * np.abs(g0r).max()
    return info, n0c, owner


@pytest.mark.parametrize("tag,leaf_cap,parts", [("50_50", 64, 2), ("50_50", 64, 4), ("50_50", 20, 8), ("21_21", 9, 4)])
def test_domain_decomposed_base_stage_shipped_meshes(checker, golden, tag, leaf_cap, parts):
    m, _ = golden(tag)
    info, n0c, owner = _check_parts(checker, m["nodes"], m["elem"], m["electrodes"], leaf_cap, parts, perm=m["elem_perm"])
    assert info[5] > 0 and info[6] >= 1                      # rank-owned segments and export slots exist


@pytest.mark.parametrize("n,per_side,leaf_cap,parts,ways", [(40, 5, 16, 8, 4), (47, 5, 20, 3, 4), (36, 9, 16, 4, 2)])
def test_domain_decomposed_base_stage_structured(checker, n, per_side, leaf_cap, parts, ways):
    cfg = meshgen.synthetic_config(n, per_side)
    nodes, elem = meshgen.structured_mesh(n)
    param = o.elem_param(nodes, elem)
    els = o.electrode_elements(param, np.array(cfg["electrode_centers"]) / 1000, cfg["electrode_radius"] / 1000)
    info, n0c, owner = _check_parts(checker, nodes, elem, els, leaf_cap, parts, ways)
    # balanced: no rank owns more than twice the mean number of elements
    counts = np.bincount(owner, minlength=parts)
    assert counts.max() <= 2.0 * counts.mean()
