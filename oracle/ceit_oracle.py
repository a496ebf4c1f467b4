"""CPU ORACLE for the CEIT Jacobian / Gauss-Newton hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in numpy/scipy, the algorithm of the reference (zehao99/CEIT, pure
Python/numpy).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it, and only as the checker or the timed CPU baseline -
never on the product path (`ceit_b200/` must not import anything from `oracle/`).

Parity pin: the reference's own tests hold NO numeric vector for this path, so
the oracle is pinned against outputs of the reference itself, run in the build container by
`tests/golden/make_golden.py` and committed as `tests/golden/ref_*.npz`
(`tests/test_oracle_golden.py` checks every function below against them).

Three levels:
  L0' `forward_dense`        - line-by-line restatement of EFEM.calculation (dense K, Python
                               assembly loop, row/col swaps, np.linalg.inv); this is what the
                               cpu_baseline times.
  L1  `forward_sparse`       - same maths with scipy.sparse + splu (for meshes whose dense K
                               does not fit); `jacobian_columns_bruteforce` loops it.
  L2  `jacobian_woodbury`    - exact rank<=3 update of one factorised system per excitation
                               (SURVEY.md §7.1), full J in seconds.
Inverse model / realtime: `inverse_model`, `reconstruct` restate Solver.py:110-124 / :95-108.

All `file:line` citations are relative to the reference repository root.
"""
import numpy as np

try:  # scipy is only needed by L1/L2
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
except Exception:  # pragma: no cover
    sp = None
    spla = None


# --------------------------------------------------------------------------------------------
# mesh model (CEIT/models/mesh.py)
# --------------------------------------------------------------------------------------------
def elem_param(nodes, elem):
    """[A, b0,b1,b2, c0,c1,c2, xbar, ybar] per element, CEIT/models/mesh.py:41-75.

    The reference inverts P=[[x,y,1]]*3 numerically (mesh.py:60-67); rows 0/1 of P^-1 are the
    shape-function gradients, written here in closed form (differences are at the 1e-16 level).
    """
    x = nodes[elem, 0]
    y = nodes[elem, 1]
    det = (x[:, 0] * (y[:, 1] - y[:, 2]) - y[:, 0] * (x[:, 1] - x[:, 2])
           + (x[:, 1] * y[:, 2] - x[:, 2] * y[:, 1]))
    area = np.abs(det / 2)                                                     # mesh.py:62
    b = np.stack([y[:, 1] - y[:, 2], y[:, 2] - y[:, 0], y[:, 0] - y[:, 1]], 1) / det[:, None]
    c = np.stack([x[:, 2] - x[:, 1], x[:, 0] - x[:, 2], x[:, 1] - x[:, 0]], 1) / det[:, None]
    return np.concatenate([area[:, None], b, c, x.mean(1)[:, None], y.mean(1)[:, None]], 1)


def electrode_elements(param, centers, radius):
    """Element lists per electrode: centroid inside the closed square, mesh.py:90-113."""
    out = []
    for cx, cy in centers:
        m = ((cx + radius) >= param[:, 7]) & (param[:, 7] >= (cx - radius)) & \
            ((cy + radius) >= param[:, 8]) & (param[:, 8] >= (cy - radius))
        idx = np.nonzero(m)[0]
        if len(idx) == 0:
            raise Exception("No element is selected, please check the input")   # mesh.py:112-113
        out.append(idx)
    return out


def detection_index(nodes, elem, bound):
    """Elements whose centroid is strictly inside the detection square, mesh.py:127-155."""
    xv = (nodes[elem[:, 0], 0] + nodes[elem[:, 1], 0] + nodes[elem[:, 2], 0]) / 3
    yv = (nodes[elem[:, 0], 1] + nodes[elem[:, 1], 1] + nodes[elem[:, 2], 1]) / 3
    return np.nonzero((np.abs(xv) < bound) & (np.abs(yv) < bound))[0]


# --------------------------------------------------------------------------------------------
# forward model (CEIT/EFEM.py, CEIT/FEMBasic.py)
# --------------------------------------------------------------------------------------------
_PATTERN = [[0, 0], [1, 1], [2, 2], [0, 1], [1, 2], [2, 0], [1, 0], [2, 1], [0, 2]]  # EFEM.py:53-54


def local_matrices(param, perm, var, omega):
    """k_ab per element (E,3,3) complex, EFEM.py:59-60,64-65 (NOT yet doubled)."""
    A = param[:, 0]
    b = param[:, 1:4]
    c = param[:, 4:7]
    stiff = perm[:, None, None] * (b[:, :, None] * b[:, None, :] + c[:, :, None] * c[:, None, :]) \
        * (4 * A)[:, None, None]
    mass = np.where(np.eye(3, dtype=bool)[None], (omega * var * A / 6)[:, None, None],
                    (omega * var * A / 12)[:, None, None])
    return stiff + 1j * mass


def assemble_dense_loop(elem, param, perm, var, omega, N):
    """Literal restatement of EFEM.construct_sparse_matrix, EFEM.py:47-69 (Python loop,
    every pattern entry added to [i][j] AND [j][i] => the whole matrix is 2x)."""
    K = np.zeros((N, N), dtype=np.complex128)
    for index, element in enumerate(elem):
        p = param[index]
        for i, j in _PATTERN:
            if i != j:
                K_ij = perm[index] * (p[1 + i] * p[1 + j] + p[4 + i] * p[4 + j]) * (4 * p[0]) + \
                    (omega * var[index] * p[0] / 12) * 1j
            else:
                K_ij = perm[index] * (p[1 + i] * p[1 + j] + p[4 + i] * p[4 + j]) * (4 * p[0]) + \
                    (omega * var[index] * p[0] / 6) * 1j
            K[element[i]][element[j]] += K_ij
            K[element[j]][element[i]] += K_ij
    return K


def assemble_csr(elem, param, perm, var, omega, N):
    """Same matrix as `assemble_dense_loop` as scipy CSR (sorted indices, duplicates summed)."""
    k = 2.0 * local_matrices(param, perm, var, omega)
    rows = np.repeat(elem, 3, axis=1).reshape(-1)
    cols = np.tile(elem, (1, 3)).reshape(-1)
    K = sp.coo_matrix((k.reshape(-1), (rows, cols)), shape=(N, N)).tocsr()
    K.sum_duplicates()
    K.sort_indices()
    return K


def dirichlet_nodes(elem, electrode_elems_i):
    """Sorted unique nodes of the excitation electrode's elements, EFEM.py:79-86."""
    return np.unique(elem[np.asarray(electrode_elems_i)].reshape(-1))


def _amplitudes(phi_abs, elem, electrodes):
    """FEMBasic.py:118-137: abs per node (EFEM.py:44) -> mean of 3 per element -> mean per electrode."""
    elem_pot = (phi_abs[elem[:, 0]] + phi_abs[elem[:, 1]] + phi_abs[elem[:, 2]]) / 3
    elec = np.array([np.mean(elem_pot[np.asarray(e)]) for e in electrodes])
    return elem_pot, elec


def forward_dense(elem, param, perm, var, omega, electrodes, i, N):
    """EFEM.calculation(i) restated literally (FEMBasic.py:67-86, EFEM.py:37-45):
    dense assembly loop, symmetric swaps moving the Dirichlet nodes to the tail
    (EFEM.py:71-96,136-142), phi_f = -inv(K_ff) K_fb 1 (EFEM.py:111-119), un-permute
    (EFEM.py:128-134).  Returns (node_potential f64[N], element_potential, electrode_potential)."""
    K = assemble_dense_loop(elem, param, perm, var, omega, N)
    order = list(range(N))
    node_list = list(dirichlet_nodes(elem, electrodes[i]))
    nb = len(node_list)
    nf = N - nb
    index = N
    for n in node_list:
        if n < nf:
            index -= 1
            while index in node_list:
                index -= 1
            K[[n, index], :] = K[[index, n], :]
            K[:, [n, index]] = K[:, [index, n]]
            order[n], order[index] = order[index], order[n]
    pot_b = np.ones((nb, 1), dtype=np.complex128)
    pot_f = -np.dot(np.dot(np.linalg.inv(K[:nf, :nf]), K[:nf, nf:]), pot_b)
    phi = np.abs(np.append(pot_f.reshape(-1), pot_b.reshape(-1)))
    node_pot = np.zeros(N)
    node_pot[np.asarray(order)] = phi
    elem_pot, elec = _amplitudes(node_pot, elem, electrodes)
    return node_pot, elem_pot, elec


def forward_sparse(K, elem, electrodes, i, return_complex=False):
    """Same solve on a CSR K with splu (L1).  K from `assemble_csr`."""
    N = K.shape[0]
    bn = dirichlet_nodes(elem, electrodes[i])
    free = np.setdiff1d(np.arange(N), bn)
    Kc = K.tocsc()
    Kff = Kc[free][:, free]
    rhs = -np.asarray(Kc[free][:, bn].sum(axis=1)).reshape(-1)
    phi = np.ones(N, dtype=np.complex128)
    phi[free] = spla.splu(Kff.tocsc()).solve(rhs)
    if return_complex:
        return phi
    a = np.abs(phi)
    elem_pot, elec = _amplitudes(a, elem, electrodes)
    return a, elem_pot, elec


def jac_rows(elec, i, L):
    """Row block of excitation i: |electrode_potential[m]| for m != i ascending, EJAC.py:127-132."""
    return np.abs(np.array([elec[m] for m in range(L) if m != i]))


def jacobian_columns_bruteforce(elem, param, perm, base_var, omega, electrodes, pairs, c, N, dense=False):
    """Perturb -> re-assemble -> solve, one (i, j) pair at a time, EJAC.py:123-133."""
    L = len(electrodes)
    out = []
    for (i, j) in pairs:
        var = np.array(base_var, dtype=np.float64)
        var[j] += c
        if dense:
            _, _, elec = forward_dense(elem, param, perm, var, omega, electrodes, i, N)
        else:
            K = assemble_csr(elem, param, perm, var, omega, N)
            _, _, elec = forward_sparse(K, elem, electrodes, i)
        out.append(jac_rows(elec, i, L))
    return np.array(out)


def jacobian_woodbury(elem, param, perm, base_var, omega, electrodes, c, N, i_range=None):
    """Full J (L(L-1) x E) through the exact rank<=3 update:
    dK_j = 2*j*omega*c*A_j*M touches only element j's 3x3 block, so with G = Kff^-1[S,S],
    D = dK[S,S], r = (dK phi0)[S]:  dphi_f = -Kff^-1[:,S] (I + D G)^-1 r."""
    L = len(electrodes)
    E = len(elem)
    J = np.zeros((L * (L - 1), E))
    K = assemble_csr(elem, param, perm, np.asarray(base_var, dtype=np.float64), omega, N).tocsc()
    M = (np.ones((3, 3)) + np.eye(3)) / 12.0
    # node -> electrode weights (abs is per node, averaging is linear afterwards)
    W = np.zeros((L, N))
    for m, es in enumerate(electrodes):
        for e in es:
            W[m, elem[e]] += 1.0 / (3 * len(es))
    for i in (range(L) if i_range is None else i_range):
        bn = dirichlet_nodes(elem, electrodes[i])
        free = np.setdiff1d(np.arange(N), bn)
        pos = -np.ones(N, dtype=np.int64)
        pos[free] = np.arange(len(free))
        Kff = K[free][:, free].toarray()
        Kinv = np.linalg.inv(Kff)
        phi0 = np.ones(N, dtype=np.complex128)
        phi0[free] = Kinv @ (-np.asarray(K[free][:, bn].sum(axis=1)).reshape(-1))
        meas = np.nonzero(W.sum(0) > 0)[0]
        meas = meas[pos[meas] >= 0]
        rows = [m for m in range(L) if m != i]
        for j in range(E):
            tri = elem[j]
            dK = 2j * omega * c * param[j, 0] * M
            S = [a for a in range(3) if pos[tri[a]] >= 0]
            phi = phi0.copy()
            if S:
                ps = pos[tri[S]]
                G = Kinv[np.ix_(ps, ps)]
                D = dK[np.ix_(S, S)]
                r = (dK @ phi0[tri])[S]
                y = np.linalg.solve(np.eye(len(S)) + D @ G, r)
                phi[meas] = phi0[meas] - Kinv[np.ix_(pos[meas], ps)] @ y
            a = np.abs(phi)
            J[i * (L - 1):(i + 1) * (L - 1), j] = np.abs(W[rows] @ a)
    return J


# --------------------------------------------------------------------------------------------
# inverse model / realtime (CEIT/Solver.py)
# --------------------------------------------------------------------------------------------
def eliminate_non_detect(J, det_index):
    """J[:, detection_index], Solver.py:76-86 / EJAC.py:142-152."""
    return np.array(J[:, np.asarray(det_index)])


def inverse_model(J, det_index, lmbda):
    """inv(Jt^T Jt + lmbda^2 I) Jt^T with Jt = J[:, det] - 1, Solver.py:110-124 (identity
    regulariser, coefficient lmbda SQUARED)."""
    Jt = eliminate_non_detect(J, det_index) - 1
    Q = np.eye(Jt.shape[1])
    return np.dot(np.linalg.inv(np.dot(Jt.T, Jt) + lmbda ** 2 * Q), Jt.T)


def inverse_model_dual(J, det_index, lmbda):
    """Algebraically identical push-through form Jt^T (Jt Jt^T + lmbda^2 I)^-1 (m x m system)."""
    Jt = eliminate_non_detect(J, det_index) - 1
    return Jt.T @ np.linalg.inv(Jt @ Jt.T + lmbda ** 2 * np.eye(Jt.shape[0]))


def reconstruct(inv_mat, delta_V):
    """Solver.solve, Solver.py:95-108: np.dot(inv_mat, delta_V); (m,) or (m,F)."""
    return np.dot(inv_mat, delta_V)


# ---------------------------------------------------------------------------------------------------
# datahandler row of SURVEY.md 8(f): the consumers of Solver.solve outputs
# ---------------------------------------------------------------------------------------------------
def recording_mean(data, electrode_num=16, contain_excitation=False):
    """BasicData.__init__ / exclude_excitation (CEIT/datahandler/Data.py:18-30): mean over the recorded
    samples (rows); with contain_excitation the channels i with i % electrode_num == 0 are dropped."""
    mean = np.mean(np.asarray(data), axis=0)
    if contain_excitation:
        mean = np.array([v for i, v in enumerate(mean) if i % electrode_num != 0])
    return mean


def center_of_position(capacitance, detection_index, elem_param, threshold_percentage=0.9):
    """DataPlotter.get_COP after the solve (CEIT/datahandler/DataHandler.py:41-56): threshold between min and
    max, then the mean centroid / value of the detection elements strictly above it, accumulated in index order
    like the reference's loop.  Raises ZeroDivisionError (as the reference does) when nothing passes."""
    capacitance = np.asarray(capacitance, dtype=np.float64)
    threshold = np.max(capacitance) * threshold_percentage + np.min(capacitance) * (1 - threshold_percentage)
    x_sum = y_sum = v_sum = 0.0
    count = 0.0
    for i, c in enumerate(capacitance):
        if c > threshold:
            idx = detection_index[i]
            x_sum += elem_param[idx][7]
            y_sum += elem_param[idx][8]
            v_sum += c
            count += 1
    return x_sum / count, y_sum / count, v_sum / count
