"""numpy MODEL of the algorithm the CUDA path implements (TEST INFRASTRUCTURE, see ceit_oracle.py).

Not a restatement of the reference: this is the B200 algorithm written with
dense numpy so that every intermediate the kernels produce (level-set blocks, block Cholesky,
selected inverse, X, S_U, padded S^-1, Y_i, phi0, J) can be compared on the CPU, and so the maths is
validated against the reference goldens before any GPU time is spent.

  U   = all electrode nodes (Dirichlet for some excitation), F0 = the rest.
  K00 = K[F0,F0] is factorised ONCE (block-tridiagonal over BFS level sets of the F0 graph).
  X   = K00^-1 K0U, S_U = K_UU - K_U0 X (Schur complement on the electrode nodes),
  G0  = selected inverse of K00 on the mesh pattern (block-tridiagonal recursion).
  Excitation i (Dirichlet set B_i, R_i = U \\ B_i):  Sinv_i = padded inverse of S_U[R_i,R_i];
  Xh = [X ; -I] (N x nU), Yh_i = Xh Sinv_i.  Then for any nodes a, b
      Kff^-1[a,b] = G0[a,b] + sum_k Yh_i[a,k] Xh[b,k],   Kff^-1[U-node k, a] = -Yh_i[a,k]
  and the per-element Woodbury step is uniform over node classes.
"""
import numpy as np


def electrode_node_sets(elem, electrodes):
    return [np.unique(elem[np.asarray(e)].reshape(-1)) for e in electrodes]


def union_nodes(node_sets, N):
    """U ordering = electrode order, first occurrence wins."""
    u_of = -np.ones(N, dtype=np.int64)
    U = []
    for s in node_sets:
        for n in s:
            if u_of[n] < 0:
                u_of[n] = len(U)
                U.append(int(n))
    return np.array(U, dtype=np.int64), u_of


def adjacency(elem, N):
    import scipy.sparse as sp
    r = np.repeat(elem, 3, axis=1).reshape(-1)
    c = np.tile(elem, (1, 3)).reshape(-1)
    A = sp.coo_matrix((np.ones(len(r)), (r, c)), shape=(N, N)).tocsr()
    A.sum_duplicates()
    return A


def level_sets(A, keep):
    """BFS level sets of the graph induced on `keep` (bool mask). Start: pseudo-peripheral node
    (lowest index node, then two sweeps to the farthest node). Restarts on new components."""
    N = A.shape[0]
    indptr, indices = A.indptr, A.indices
    level = -np.ones(N, dtype=np.int64)

    def bfs(start, base, commit):
        lv = {start: base}
        frontier = [start]
        order = [[start]]
        while frontier:
            nxt = []
            for v in frontier:
                for w in indices[indptr[v]:indptr[v + 1]]:
                    if keep[w] and level[w] < 0 and w not in lv:
                        lv[w] = lv[v] + 1
                        nxt.append(int(w))
            if nxt:
                nxt.sort()
                order.append(nxt)
            frontier = nxt
        if commit:
            for k, v in lv.items():
                level[k] = v
        return order

    levels = []
    for s in range(N):
        if not keep[s] or level[s] >= 0:
            continue
        o = bfs(s, 0, False)
        far = o[-1][0]
        o = bfs(far, 0, False)
        far = o[-1][0]
        o = bfs(far, len(levels), True)
        levels += o
    return levels


def merge_levels(levels, target):
    """Merge consecutive levels until each block has >= target nodes (keeps block-tridiagonal)."""
    blocks, cur = [], []
    for lv in levels:
        cur += lv
        if len(cur) >= target:
            blocks.append(cur)
            cur = []
    if cur:
        if blocks and len(cur) < target // 2:
            blocks[-1] += cur
        else:
            blocks.append(cur)
    return blocks


def chol_sym(A):
    """L with L L^T = A (plain transpose; complex sqrt for complex symmetric A)."""
    n = A.shape[0]
    L = np.zeros_like(A)
    A = A.copy()
    for j in range(n):
        d = np.sqrt(A[j, j])
        L[j, j] = d
        L[j + 1:, j] = A[j + 1:, j] / d
        A[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
    return L


def block_tridiag_factor(K00, blk_ptr):
    """Block Cholesky of a block-tridiagonal matrix. Returns Linv_k, Lsub_k (= L_{k+1,k})."""
    nb = len(blk_ptr) - 1
    Linv, Lsub = [], []
    D = K00[blk_ptr[0]:blk_ptr[1], blk_ptr[0]:blk_ptr[1]].copy()
    for k in range(nb):
        L = chol_sym(D)
        Li = np.linalg.inv(L)
        Linv.append(Li)
        if k + 1 < nb:
            B = K00[blk_ptr[k + 1]:blk_ptr[k + 2], blk_ptr[k]:blk_ptr[k + 1]]
            Ls = B @ Li.T
            Lsub.append(Ls)
            D = K00[blk_ptr[k + 1]:blk_ptr[k + 2], blk_ptr[k + 1]:blk_ptr[k + 2]] - Ls @ Ls.T
    return Linv, Lsub


def block_tridiag_solve(Linv, Lsub, blk_ptr, RHS):
    nb = len(Linv)
    Z = [None] * nb
    for k in range(nb):
        t = RHS[blk_ptr[k]:blk_ptr[k + 1]]
        if k > 0:
            t = t - Lsub[k - 1] @ Z[k - 1]
        Z[k] = Linv[k] @ t
    X = np.zeros_like(RHS)
    for k in range(nb - 1, -1, -1):
        t = Z[k]
        if k + 1 < nb:
            t = t - Lsub[k].T @ X[blk_ptr[k + 1]:blk_ptr[k + 2]]
        X[blk_ptr[k]:blk_ptr[k + 1]] = Linv[k].T @ t
    return X


def block_tridiag_selinv(Linv, Lsub, blk_ptr):
    """Diagonal blocks Sig_kk and sub-diagonal blocks Sig_{k+1,k} of the inverse."""
    nb = len(Linv)
    Sd = [None] * nb
    So = [None] * (nb - 1)
    Sd[nb - 1] = Linv[nb - 1].T @ Linv[nb - 1]
    for k in range(nb - 2, -1, -1):
        P = Lsub[k] @ Linv[k]                      # B_k D_k^-1
        So[k] = -Sd[k + 1] @ P
        Sd[k] = Linv[k].T @ Linv[k] - P.T @ So[k]
    return Sd, So


class SchurModel:
    def __init__(self, nodes, elem, electrodes, K, target_block=64):
        """K: scipy CSR (N x N), complex or real (the doubled reference matrix)."""
        self.N = N = len(nodes)
        self.elem = elem
        self.L = len(electrodes)
        self.node_sets = electrode_node_sets(elem, electrodes)
        self.U, self.u_of = union_nodes(self.node_sets, N)
        self.nU = len(self.U)
        keep = self.u_of < 0
        A = adjacency(elem, N)
        self.levels = level_sets(A, keep)
        self.blocks = merge_levels(self.levels, target_block)
        self.perm0 = np.array([n for b in self.blocks for n in b], dtype=np.int64)
        self.N0 = len(self.perm0)
        assert self.N0 + self.nU == N
        self.blk_ptr = np.concatenate([[0], np.cumsum([len(b) for b in self.blocks])])
        self.pos = -np.ones(N, dtype=np.int64)
        self.pos[self.perm0] = np.arange(self.N0)
        self.pos[self.U] = self.N0 + np.arange(self.nU)
        full = np.concatenate([self.perm0, self.U])
        Kp = K.tocsr()[full][:, full].toarray()
        real = np.abs(Kp.imag).max() == 0 if np.iscomplexobj(Kp) else True
        if real:
            Kp = Kp.real.copy()
        self.real = real
        N0 = self.N0
        self.K00, self.K0U, self.KUU = Kp[:N0, :N0], Kp[:N0, N0:], Kp[N0:, N0:]
        # check block-tridiagonal structure
        blk_of = np.repeat(np.arange(len(self.blocks)), np.diff(self.blk_ptr))
        r, c = np.nonzero(self.K00)
        assert np.abs(blk_of[r] - blk_of[c]).max() <= 1
        self.Linv, self.Lsub = block_tridiag_factor(self.K00, self.blk_ptr)
        self.X = block_tridiag_solve(self.Linv, self.Lsub, self.blk_ptr, self.K0U)
        self.SU = self.KUU - self.K0U.T @ self.X
        self.Sd, self.So = block_tridiag_selinv(self.Linv, self.Lsub, self.blk_ptr)
        self.Xh = np.concatenate([self.X, -np.eye(self.nU, dtype=self.X.dtype)], 0)   # rows in pos order
        # node -> electrode weights on U
        self.W = np.zeros((self.L, self.nU))
        for m, es in enumerate(electrodes):
            for e in es:
                self.W[m, self.u_of[elem[e]]] += 1.0 / (3 * len(es))
        self.blk_of = blk_of

    def G0(self, a, b):
        """K00^-1[pos a, pos b] for positions inside F0 (must be in same/adjacent blocks)."""
        if a >= self.N0 or b >= self.N0:
            return 0.0
        ka, kb = self.blk_of[a], self.blk_of[b]
        if ka == kb:
            return self.Sd[ka][a - self.blk_ptr[ka], b - self.blk_ptr[ka]]
        if ka == kb + 1:
            return self.So[kb][a - self.blk_ptr[ka], b - self.blk_ptr[kb]]
        if kb == ka + 1:
            return self.So[ka][b - self.blk_ptr[kb], a - self.blk_ptr[ka]]
        raise AssertionError("pair outside block-tridiagonal pattern")

    def excitation(self, i):
        """Padded Sinv_i (nU x nU), Yh_i (N x nU), phi0 (N, in pos order)."""
        inB = np.zeros(self.nU, dtype=bool)
        inB[self.u_of[self.node_sets[i]]] = True
        St = self.SU.copy()
        St[inB, :] = 0
        St[:, inB] = 0
        St[inB, inB] = 1
        L = chol_sym(St)
        Li = np.linalg.inv(L)
        Sinv = Li.T @ Li
        Sinv[inB, :] = 0
        Sinv[:, inB] = 0
        Yh = self.Xh @ Sinv
        phiU = -Sinv @ (self.SU[:, inB].sum(1))
        phiU[inB] = 1
        phi0 = np.concatenate([-self.X @ phiU, phiU])
        return Sinv, Yh, phi0, inB

    def jacobian(self, area, omega, c, i_range=None):
        L, E = self.L, len(self.elem)
        J = np.zeros((L * (L - 1), E))
        M = (np.ones((3, 3)) + np.eye(3)) / 12.0
        base = np.zeros(L * (L - 1))
        for i in (range(L) if i_range is None else i_range):
            Sinv, Yh, phi0, inB = self.excitation(i)
            rows = [m for m in range(L) if m != i]
            base[i * (L - 1):(i + 1) * (L - 1)] = np.abs(self.W[rows] @ np.abs(phi0[self.N0:]))
            for j in range(E):
                p = self.pos[self.elem[j]]
                G = np.array([[self.G0(a, b) for b in p] for a in p]) + Yh[p] @ self.Xh[p].T
                D = 2j * omega * c * area[j] * M
                r = D @ phi0[p]
                y = np.linalg.solve(np.eye(3) + D @ G, r)
                dU = Yh[p].T @ y                          # delta at the U nodes
                amp = np.abs(phi0[self.N0:] + dU)
                J[i * (L - 1):(i + 1) * (L - 1), j] = np.abs(self.W[rows] @ amp)
        return J, base


# ------------------------------------------------------------------------------------------------------------------
# Partitioned elimination of the block-tridiagonal chain: the model of how G ranks would share the level-2 system
# (not on the product path yet).  G-1 single blocks are chosen as INTERFACES; removing them leaves G
# mutually disconnected SEGMENTS, one per rank.  Rank g, independently:
#     factorises its segment A_g (block-tridiagonal, the code above),
#     W_g = A_g^-1 C_g      ("spikes": C_g couples the segment's first / last block to the interface on its left / right),
#     t_g = A_g^-1 R_g,     S_g = C_g^T W_g,     r_g = C_g^T t_g.
# One exchange (all-gather of S_g: <= 2x2 interface blocks, and r_g: <= 2 interface row blocks per rank) builds the
# REDUCED system on the interfaces, A_red = A_I - sum_g S_g (block-tridiagonal again, G-1 blocks), which every rank
# solves redundantly.  Then, again independently per rank,
#     x_g = t_g - W_g x_I[lr],       Sigma_g = A_g^-1 + W_g Sigma_I[lr,lr] W_g^T   (only its tridiagonal pattern),
#     Sigma[g-block, interface] = -W_g Sigma_I[lr, interface]                      (only the adjacent blocks),
#     R^T M^-1 R = sum_g R_g^T t_g + (R_I - sum r_g)^T A_red^-1 (R_I - sum r_g)     (the S_U contribution).
# Serial chain per rank: nb/G blocks + G-1 reduced blocks instead of nb/2 + 1.
# ------------------------------------------------------------------------------------------------------------------
def choose_interfaces(nb, G):
    """G-1 interface blocks spread over 0..nb-1 with non-empty segments between them (fewer if nb is too small)."""
    G = max(1, min(G, (nb + 1) // 2))
    if G == 1:
        return []
    cuts = sorted({int(round((g * nb) / G - 0.5)) for g in range(1, G)})
    cuts = [c for c in cuts if 0 < c < nb - 1]
    out = []
    for c in cuts:                                   # keep at least one block between consecutive interfaces
        if not out or c - out[-1] >= 2:
            out.append(c)
    return out


def partitioned_chain(M, blk_ptr, RHS, G):
    """Model of the G-rank elimination of the SPD block-tridiagonal M (dense storage, blocks blk_ptr) with right-hand
    sides RHS.  Returns X = M^-1 RHS, the diagonal / sub-diagonal blocks of M^-1 and RHS^T M^-1 RHS, computed only from
    per-segment quantities, the exchanged (S_g, r_g) and the reduced interface system."""
    nb = len(blk_ptr) - 1
    sl = lambda k: slice(blk_ptr[k], blk_ptr[k + 1])
    itf = choose_interfaces(nb, G)
    bounds = [-1] + itf + [nb]
    segs = [list(range(bounds[g] + 1, bounds[g + 1])) for g in range(len(bounds) - 1)]
    nI = len(itf)
    X = np.zeros_like(RHS)
    Sd, So = [None] * nb, [None] * (nb - 1)
    quad = np.zeros((RHS.shape[1], RHS.shape[1]), dtype=RHS.dtype)
    per_rank, S_all, r_all = [], [], []
    # ---- phase 1 (per rank, no communication) ----------------------------------------------------------------
    for g, seg in enumerate(segs):
        rows = np.concatenate([np.arange(blk_ptr[k], blk_ptr[k + 1]) for k in seg])
        ptr = np.concatenate([[0], np.cumsum([blk_ptr[k + 1] - blk_ptr[k] for k in seg])])
        A = M[np.ix_(rows, rows)]
        Linv, Lsub = block_tridiag_factor(A, ptr)
        left = bounds[g] if bounds[g] >= 0 else None
        right = bounds[g + 1] if bounds[g + 1] < nb else None
        cols = [k for k in (left, right) if k is not None]
        C = np.zeros((len(rows), sum(blk_ptr[k + 1] - blk_ptr[k] for k in cols)), dtype=M.dtype)
        off, col_sl = 0, {}
        for k in cols:
            w = blk_ptr[k + 1] - blk_ptr[k]
            col_sl[k] = slice(off, off + w)
            C[:, off:off + w] = M[rows, sl(k)]          # non-zero only in the segment block adjacent to interface k
            off += w
        both = block_tridiag_solve(Linv, Lsub, ptr, np.concatenate([C, RHS[rows]], axis=1))
        W, t = both[:, :C.shape[1]], both[:, C.shape[1]:]
        quad += RHS[rows].T @ t
        S_all.append((cols, col_sl, C.T @ W))            # exchanged: (n_l + n_r)^2
        r_all.append((cols, col_sl, C.T @ t))            # exchanged: (n_l + n_r) x nrhs
        per_rank.append((seg, rows, ptr, Linv, Lsub, W, t, cols, col_sl))
    # ---- phase 2 (after ONE all-gather; every rank redundantly): reduced interface system -------------------------
    if nI:
        iptr = np.concatenate([[0], np.cumsum([blk_ptr[k + 1] - blk_ptr[k] for k in itf])])
        ipos = {k: q for q, k in enumerate(itf)}
        isl = lambda k: slice(iptr[ipos[k]], iptr[ipos[k] + 1])
        irows = np.concatenate([np.arange(blk_ptr[k], blk_ptr[k + 1]) for k in itf])
        Ared = M[np.ix_(irows, irows)].copy()            # block diagonal: consecutive interfaces are never adjacent
        Rred = RHS[irows].copy()
        for (cols, col_sl, S), (_, _, r) in zip(S_all, r_all):
            for a in cols:
                Rred[isl(a)] -= r[col_sl[a]]
                for b in cols:
                    Ared[isl(a), isl(b)] -= S[col_sl[a], col_sl[b]]
        Li, Ls = block_tridiag_factor(Ared, iptr)
        xI = block_tridiag_solve(Li, Ls, iptr, Rred)
        SdI, SoI = block_tridiag_selinv(Li, Ls, iptr)
        quad += Rred.T @ xI
        for q, k in enumerate(itf):
            X[sl(k)] = xI[iptr[q]:iptr[q + 1]]
            Sd[k] = SdI[q]

        def sig_I(a, b):                                  # Sigma_I block (a, b) for |pos(a) - pos(b)| <= 1
            qa, qb = ipos[a], ipos[b]
            if qa == qb:
                return SdI[qa]
            return SoI[qb] if qa == qb + 1 else SoI[qa].T
    # ---- phase 3 (per rank, no communication): back substitution and selected inverse of the segment -----------
    for seg, rows, ptr, Linv, Lsub, W, t, cols, col_sl in per_rank:
        x = t.copy()
        nc = W.shape[1]
        SigLR = np.zeros((nc, nc), dtype=M.dtype)
        if cols:
            xlr = np.concatenate([X[sl(k)] for k in cols])
            x -= W @ xlr
            for a in cols:
                for b in cols:
                    SigLR[col_sl[a], col_sl[b]] = sig_I(a, b)
        X[rows] = x
        sd, so = block_tridiag_selinv(Linv, Lsub, ptr)
        V = W @ SigLR                                    # segment rows x (n_l + n_r)
        for q, k in enumerate(seg):
            rq = slice(ptr[q], ptr[q + 1])
            Sd[k] = sd[q] + V[rq] @ W[rq].T
            if q + 1 < len(seg):
                rn = slice(ptr[q + 1], ptr[q + 2])
                So[k] = so[q] + V[rn] @ W[rq].T
        # blocks between the segment's end blocks and its interfaces: Sigma[seg row, interface] = -(W Sigma_I)[., interface]
        for a in cols:
            if a == seg[0] - 1:                          # interface on the left: sub-diagonal block (seg[0], a)
                So[a] = -V[slice(ptr[0], ptr[1]), col_sl[a]]
            if a == seg[-1] + 1:                         # interface on the right: sub-diagonal block (a, seg[-1])
                So[seg[-1]] = -V[slice(ptr[len(seg) - 1], ptr[len(seg)]), col_sl[a]].T
    return X, Sd, So, quad
