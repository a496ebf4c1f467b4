"""Multi-GPU sharding of the Jacobian (one process per GPU, torch.distributed for the plumbing).

Every (excitation i, element j) column is independent once the base stage is factorised, so the
Jacobian shards with no data-path collective: excitation-major (contiguous row blocks of J, the
reference's own `calc_from`/`calc_end` axis, EJAC.py:107-119) and, when there are more ranks than
excitations or for balance, element ranges inside an excitation.  The single exchange of the path is
the all-gather that replicates J for the inverse model.
"""
import numpy as np


def shard_ranges(L, E, world_size, rank):
    """Return (i_from, i_to, e_from, e_to) for `rank`.

    world_size <= L : contiguous excitation blocks (sizes differ by at most 1), all elements.
    world_size  > L : ranks are grouped per excitation; each group splits the element range.
    """
    if world_size <= L:
        base, rem = divmod(L, world_size)
        i_from = rank * base + min(rank, rem)
        i_to = i_from + base + (1 if rank < rem else 0)
        return i_from, i_to, 0, E
    per = world_size // L            # ranks per excitation (the last excitations may get one more)
    extra = world_size - per * L
    bounds = np.concatenate([[0], np.cumsum([per + (1 if i >= L - extra else 0) for i in range(L)])])
    i = int(np.searchsorted(bounds, rank, side="right") - 1)
    g = int(bounds[i + 1] - bounds[i])
    k = rank - int(bounds[i])
    base, rem = divmod(E, g)
    e_from = k * base + min(k, rem)
    e_to = e_from + base + (1 if k < rem else 0)
    return i, i + 1, e_from, e_to


def all_shards(L, E, world_size):
    return [shard_ranges(L, E, world_size, r) for r in range(world_size)]


def gather_jacobian(J_local, L, E, group=None):
    """Replicate J (L(L-1) x E) on every rank from the per-rank shards.

    J_local: full-size tensor with this rank's block filled in (cuda for NCCL, cpu for gloo).
    Excitation-major shards are contiguous row blocks -> one all_gather of padded blocks;
    element-range shards (world_size > L) fall back to an all_reduce(SUM) of zero-padded blocks."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if ws == 1:
        return J_local
    shards = all_shards(L, E, ws)
    if ws <= L:
        rows = [(s[1] - s[0]) * (L - 1) for s in shards]
        mx = max(rows)
        mine = torch.zeros((mx, E), dtype=J_local.dtype, device=J_local.device)
        r0 = shards[rank][0] * (L - 1)
        mine[: rows[rank]] = J_local[r0: r0 + rows[rank]]
        out = torch.empty((ws * mx, E), dtype=J_local.dtype, device=J_local.device)
        dist.all_gather_into_tensor(out, mine, group=group)
        for r, s in enumerate(shards):
            J_local[s[0] * (L - 1): s[1] * (L - 1)] = out[r * mx: r * mx + rows[r]]
        return J_local
    i_from, i_to, e_from, e_to = shards[rank]
    mask = torch.zeros_like(J_local)
    mask[i_from * (L - 1): i_to * (L - 1), e_from:e_to] = J_local[i_from * (L - 1): i_to * (L - 1), e_from:e_to]
    dist.all_reduce(mask, op=dist.ReduceOp.SUM, group=group)
    J_local.copy_(mask)
    return J_local


class ShardedStep:
    """The exchange half of a multi-GPU step (one process per GPU, NCCL through torch.distributed).

    After the local part (each rank's excitation rows of J), the inverse model needs, per rank, ALL rows of J at
    that rank's slice of the detection elements, and the reference contract wants J replicated.  Two collectives:

      * critical path - an ALL-TO-ALL of the row block restricted to each peer's detection columns: 1/ws of the
        volume of an all-gather; it lands directly in the rows of the compact matrix J_r (m x n_det_local), from
        which the rank forms its partial Gram matrix (one all-reduce of m x m doubles) and its rows of inv_mat;
      * off the critical path - the replication of J as an IN-PLACE all-gather of the row blocks (no staging copies)
        on a second communicator, so it overlaps the Gram / Cholesky / inverse-model GEMMs; `finish()` joins it.

    Falls back to `gather_jacobian` + the replicated-J form when the shards are not equal row blocks (ws does not
    divide L, or ws > L: element-range shards).
    """

    def __init__(self, L, E, det, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.ws = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.L, self.E, self.m = L, E, L * (L - 1)
        self.shards = all_shards(L, E, self.ws)
        self.equal_rows = self.ws <= L and L % self.ws == 0
        n_det = det.numel()
        self.bounds = [(n_det * r) // self.ws for r in range(self.ws + 1)]
        self.det = det
        self.det_slices = [det[self.bounds[q]:self.bounds[q + 1]].contiguous() for q in range(self.ws)]
        self.det_local = self.det_slices[self.rank]
        self.nd_local = self.det_local.numel()
        self._work = None
        self.nccl = dist.get_backend(group) == "nccl"
        if self.equal_rows:
            # second communicator: the replication all-gather must not queue in front of the Gram all-reduce
            self.repl_group = dist.new_group(ranks=list(range(self.ws)), backend="nccl") if self.nccl else None
            self.J_r = torch.empty((self.m, max(self.nd_local, 1)), dtype=torch.float64, device=det.device)
            self.cols_local = torch.arange(self.nd_local, dtype=torch.int64, device=det.device)
            rows = self.m // self.ws
            self.recv = [self.J_r[p * rows:(p + 1) * rows] for p in range(self.ws)]
            self.send = [torch.empty((rows, self.det_slices[q].numel()), dtype=torch.float64, device=det.device)
                         for q in range(self.ws)]
            self.nvlink_bytes_per_step = 8 * rows * (n_det - self.nd_local) + 8 * rows * E * (self.ws - 1)

    def exchange(self, J):
        """Start both collectives; returns the matrix / index pair to feed the sharded inverse model with."""
        torch, dist = self.torch, self.dist
        if not self.equal_rows:
            gather_jacobian(J, self.L, self.E, self.group)
            return J, self.det_local
        rows = self.m // self.ws
        r0 = self.rank * rows
        mine = J[r0:r0 + rows]
        for q in range(self.ws):
            torch.index_select(mine, 1, self.det_slices[q], out=self.send[q])
        if self.nccl:
            dist.all_to_all(self.recv, self.send, group=self.group)
            self._work = dist.all_gather_into_tensor(J, mine, group=self.repl_group, async_op=True)
        else:
            # gloo (CPU tests of the host logic): no all_to_all, no in-place all-gather
            reqs = []
            for q in range(self.ws):
                if q == self.rank:
                    self.recv[q].copy_(self.send[q])
                else:
                    reqs.append(dist.isend(self.send[q], q, group=self.group))
                    reqs.append(dist.irecv(self.recv[q], q, group=self.group))
            for r in reqs:
                r.wait()
            dist.all_gather_into_tensor(J, mine.clone(), group=self.group)
        return self.J_r, self.cols_local

    def finish(self):
        """Join the replication of J (the current stream waits for the all-gather)."""
        if self._work is not None:
            self._work.wait()
            self._work = None


class _DevMem:
    """A library-owned device allocation seen as a 1-D float64 array (zero copy, __cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


class DomainStep:
    """One material state on G GPUs by DOMAIN DECOMPOSITION (north_star: "the element loop shards by element range").

    The reference's parallel axis is the excitation range (`calc_from` / `calc_end`, CEIT/EJAC.py:107-119); sharding
    the excitations leaves the base stage replicated.  Here the mesh itself is decomposed: the plan of rank r
    (`Plan(..., parts=G, part=r)`) owns the sub-trees of one region of the elimination forest, the top of the forest is
    replicated.  Per step:

      assemble -> factor_phase(1) -> exchange (ONE all-gather of the sub-tree roots' update matrices, ONE all-reduce of
      the top's right-hand-side rows + the partial S_U) -> factor_phase(2) -> jacobian: the columns of the rank's OWN
      elements for all L excitations -> inverse model: partial Gram matrix of the rank's own detection columns, one
      all-reduce of m x m, the rank's rows of inv_mat.

    The base stage, the per-excitation stage, the Woodbury kernel and the inverse model all work on 1/G of the mesh; J
    never has to be exchanged for the inverse model - `gather_jacobian()` replicates it only because the reference's
    contract returns the whole matrix (it can overlap the inverse model on a second communicator).
    """

    def __init__(self, nodes, elem, electrodes, det, group=None, plan=None):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.torch, self.dist, self._lib = torch, dist, _lib
        self.group = group
        self.ws = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.plan = plan if plan is not None else _lib.Plan(nodes, elem, electrodes, parts=self.ws, part=self.rank)
        self.L, self.E = self.plan.L, self.plan.E
        self.m = self.L * (self.L - 1)
        dev = det.device
        own = torch.from_numpy(self.plan.own_elements()).to(dev)
        self.own_idx = torch.nonzero(own).flatten()                       # element ids of this rank, ascending
        counts = torch.zeros(self.ws, dtype=torch.int64, device=dev)
        counts[self.rank] = self.own_idx.numel()
        dist.all_reduce(counts, group=group)
        self.counts = [int(c) for c in counts.tolist()]
        self.e_max = max(self.counts)
        det_own = own[det]
        self.det = det
        self.det_pos = torch.nonzero(det_own).flatten()                    # rows of inv_mat this rank produces
        self.det_local = det[det_own].contiguous()                         # their element ids (columns of J)
        self._xg = self._red = None
        self._pack = torch.zeros((self.m, self.e_max), dtype=torch.float64, device=dev)
        self._all = torch.empty((self.ws, self.m, self.e_max), dtype=torch.float64, device=dev)
        self._idx_all = None
        self.nccl = dist.get_backend(group) == "nccl"
        self.repl_group = dist.new_group(ranks=list(range(self.ws)), backend="nccl") if self.nccl else group
        self._work = None
        self.exchange_bytes = 0

    # ---- base stage -------------------------------------------------------------------------------------------
    def _buffers(self):
        info, g, r = self.plan.exchange_info()
        per, nred = int(info[2]), int(info[3])
        dev = self.det.device
        self._xg = self.torch.as_tensor(_DevMem(g, max(per * self.ws, 1)), device=dev) if g and per else None
        self._red = self.torch.as_tensor(_DevMem(r, nred), device=dev) if r and nred else None
        self._per = per
        self._gen = self.plan.generation()
        self.exchange_bytes = 8 * (per * (self.ws - 1) + 2 * nred)

    def base_exchange(self):
        """The one exchange of the base stage (between factor_phase(1) and factor_phase(2)), on the current stream."""
        if self._red is None or self._gen != self.plan.generation():
            self._buffers()
        if self._xg is not None:
            mine = self._xg[self.rank * self._per:(self.rank + 1) * self._per]
            self.dist.all_gather_into_tensor(self._xg, mine, group=self.group)
        self.dist.all_reduce(self._red, op=self.dist.ReduceOp.SUM, group=self.group)

    def factor(self, perm, var, omega):
        self.plan.assemble(perm, var, omega)
        self.plan.factor_phase(1)
        self.base_exchange()
        self.plan.factor_phase(2)

    # ---- Jacobian columns of this rank's elements ----------------------------------------------------------------
    def jacobian(self, c, omega, J, base=None):
        self.plan.jacobian_block(0, self.L, 0, self.E, c, omega, J, J.stride(0), base)

    def inverse_model(self, J, lmbda, shift=1.0, flag=None):
        """This rank's rows of inv_mat (rows `det_pos` of the whole model), from its own columns of J."""
        return self._lib.inverse_model_sharded(J, self.det_local, lmbda, shift=shift, group=self.group, flag=flag)

    # ---- replication of J (output contract only) ------------------------------------------------------------------
    def gather_jacobian(self, J, async_op=False):
        """Replicate J: pack this rank's columns, all-gather, scatter the other ranks' columns into place.  The whole
        chain runs on a side stream (and its own communicator), so with async_op=True it overlaps whatever the caller
        enqueues next on the current stream (the inverse model) until `finish()`."""
        torch, dist = self.torch, self.dist
        if self._idx_all is None:
            idx = torch.full((self.ws, self.e_max), -1, dtype=torch.int64, device=J.device)
            idx[self.rank, :self.own_idx.numel()] = self.own_idx
            mine = idx[self.rank].clone()
            dist.all_gather_into_tensor(idx.view(-1), mine, group=self.group)    # flat views: gloo and NCCL agree on them
            self._idx_all = [idx[r, :self.counts[r]].contiguous() for r in range(self.ws)]
            self._side = torch.cuda.Stream(device=J.device) if J.is_cuda else None
        n = self.own_idx.numel()
        if self._side is None:                                   # gloo / CPU: everything in line
            if n:
                torch.index_select(J, 1, self.own_idx, out=self._pack[:, :n])
            dist.all_gather_into_tensor(self._all.view(-1), self._pack.view(-1), group=self.group)
            self._scatter(J)
            return J
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            if n:
                torch.index_select(J, 1, self.own_idx, out=self._pack[:, :n])
            dist.all_gather_into_tensor(self._all.view(-1), self._pack.view(-1), group=self.repl_group)
            self._scatter(J)
        self._work = True
        if not async_op:
            self.finish()
        return J

    def _scatter(self, J):
        for r in range(self.ws):
            if r != self.rank and self.counts[r]:
                J.index_copy_(1, self._idx_all[r], self._all[r][:, :self.counts[r]])

    def finish(self):
        """Join the replication of J: the current stream waits for the side stream."""
        if self._work:
            self.torch.cuda.current_stream().wait_stream(self._side)
            self._work = None
