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
