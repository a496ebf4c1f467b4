"""world_size-2 (and 3) gloo run of the Jacobian all-gather on CPU tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ceit_b200.sharding import ShardedStep, gather_jacobian, shard_ranges


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, L, E, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    full = torch.arange(L * (L - 1) * E, dtype=torch.float64).reshape(L * (L - 1), E) * 0.5 + 1.0
    i0, i1, e0, e1 = shard_ranges(L, E, ws, rank)
    J = torch.zeros_like(full)
    J[i0 * (L - 1): i1 * (L - 1), e0:e1] = full[i0 * (L - 1): i1 * (L - 1), e0:e1]
    out = gather_jacobian(J, L, E)
    q.put((rank, bool(torch.equal(out, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("ws,L,E", [(2, 5, 37), (3, 4, 20), (3, 2, 11)])
def test_gather_jacobian_gloo(ws, L, E):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, L, E, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)


def _worker_step(rank, ws, port, L, E, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    m = L * (L - 1)
    full = torch.arange(m * E, dtype=torch.float64).reshape(m, E) * 0.25 + 1.0
    det = torch.tensor([e for e in range(E) if e % 3 != 1], dtype=torch.int64)
    i0, i1, e0, e1 = shard_ranges(L, E, ws, rank)
    J = torch.zeros_like(full)
    J[i0 * (L - 1): i1 * (L - 1), e0:e1] = full[i0 * (L - 1): i1 * (L - 1), e0:e1]
    step = ShardedStep(L, E, det)
    Jr, cols = step.exchange(J)
    step.finish()
    lo, hi = step.bounds[rank], step.bounds[rank + 1]
    ok = bool(torch.equal(J, full))                                       # J replicated
    ok &= bool(torch.equal(Jr.index_select(1, cols), full[:, det[lo:hi]]))  # all rows at this rank's detection slice
    ok &= sum(step.bounds[r + 1] - step.bounds[r] for r in range(ws)) == det.numel()
    q.put((rank, ok, step.equal_rows))
    dist.destroy_process_group()


@pytest.mark.parametrize("ws,L,E,equal", [(2, 4, 37, True), (3, 6, 20, True), (3, 4, 20, False), (3, 2, 11, False)])
def test_sharded_step_exchange_gloo(ws, L, E, equal):
    """all-to-all of the detection-column slices + replication of J (equal row blocks), and the gather fallback"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_step, args=(r, ws, port, L, E, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res) and all(eq == equal for _, _, eq in res)


def _worker_domain(rank, ws, port, q):
    """Host logic of the domain-decomposed schedule on CPU tensors: element ownership from the decomposed plans
    (host tables only), replication of J from the ranks' own columns, the detection split that feeds the sharded
    inverse model."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from ceit_b200 import _lib
        from ceit_b200.sharding import DomainStep
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        m = dict(np.load(os.path.join(root, "tests", "golden", "mesh_50_50.npz")))
        ptr, ee = m["electrode_ptr"], m["electrode_elems"]
        els = [ee[ptr[i]:ptr[i + 1]] for i in range(16)]
        E, L = len(m["elem"]), 16
        mrows = L * (L - 1)
        det = torch.tensor(m["detection_index"], dtype=torch.int64)
        plan = _lib.Plan(m["nodes"], m["elem"], els, upload=False, parts=ws, part=rank)
        step = DomainStep(m["nodes"], m["elem"], els, det, plan=plan)
        full = (torch.arange(mrows * E, dtype=torch.float64).reshape(mrows, E) * 0.125 + 1.0)
        J = torch.zeros_like(full)
        J[:, step.own_idx] = full[:, step.own_idx]                        # what the rank's Woodbury kernel writes
        for _ in range(2):                                                # buffers are reusable
            step.gather_jacobian(J, async_op=True)
            step.finish()
        ok = bool(torch.equal(J, full))
        # detection split: the ranks' rows of inv_mat partition the detection elements, in the order of `det`
        n_det = torch.tensor([step.det_pos.numel()], dtype=torch.int64)
        dist.all_reduce(n_det)
        ok &= int(n_det.item()) == det.numel()
        ok &= bool(torch.equal(det[step.det_pos], step.det_local))
        own = torch.zeros(E, dtype=torch.bool)
        own[step.own_idx] = True
        ok &= bool(own[step.det_local].all())
        ok &= sum(step.counts) == E and step.counts[rank] == step.own_idx.numel()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws", [2, 3])
def test_domain_step_host_logic_gloo(ws):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_domain, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)
