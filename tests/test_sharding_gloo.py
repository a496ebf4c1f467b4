"""world_size-2 (and 3) gloo run of the Jacobian all-gather on CPU tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ceit_b200.sharding import gather_jacobian, shard_ranges


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
