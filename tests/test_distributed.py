"""CPU, world_size 2 over gloo: the data-parallel exchange step (one all-reduce of the flat gradient bucket)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from slowfast_b200.engine import allreduce_flat_gradients, flat_offsets
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(6, 5), nn.BatchNorm1d(5), nn.Linear(5, 3))
    params = list(net.parameters())
    offsets, total = flat_offsets(params)  # slots start on 256-byte boundaries
    assert all(o % 64 == 0 for o in offsets) and total >= sum(p.numel() for p in params)
    # each rank's "backward" fills a flat bucket with rank-dependent values; grads start as private copies
    flat = torch.arange(total, dtype=torch.float32) * (rank + 1)
    for p, off in zip(params, offsets):
        p.grad = flat[off:off + p.numel()].view_as(p).clone()
    allreduce_flat_gradients(flat, params)
    expect = torch.arange(total, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(flat, expect)
    for p, off in zip(params, offsets):  # param.grad must now alias the bucket and hold the averaged gradient
        ok = ok and p.grad.data_ptr() == flat.data_ptr() + 4 * off and torch.allclose(p.grad.flatten(),
                                                                                    expect[off:off + p.numel()])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
