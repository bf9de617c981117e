"""CPU, world_size 2 over gloo: the flat gradient bucket, the single all-reduce and the replica-consistent Adam
update of straps_b200.parallel (the N>1 path of BASELINE config 4; NCCL replaces gloo on the GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, PKG)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from straps_b200.parallel import DataParallelAdam, shard
    torch.manual_seed(100 + rank)                       # replicas start DIFFERENT on purpose: broadcast must fix it
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    opt = DataParallelAdam(model.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(7)
    X = torch.randn(8, 6, generator=g)
    Y = torch.randn(8, 3, generator=g)
    xs, ys = shard(X, rank, world), shard(Y, rank, world)
    for _ in range(3):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(xs), ys, reduction='sum') / X.shape[0]
        loss.backward()
        opt.step()
    flat = opt.bucket.params.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        torch.save({'params': gathered, 'grads': opt.bucket.grads.clone(), 'numel': opt.bucket.numel}, out)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_adam_world2(tmp_path):
    out = str(tmp_path / 'r0.pt')
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    # every rank ends with bit-identical parameters
    assert torch.equal(res['params'][0], res['params'][1])
    # and they equal single-process training on the full batch with torch.optim.Adam from rank 0's initial weights
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(7)
    X = torch.randn(8, 6, generator=g)
    Y = torch.randn(8, 3, generator=g)
    for _ in range(3):
        opt.zero_grad()
        # each rank divides its shard's summed loss by the GLOBAL batch, the all-reduce sums, Adam divides by world:
        # that is mean-over-ranks of per-rank gradients = the full-batch gradient / world ... so compare like for like
        loss = torch.nn.functional.mse_loss(model(X), Y, reduction='sum') / X.shape[0] / 2
        loss.backward()
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert res['numel'] == ref.numel()
    assert torch.allclose(res['params'][0], ref, rtol=1e-5, atol=1e-6)


def test_bucket_views_and_shard():
    import sys
    sys.path.insert(0, PKG)
    from straps_b200.parallel import FlatBucket, shard
    lin = torch.nn.Linear(4, 3)
    w0 = lin.weight.detach().clone()
    b = FlatBucket(lin.parameters())
    assert b.numel == 15 and torch.equal(lin.weight.detach(), w0)
    lin(torch.ones(2, 4)).sum().backward()
    assert torch.equal(b.grads[:12].view(3, 4), lin.weight.grad) and float(b.grads.abs().sum()) > 0
    # zero_grad drops every .grad (the next backward WRITES its gradients); gather() puts whatever autograd produced into the bucket
    b.zero_grad()
    assert lin.weight.grad is None and float(b.grads.abs().sum()) == 0
    lin(torch.ones(2, 4)).sum().backward()
    assert lin.weight.grad.data_ptr() != b.grads.data_ptr()          # autograd allocated its own tensor ...
    b.gather()
    assert lin.weight.grad.data_ptr() == b.grads.data_ptr()          # ... gather() copied it into the slot and re-homed .grad
    assert torch.equal(b.grads[:12].view(3, 4), torch.ones(3, 4) * 2) and torch.equal(b.grads[12:], torch.ones(3) * 2)
    b.zero_grad()
    b.gather()                                                       # no backward at all: zero gradients, views restored
    assert float(lin.weight.grad.abs().sum()) == 0 and lin.weight.grad.data_ptr() == b.grads.data_ptr()
    x = torch.arange(12).view(6, 2)
    assert shard(x, 1, 3).tolist() == [[4, 5], [6, 7]]
    with pytest.raises(ValueError):
        shard(x, 0, 4)
