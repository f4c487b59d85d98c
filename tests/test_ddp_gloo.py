"""CPU, world_size 2, gloo: the N>1 host logic -- crystal sharding by edge count and the single
flat-gradient all-reduce -- on a small stand-in module (the CartNet kernels themselves need a GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cartnet_b200.ddp import FlatGradAllReduce, broadcast_module, shard_by_edges


def test_shard_by_edges_is_balanced_and_deterministic():
    counts = [1000, 10, 20, 990, 500, 480, 30, 5, 700, 300]
    parts = shard_by_edges(counts, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(counts)))
    loads = [sum(counts[i] for i in p) for p in parts]
    assert max(loads) <= 1.05 * (sum(counts) / 4) + max(counts) * 0  # LPT is within 4/3 of optimal; here nearly exact
    assert max(loads) - min(loads) <= 120
    assert parts == shard_by_edges(counts, 4)
    assert shard_by_edges([5, 5], 4) == [[0], [1], [], []]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different init per rank ...
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.SiLU(), torch.nn.Linear(16, 3))
    broadcast_module(model, 0)                          # ... made identical here
    sync = FlatGradAllReduce(model.parameters())
    torch.manual_seed(7)
    data = torch.randn(2 * 12, 8)
    target = torch.randn(2 * 12, 3)
    xs, ys = data[rank * 12:(rank + 1) * 12], target[rank * 12:(rank + 1) * 12]    # each rank: its own shard
    sync.zero()
    loss = torch.nn.functional.l1_loss(model(xs), ys)
    loss.backward()
    assert all(p.grad.data_ptr() >= sync.flat.data_ptr() for p in model.parameters())   # grads landed in the flat buffer
    local = sync.flat.clone()
    sync.allreduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    ok = torch.allclose(sync.flat, expect, atol=1e-7) and not torch.allclose(local, expect)
    w0 = [torch.zeros_like(model[0].weight) for _ in range(world)]
    dist.all_gather(w0, model[0].weight.data)
    ok = ok and torch.equal(w0[0], w0[1])
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
