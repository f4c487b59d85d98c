#!/usr/bin/env python
"""Why does the blocking end-to-end loop of bench.py vary from box to box? Per-step host wall times of the e2e loop
(prefetcher + step + loss.item()), the device-resident loop and the allocator's cudaMalloc count, in one process."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import cartnet_b200
from cartnet_b200 import DevicePrefetcher, cartnet as CN
from cartnet_b200.ddp import FlatGradAllReduce

dev = torch.device("cuda:0")
nb = 2
host = [bench.make_host_batch(bench.rank_structures(64, 2 + i, 0, 1, dev), 2 + i, dev) for i in range(nb)]
devb = [bench.shallow(b.clone()).to(dev) for b in host]
torch.manual_seed(0)
model = cartnet_b200.CartNet(256, 64, 4, precision="bf16").to(dev).train()
sync = FlatGradAllReduce(model.parameters(), direct=True)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
for b in devb:
    CN.get_plan(b)


def step(b):
    sync.zero()
    pred, true = model(b)
    loss = torch.nn.functional.l1_loss(pred, true)
    loss.backward()
    opt.step()
    return loss


def mallocs():
    return torch.cuda.memory_stats(dev)["num_device_alloc"]


for i in range(5):
    step(bench.shallow(devb[i % nb]))
torch.cuda.synchronize()
for rep in range(3):
    m0 = mallocs()
    t0 = time.perf_counter()
    for i in range(10):
        step(bench.shallow(devb[i % nb]))
    torch.cuda.synchronize()
    print("resident loop: %.2f ms/step, cudaMallocs %d" % ((time.perf_counter() - t0) * 100, mallocs() - m0))
for rep in range(4):
    feed = DevicePrefetcher((host[i % nb] for i in range(12)), dev)
    ts = []
    m0 = mallocs()
    torch.cuda.synchronize()
    for i in range(12):
        t0 = time.perf_counter()
        b = next(feed)
        t1 = time.perf_counter()
        loss = step(b)
        t2 = time.perf_counter()
        v = float(loss.item())
        t3 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1, t3 - t2))
    a = np.array(ts[2:]) * 1e3
    print("e2e loop: %.2f ms/step (prefetch %.2f, issue %.2f, wait %.2f; worst step %.2f), cudaMallocs %d, reserved %.1f GB" % (
        a.sum(1).mean(), a[:, 0].mean(), a[:, 1].mean(), a[:, 2].mean(), a.sum(1).max(), mallocs() - m0,
        torch.cuda.memory_reserved(dev) / 1e9))
