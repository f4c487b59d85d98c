#!/usr/bin/env python
"""Which host synchronisations does a training step incur on each input path? Runs two steps per path under
torch.cuda.set_sync_debug_mode("warn") (after a warm step) and prints the distinct warning sites.
Paths: device-resident batch; DevicePrefetcher over pinned host batches; DeviceDataset.collate."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import cartnet_b200
from cartnet_b200 import DeviceDataset, DevicePrefetcher
dev = torch.device("cuda", 0)
hbs = [bench.make_host_batch(bench.rank_structures("adp", 16, 2 + i, 0, 1, dev), 2 + i, dev) for i in range(2)]
torch.manual_seed(0)
model = cartnet_b200.CartNet(256, 64, 4, precision="bf16x3").to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)

def step(b):
    opt.zero_grad(set_to_none=True)
    pred, true = model(b)
    loss = cartnet_b200.compute_loss(pred, true)[0]
    loss.backward()
    opt.step()
    return loss

def run(name, make_batches):
    batches = make_batches()
    step(next(batches))                      # warm
    torch.cuda.synchronize()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        torch.cuda.set_sync_debug_mode("warn")
        try:
            for _ in range(2):
                step(next(batches))
        finally:
            torch.cuda.set_sync_debug_mode("default")
    sites = sorted({"%s:%d %s" % (os.path.relpath(x.filename, ROOT) if x.filename.startswith(ROOT) else os.path.basename(x.filename), x.lineno, str(x.message)[:60]) for x in w if "ynchroniz" in str(x.message)})
    print("%-28s %d synchronising calls in 2 steps" % (name, len([x for x in w if "ynchroniz" in str(x.message)])))
    for s in sites:
        print("     ", s)

run("device-resident batches", lambda: iter([bench.shallow(hbs[i % 2]).to(dev) for i in range(3)]))
run("DevicePrefetcher (pinned)", lambda: iter(DevicePrefetcher((hbs[i % 2] for i in range(3)), dev)))
ds = DeviceDataset.from_batch(hbs[0], dev)
run("DeviceDataset.collate", lambda: (ds.collate(list(range(16))) for _ in range(3)))
