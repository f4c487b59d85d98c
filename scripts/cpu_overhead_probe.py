#!/usr/bin/env python
"""How long does the HOST need to issue one training step (no syncs), versus the device time of the step?"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import cartnet_b200
from cartnet_b200 import cartnet as CN
from cartnet_b200.ddp import FlatGradAllReduce

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hb = bench.make_host_batch(bench.rank_structures(B, 2, 0, 1, dev), 2, dev)
db = bench.shallow(hb.clone()).to(dev)
torch.manual_seed(0)
model = cartnet_b200.CartNet(256, 64, 4, precision="bf16").to(dev).train()
sync = FlatGradAllReduce(model.parameters(), direct=True)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
CN.get_plan(db)

def step():
    sync.zero()
    pred, true = model(bench.shallow(db))
    loss = torch.nn.functional.l1_loss(pred, true)
    loss.backward()
    opt.step()

for _ in range(5):
    step()
torch.cuda.synchronize()
for phase in ("all",):
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    t_issue = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    print("host issue time %.2f ms/step ; device time %.2f ms/step" % (t_issue * 1e3, e0.elapsed_time(e1) / 10))

torch.cuda.set_sync_debug_mode("warn")      # any remaining host<->device synchronisation inside a step is printed
step()
torch.cuda.set_sync_debug_mode("default")
torch.cuda.synchronize()

# split of the host time
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
