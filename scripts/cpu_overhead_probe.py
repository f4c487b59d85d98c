#!/usr/bin/env python
"""How long does the HOST need to issue one training step (no syncs), versus the device time of the step?
usage: python scripts/cpu_overhead_probe.py [workload=adp_train] [precision=bf16x3]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import cartnet_b200

name = sys.argv[1] if len(sys.argv) > 1 else "adp_train"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
wl = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
hb = bench.make_host_batch(bench.rank_structures(wl["shape"], wl["batch"], wl["seed"], 0, 1, dev), wl["seed"], dev, cholesky=wl["model"].get("cholesky", True))
db = bench.shallow(hb).to(dev)
torch.manual_seed(0)
model = cartnet_b200.CartNet(256, 64, 4, precision=prec, **wl["model"]).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)

def step():
    opt.zero_grad(set_to_none=True)
    pred, true = model(bench.shallow(db))
    loss = cartnet_b200.compute_loss(pred, true)[0]
    loss.backward()
    opt.step()

for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
t_issue = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
print("%s %s: host issue %.2f ms/step, device %.2f ms/step (E = %d)" % (name, prec, t_issue * 1e3, e0.elapsed_time(e1) / 10, db.edge_index.shape[1]))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
