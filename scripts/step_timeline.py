#!/usr/bin/env python
"""Where does one training step go on the device? Kineto (torch.profiler) trace of a few steps: every kernel
(ours AND torch's), grouped by name, plus the GPU-idle share of the step. Not a bench number (profiler on)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import cartnet_b200
from cartnet_b200 import cartnet as CN
from cartnet_b200.ddp import FlatGradAllReduce
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STEPS = 4
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        step()
    torch.cuda.synchronize()

evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print("span %.3f ms/step, kernel-busy %.3f ms/step, idle %.3f ms/step, %d device events/step" %
      ((t1 - t0) / STEPS / 1e3, busy / STEPS / 1e3, (t1 - t0 - busy) / STEPS / 1e3, len(evs) // STEPS))
agg = {}
for e in evs:
    d = agg.setdefault(e.name[:90], [0.0, 0])
    d[0] += e.time_range.end - e.time_range.start
    d[1] += 1
ours = ("tc_nt", "tc_tn", "colreduce", "edge_", "segment_sum", "node_update", "sgemm", "pack_weights", "bias_grad", "cast_",
        "finalize", "nlist", "gate_", "dsilu", "rbf", "colstats", "splitk")
tot_ours = sum(v[0] for k, v in agg.items() if any(o in k for o in ours))
print("ours %.3f ms/step ; other (torch / memcpy / memset) %.3f ms/step" % (tot_ours / STEPS / 1e3, (busy - tot_ours) / STEPS / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    print("%9.1f us/step  x%5.1f  %s" % (v[0] / STEPS, v[1] / STEPS, k))
# gaps: idle time preceding each kernel, attributed to that kernel's name
gaps = {}
for a, b in zip(evs[:-1], evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 0:
        d = gaps.setdefault(b.name[:70], [0.0, 0])
        d[0] += g
        d[1] += 1
print("--- largest idle gaps (before kernel)")
for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%9.1f us/step  x%5.1f  %s" % (v[0] / STEPS, v[1] / STEPS, k))
