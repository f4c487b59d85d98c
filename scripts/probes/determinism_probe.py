#!/usr/bin/env python
"""Bit-reproducibility of one full-size ADP-64 training step (bench batch): run it three times from the same state and
compare loss, prediction and every gradient bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import cartnet_b200
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
dev = torch.device("cuda", 0)
hb = bench.make_host_batch(bench.rank_structures("adp", 64, 2, 0, 1, dev), 2, dev)
runs = []
for r in range(3):
    torch.manual_seed(0)
    model = cartnet_b200.CartNet(256, 64, 4, precision=prec).to(dev).train()
    pred, true = model(bench.shallow(hb).to(dev))
    loss = cartnet_b200.compute_loss(pred, true)[0]
    loss.backward()
    runs.append((loss.detach().clone(), pred.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
for r in (1, 2):
    bad = [k for k in runs[0][2] if not torch.equal(runs[0][2][k], runs[r][2][k])]
    print("run", r, "loss equal", bool(torch.equal(runs[0][0], runs[r][0])), "pred equal", bool(torch.equal(runs[0][1], runs[r][1])), "gradients that differ:", bad)
