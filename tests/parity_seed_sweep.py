#!/usr/bin/env python
"""fp32-mode parity against the CPU oracle over several seeds / batch shapes (training step + eval forward)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
import cartnet_b200
from oracle import cartnet_oracle as O, fixtures

torch.set_num_threads(os.cpu_count())
kw = dict(invariant=False, temperature=True, use_envelope=True, atom_types=True, cholesky=True)
worst = dict(pred=0, eval=0, e=0, x=0, grad=0)
for seed in range(100, 108):
    shape = ["adp", "mp", "jarvis"][seed % 3]
    count = [3, 6, 10][seed % 3]
    batch = fixtures.make_oracle_batch(shape, count, seed)
    torch.manual_seed(seed)
    orc = O.OracleCartNet(256, 64, 4, **kw)
    sd = fixtures.make_state_dict(orc.state_dict(), seed)
    orc.load_state_dict(sd)
    model = cartnet_b200.CartNet(256, 64, 4, precision="fp32", **kw)
    model.load_state_dict(sd); model.cuda()
    ref = common.run_train_step(orc, batch)
    got = common.run_train_step(model, batch.clone().to("cuda"))
    scale = max(float(v.abs().max()) for v in ref["grads"].values())
    gerr = max(float((got["grads"][k].cpu() - g).abs().max()) / (float(g.abs().max()) + 1e-3 * scale) for k, g in ref["grads"].items())
    errs = dict(pred=common.rel_err(got["pred"], ref["pred"]), eval=common.rel_err(got["pred_eval"], ref["pred_eval"]),
                e=common.rel_err(got["e"], ref["e"]), x=common.rel_err(got["x"], ref["x"]), grad=gerr)
    for k in worst:
        worst[k] = max(worst[k], errs[k])
    print("seed %d %-6s N=%4d E=%6d " % (seed, shape, batch.num_nodes, batch.num_edges) + " ".join("%s %.2e" % kv for kv in errs.items()), flush=True)
print("worst", " ".join("%s %.2e" % kv for kv in worst.items()))
