#!/usr/bin/env python
"""Prints the parity of every precision mode against the reference's golden vectors (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import cartnet_b200  # noqa: E402
from oracle import fixtures  # noqa: E402

gm = np.load(os.path.join(ROOT, "tests", "golden", "model.npz"))
for name in common.MODEL_CASES:
    shape, sizes, seed, kw, lrad = common.MODEL_CASES[name]
    batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes), cholesky=kw["cholesky"],
                                        temperature=kw["temperature"]).to("cuda")
    for precision in ("fp32", "bf16x3", "tf32", "bf16"):
        torch.manual_seed(0)
        model = cartnet_b200.CartNet(common.DIM_IN, common.DIM_RBF, common.NUM_LAYERS, radius=lrad, precision=precision, **kw)
        model.load_state_dict(fixtures.make_state_dict(model.state_dict(), seed))
        model.cuda()
        try:
            res = common.run_train_step(model, batch0)
        except Exception as ex:  # noqa: BLE001
            print("%-16s %-5s FAILED: %s" % (name, precision, str(ex)[:200]))
            continue
        pre = name + "/"
        e_pred = common.rel_err(res["pred"], torch.from_numpy(gm[pre + "pred"]))
        e_eval = common.rel_err(res["pred_eval"], torch.from_numpy(gm[pre + "pred_eval"]))
        e_e = common.rel_err(torch.from_numpy(fixtures.subsample_rows(res["e"].cpu(), 32)), torch.from_numpy(gm[pre + "e_out_rows"]))
        gk = [k[len(pre + "grad/"):] for k in gm.files if k.startswith(pre + "grad/")]
        worst, wk = 0.0, ""
        for k in gk:
            ref = torch.from_numpy(gm[pre + "grad/" + k])
            g = res["grads"][k].cpu()
            g2 = g.reshape(g.shape[0], -1)
            mine = torch.from_numpy(fixtures.subsample_rows(g2, 8)) if g2.numel() > 4096 else g
            if float(ref.abs().max()) < 1e-4 * max(float(gm[pre + "gradnorm/" + q]) for q in gk):
                continue
            r = common.rel_err(mine, ref)
            if not (r <= worst):      # NaN must surface as the worst entry
                worst, wk = r, k
        print("%-16s %-6s train pred %.2e | eval pred %.2e | edge_attr %.2e | worst grad %.2e (%s)" % (
            name, precision, e_pred, e_eval, e_e, worst, wk))
