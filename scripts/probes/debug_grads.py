#!/usr/bin/env python
"""Per-layer input gradients (x, edge_attr) of one training step, native layer orchestration vs the Python composition:
NaN counts and max |difference|. A debugging aid for the bf16x3 path (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import cartnet_b200  # noqa: E402
from cartnet_b200 import functional as CF  # noqa: E402
from oracle import fixtures  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
shape, sizes, seed, kw, lrad = common.MODEL_CASES["adp"]
batch0 = fixtures.make_oracle_batch(shape, len(sizes), seed, sizes=np.array(sizes)).to("cuda")
res = {}
for native, pr in ((True, prec), (False, prec), (True, "fp32")):
    CF.USE_NATIVE_LAYER = native
    torch.manual_seed(0)
    model = cartnet_b200.CartNet(common.DIM_IN, common.DIM_RBF, common.NUM_LAYERS, radius=lrad, precision=pr, **kw)
    model.load_state_dict(fixtures.make_state_dict(model.state_dict(), seed))
    model.cuda().train()
    cap = {}
    for i, layer in enumerate(model.layers):
        def pre(mod, args, i=i):
            b = args[0]
            for nm in ("x", "edge_attr"):
                t = getattr(b, nm)
                if t.requires_grad:
                    t.register_hook(lambda g, key=(i, nm): cap.__setitem__(key, g.detach().clone()))
        layer.register_forward_pre_hook(pre)
    out = model(batch0.clone())
    pred = out[0] if isinstance(out, tuple) else out
    pred.float().abs().sum().backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    res[len(res)] = (native, cap, grads)
for idx, (native, cap, grads) in res.items():
    print("run", idx, "native" if native else "python")
    for key in sorted(cap):
        g = cap[key]
        print("   layer %d d%-9s nan %7d  max %.4e" % (key[0], key[1], int(torch.isnan(g).sum()), float(torch.nan_to_num(g).abs().max())))
    bad = [k for k, g in grads.items() if torch.isnan(g).any()]
    print("   NaN parameter gradients:", bad)
(_, c0, g0), (_, c1, g1), (_, c2, g2) = res[0], res[1], res[2]
for key in sorted(c0):
    print("layer %d d%-9s max|native - fp32| %.3e   max|python - fp32| %.3e" % (
        key[0], key[1], float(torch.nan_to_num(c0[key] - c2[key]).abs().max()), float(torch.nan_to_num(c1[key] - c2[key]).abs().max())))
for k in g2:
    a, b = float(torch.nan_to_num(g0[k] - g2[k]).abs().max()), float(torch.nan_to_num(g1[k] - g2[k]).abs().max())
    m = float(g2[k].abs().max())
    if max(a, b) > 5e-3 * m + 1e-6:
        print("  %-40s |fp32| %.3e  native err %.3e  python err %.3e" % (k, m, a, b))
