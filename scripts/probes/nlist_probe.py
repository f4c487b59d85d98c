"""ncu target: the cell-list neighbour kernels on one large crystal. usage: python scripts/probes/nlist_probe.py [atoms] [builds]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cartnet_b200 import ops, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
s = synthetic.make_structures("supercell", 1, 7, sizes=np.array([n]))[0]
pos, cell, nat = torch.from_numpy(s["pos"]).cuda(), torch.from_numpy(s["cell"][None]).cuda(), torch.tensor([n]).cuda()
for _ in range(reps):
    out = ops.nlist_build(pos, cell, nat, 5.0, batch_max_reps=False, want_cart=True, want_i32=True, cells=True)
torch.cuda.synchronize()
print("atoms", n, "edges", out["edge_index"].shape[1])
