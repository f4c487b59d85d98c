#!/usr/bin/env python
"""Graph-build throughput of the neighbour-list kernels (count + scan + fill, incl. the size read-back)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cartnet_b200 import ops, synthetic

def build_graph(pos, cell, nat, radius, cells=True):
    return ops.nlist_build(pos, cell, nat, radius, batch_max_reps=False, want_cart=True, want_i32=True, cells=cells)

def run(shape, count, seed, sizes=None, reps=5, cells=True):
    structs = synthetic.make_structures(shape, count, seed, sizes=sizes)
    pos = torch.from_numpy(np.concatenate([s["pos"] for s in structs])).cuda()
    cell = torch.from_numpy(np.stack([s["cell"] for s in structs])).cuda()
    nat = torch.tensor([len(s["z"]) for s in structs]).cuda()
    for _ in range(2):
        out = build_graph(pos, cell, nat, 5.0, cells)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = build_graph(pos, cell, nat, 5.0, cells)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    E = out["edge_index"].shape[1]
    print("%-9s " % ("cells" if cells else "all-pairs") + "%-10s crystals %4d atoms %7d edges %9d : %7.3f ms  -> %9.0f graphs/s %6.1f M edges/s" % (
        shape, count, pos.shape[0], E, dt * 1e3, count / dt, E / dt / 1e6))

run("adp", 64, 2)
run("adp", 1024, 2)
run("jarvis", 4096, 3)
run("mp", 4096, 4)
for cells in (False, True):
    run("supercell", 1, 5, cells=cells)
    run("supercell", 8, 5, cells=cells)
    run("supercell", 1, 6, sizes=np.array([20000]), cells=cells)
    run("supercell", 1, 7, sizes=np.array([100000]), cells=cells, reps=2)
