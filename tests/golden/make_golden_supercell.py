"""Generates tests/golden/supercell5000.npz: BASELINE configs[4] at its full size (one 5000-atom crystal, ~275 k
edges) from the CPU ORACLE (chunked graph build + OracleCartNet eval forward). Too slow for a test (minutes, GBs), so it
is run once in the authoring container:

    python tests/golden/make_golden_supercell.py

Stored: sha256 of edge_index / unit_cell / direction (graph parity is bit-exact), E, the eval-mode prediction of the
4-layer model with fixtures.make_state_dict(seed 5) weights, and 64 sampled rows of the final edge features."""
import hashlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cartnet_oracle as O  # noqa: E402
from oracle import fixtures  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    seed, n = 5, 5000
    t0 = time.time()
    batch = fixtures.make_oracle_batch("supercell", 1, seed, sizes=np.array([n]))
    print("oracle graph: %d edges in %.1f s" % (batch.num_edges, time.time() - t0))
    from cartnet_b200 import synthetic
    s = synthetic.make_structures("supercell", 1, seed, sizes=np.array([n]))[0]
    ei, uc, dist, direction = O.radius_graph_pbc_oracle(s["pos"], s["cell"][None], [n], 5.0)
    assert np.array_equal(ei, batch.edge_index.numpy())
    kw = dict(invariant=False, temperature=True, use_envelope=True, atom_types=True, cholesky=True)
    torch.manual_seed(0)
    orc = O.OracleCartNet(256, 64, 4, **kw)
    orc.load_state_dict(fixtures.make_state_dict(orc.state_dict(), seed))
    orc.eval()
    t0 = time.time()
    with torch.no_grad():
        b = batch.clone()
        pred, _ = orc(b)
    print("oracle eval forward %.1f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "supercell5000.npz"),
                        num_edges=np.int64(batch.num_edges), edge_index_sha=sha(ei), unit_cell_sha=sha(uc), direction_sha=sha(direction),
                        row_count=np.bincount(ei[1], minlength=n).astype(np.int32),
                        pred=pred.numpy(), e_rows=fixtures.subsample_rows(b.edge_attr, 64), x_rows=fixtures.subsample_rows(b.x, 64))
    print("wrote supercell5000.npz")


if __name__ == "__main__":
    main()
