"""TEST INFRASTRUCTURE ONLY -- seeded inputs and weights shared by tests/golden/make_golden.py,
tests/ and the smoke / cpu_baseline legs. Graphs here are built with the ORACLE graph
builder (never the CUDA one) so that model parity does not depend on graph parity."""
from __future__ import annotations

import zlib

import numpy as np
import torch

from cartnet_b200 import synthetic
from cartnet_b200.batch import CrystalBatch, collate
from oracle import cartnet_oracle as O


def make_oracle_batch(shape: str, count: int, seed: int, sizes=None, radius: float = 5.0,
                      cholesky: bool = True, temperature: bool = True) -> CrystalBatch:
    structs = synthetic.make_structures(shape, count, seed, sizes=sizes)
    rng = np.random.default_rng(seed + 7919)
    items = []
    for s in structs:
        n = len(s["z"])
        ei, _, _, direction = O.radius_graph_pbc_oracle(s["pos"], s["cell"][None], [n], radius)
        cd, cdir = O.edge_vectors(torch.from_numpy(direction))
        it = {"x": torch.from_numpy(s["z"]), "pos": torch.from_numpy(s["pos"]),
              "cell": torch.from_numpy(s["cell"]), "edge_index": torch.from_numpy(ei),
              "cart_dist": cd, "cart_dir": cdir}
        mask = s["z"] != 1
        it["non_H_mask"] = torch.from_numpy(mask)
        if temperature:
            it["temperature"] = torch.tensor(s["temperature"])
        if cholesky:
            it["y"] = torch.from_numpy(synthetic.adp_targets(int(mask.sum()), rng))
        else:
            it["y"] = torch.tensor(np.float32(rng.standard_normal()))
        items.append(it)
    return collate(items)


def make_state_dict(template: dict, seed: int) -> dict:
    """Deterministic, torch-version-independent weights for every floating tensor in a
    CartNet state dict: U(+-1/sqrt(fan_in)) for >=2-D, small perturbations for BN affine /
    running stats so that eval mode is exercised with non-trivial statistics."""
    out = {}
    for k in sorted(template.keys()):
        v = template[k]
        if not v.dtype.is_floating_point:
            out[k] = v.clone()
            continue
        rng = np.random.default_rng([seed, zlib.crc32(k.encode())])
        shape = tuple(v.shape)
        if k.endswith("rbf.means") or k.endswith("rbf.betas"):
            out[k] = v.clone()
        elif k.endswith("running_var"):
            out[k] = torch.from_numpy(rng.uniform(0.5, 1.5, size=shape).astype(np.float32))
        elif k.endswith("running_mean"):
            out[k] = torch.from_numpy(rng.normal(0, 0.1, size=shape).astype(np.float32))
        elif ".norm" in k and k.endswith("weight"):
            out[k] = torch.from_numpy(rng.uniform(0.8, 1.2, size=shape).astype(np.float32))
        elif ".norm" in k and k.endswith("bias"):
            out[k] = torch.from_numpy(rng.normal(0, 0.1, size=shape).astype(np.float32))
        else:
            fan_in = shape[1] if len(shape) >= 2 else shape[0]
            if k.endswith("bias") and len(shape) == 1:
                fan_in = max(shape[0], 16)
            b = 1.0 / np.sqrt(fan_in)
            out[k] = torch.from_numpy(rng.uniform(-b, b, size=shape).astype(np.float32))
    return out


def subsample_rows(t: torch.Tensor, k: int = 64) -> np.ndarray:
    """Rows 0, step, 2*step, ... (at most k) -- keeps fixtures small."""
    n = t.shape[0]
    if n == 0:
        return t.detach().numpy()
    step = max(1, n // k)
    return t.detach()[::step][:k].contiguous().numpy()
