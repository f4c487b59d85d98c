"""Drop-in replacement for the reference's `dataset/utils.py::radius_graph_pbc`
(/root/reference/dataset/utils.py:57-237) running on the GPU neighbour-list kernels
(cartnet_b200/csrc/nlist.cu). Same signature, same 4-tuple, bit-exact edge set and order.
"""
from __future__ import annotations

import torch

from . import ops


def _pbc_mask(data, pbc):
    """utils.py:67-77: the batch's own `pbc` overrides the argument, mixed settings are an error."""
    pbc = list(pbc)
    if hasattr(data, "pbc") and data.pbc is not None:
        p = torch.atleast_2d(data.pbc)
        for i in range(3):
            if not torch.any(p[:, i]).item():
                pbc[i] = False
            elif torch.all(p[:, i]).item():
                pbc[i] = True
            else:
                raise RuntimeError("Different structures in the batch have different PBC configurations. "
                                   "This is not currently supported.")
    return sum((1 << i) for i in range(3) if pbc[i])


def radius_graph_pbc(data, radius, max_num_neighbors_threshold=None, enforce_max_neighbors_strictly: bool = False,
                     pbc=[True, True, True]):
    """Returns (edge_index [2,E] i64, unit_cell [E,3] f32, dist [E] f32, direction [E,3] f32).
    `data` needs .pos [N,3], .cell [B,3,3], .natoms [B] (and optionally .pbc); tensors must be on a
    CUDA device. Like the reference, a multi-crystal call searches max(rep) cells for every crystal.
    `max_num_neighbors_threshold` applies the reference's kNN cap (degenerate neighbours within 0.01 A^2 of the cut are
    kept together unless `enforce_max_neighbors_strictly`)."""
    out = ops.nlist_build(data.pos, data.cell, data.natoms, float(radius), pbc_mask=_pbc_mask(data, pbc),
                          batch_max_reps=True, want_cart=False)
    ei, uc, dist, direction = out["edge_index"], out["unit_cell"], out["dist"], out["direction"]
    if max_num_neighbors_threshold is not None and max_num_neighbors_threshold > 0:     # utils.py:215-233
        keep, _ = ops.nlist_knn_mask(direction, out["row_ptr"], int(data.pos.shape[0]), int(max_num_neighbors_threshold),
                                     strict=enforce_max_neighbors_strictly)
        ei, uc, dist, direction = ei[:, keep], uc[keep], dist[keep], direction[keep]    # masked_select, like the reference
    return ei, uc, dist, direction


def build_graph(pos, cell, natoms, radius: float = 5.0, pbc=(True, True, True)):
    """Batched graph build for many crystals in ONE launch with per-crystal repeat counts -- what the
    reference's data sets get by calling radius_graph_pbc once per crystal
    (/root/reference/dataset/figshare_dataset.py:64-68) -- plus the callers' post-processing
    (cart_dist, cart_dir) and the int32 CSR views the layer kernels use. Node ids are global."""
    mask = sum((1 << i) for i in range(3) if pbc[i])
    return ops.nlist_build(pos, cell, natoms, float(radius), pbc_mask=mask, batch_max_reps=False,
                           want_cart=True, want_i32=True)
