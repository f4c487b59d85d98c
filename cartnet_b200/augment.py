"""SO(3) augmentation and Monte-Carlo rotation on the device -- SURVEY.md §8(f)4.

The reference rotates one sample at a time on the host with `roma.utils.random_rotmat`
(/root/reference/dataset/datasetADP.py:33-39: y -> R^T y R, cart_dir -> cart_dir R, cell -> cell R) and, in its
Monte-Carlo evaluation, a whole batch with a single rotation (/root/reference/main.py:94-97). Here one rotation per
CRYSTAL of a collated batch is drawn and applied on the GPU without any host synchronisation, so augmentation can sit
between the (device-side) graph build and the model. Edge distances are rotation invariant and are left untouched.
"""
from __future__ import annotations

import torch


def random_rotations(n: int, device, generator: torch.Generator | None = None, dtype=torch.float32) -> torch.Tensor:
    """n rotation matrices [n,3,3], uniform on SO(3) (unit quaternions from normalised Gaussians)."""
    q = torch.randn(n, 4, device=device, dtype=dtype, generator=generator)
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.view(n, 3, 3)


def rotate_batch_(batch, R: torch.Tensor):
    """In place: applies R[g] ([B,3,3], or one [3,3] for the whole batch like main.py:96) to every crystal g of a collated
    batch: cart_dir (per edge, by the crystal of its destination atom), cell, pos (row vectors: v -> v R) and, when
    present, the 3x3 ADP targets y (per non-H atom: y -> R^T y R). Returns the batch."""
    if R.dim() == 2:
        batch.cart_dir = batch.cart_dir @ R
        if getattr(batch, "cell", None) is not None:
            batch.cell = batch.cell @ R
        if getattr(batch, "pos", None) is not None:
            batch.pos = batch.pos @ R
        y = getattr(batch, "y", None)
        if y is not None and y.dim() == 3:
            batch.y = R.transpose(-1, -2) @ y @ R
        return batch
    node_g = batch.batch                                         # crystal of every atom
    edge_g = node_g.index_select(0, batch.edge_index[1])         # crystal of every edge (edges never cross crystals)
    batch.cart_dir = torch.bmm(batch.cart_dir.unsqueeze(1), R.index_select(0, edge_g)).squeeze(1)
    if getattr(batch, "cell", None) is not None:
        batch.cell = torch.bmm(batch.cell, R)
    if getattr(batch, "pos", None) is not None:
        batch.pos = torch.bmm(batch.pos.unsqueeze(1), R.index_select(0, node_g)).squeeze(1)
    y = getattr(batch, "y", None)
    if y is not None and y.dim() == 3:
        idx = getattr(batch, "non_H_index", None)
        atom_g = node_g.index_select(0, idx) if idx is not None else node_g[batch.non_H_mask]
        Ry = R.index_select(0, atom_g)
        batch.y = Ry.transpose(-1, -2) @ y @ Ry
    return batch


def augment_(batch, generator: torch.Generator | None = None):
    """One independent uniform rotation per crystal (the batched equivalent of DatasetADP.augment_data)."""
    B = int(batch.natoms.numel()) if getattr(batch, "natoms", None) is not None else int(batch.cell.shape[0])
    return rotate_batch_(batch, random_rotations(B, batch.cart_dir.device, generator))
