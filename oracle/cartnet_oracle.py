"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the CartNet hot path.

This file is the parity oracle for the CUDA path in cartnet_b200/. It is imported
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg; the product package never imports it (and fails loudly when the CUDA library is
missing instead of falling back to anything here).

Every function cites the reference lines it restates (paths relative to
/root/reference). Pinning: the reference ships no tests or golden vectors
("parity unpinned" upstream), so this oracle is pinned against outputs
of the UNMODIFIED reference files run in the authoring container through
oracle/ref_loader.py; the vectors live in tests/golden/ and were produced by
tests/golden/make_golden.py.

Two parts:
  * graph build  -- numpy, explicit fp32, no FMA, chunked over destination rows so it
                    also runs for n >~ 2000 atoms where the reference's O(n^2 C)
                    temporaries no longer fit (dataset/utils.py:57-237).
  * model        -- plain torch (CPU, any dtype) restatement of Encoder / CartNet_layer /
                    heads with the reference's op order (cat -> Linear -> ...), so autograd
                    supplies the backward oracle (models/cartnet.py:14-327).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

f32 = np.float32


# ----------------------------------------------------------------------------------------
# graph build  (dataset/utils.py:57-237)
# ----------------------------------------------------------------------------------------
def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1],
                     a[2] * b[0] - a[0] * b[2],
                     a[0] * b[1] - a[1] * b[0]], dtype=f32)


def _norm3(v):
    v = v.astype(f32)
    return f32(np.sqrt(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2])))


def cell_repeats(cell, radius, pbc=(True, True, True)):
    """rep_k = ceil(radius * ||(a_i x a_j) / V||_2)   (dataset/utils.py:135-156).
    The division by the signed volume is applied per component BEFORE the norm."""
    cell = np.asarray(cell, dtype=f32).reshape(3, 3)
    c23 = _cross(cell[1], cell[2])
    vol = f32(f32(f32(cell[0][0] * c23[0]) + f32(cell[0][1] * c23[1])) + f32(cell[0][2] * c23[2]))
    reps = []
    crosses = (c23, _cross(cell[2], cell[0]), _cross(cell[0], cell[1]))
    for k in range(3):
        if pbc[k]:
            inv = _norm3((crosses[k] / vol).astype(f32))
            reps.append(int(np.ceil(f32(f32(radius) * inv))))
        else:
            reps.append(0)
    return tuple(reps)


def offset_sum_mode(num_cells: int) -> int:
    """Summation order of pbc_offsets = bmm(cell^T, unit_cell) (dataset/utils.py:181-182)
    as executed by ATen on the authoring host (probed): products are rounded
    separately (no FMA); C < 45 -> (t0+t1)+t2 ; C >= 45 (9C >= 400, MKL sgemm) -> (t0+t2)+t1."""
    return 0 if num_cells < 45 else 1


def cell_offsets(cell, reps, mode=None):
    """unit_cell [C,3] f32 in cartesian_prod order (a1 slowest, dataset/utils.py:166-170)
    and the Cartesian offsets [C,3] f32 with the reference's rounding."""
    cell = np.asarray(cell, dtype=f32).reshape(3, 3)
    r1, r2, r3 = reps
    u = np.stack(np.meshgrid(np.arange(-r1, r1 + 1), np.arange(-r2, r2 + 1),
                             np.arange(-r3, r3 + 1), indexing="ij"), axis=-1).reshape(-1, 3).astype(f32)
    if mode is None:
        mode = offset_sum_mode(len(u))
    t = [(u[:, k:k + 1] * cell[k][None, :]).astype(f32) for k in range(3)]
    if mode == 0:
        off = ((t[0] + t[1]).astype(f32) + t[2]).astype(f32)
    else:
        off = ((t[0] + t[2]).astype(f32) + t[1]).astype(f32)
    return u, off


def radius_graph_pbc_oracle(pos, cell, natoms, radius, pbc=(True, True, True), chunk_rows=64,
                            reps=None, max_num_neighbors_threshold=None, enforce_max_neighbors_strictly=False):
    """Restates radius_graph_pbc (dataset/utils.py:57-237) with max_num_neighbors_threshold=None.

    pos [N,3] f32, cell [B,3,3] f32 (rows = lattice vectors), natoms [B] int.
    Returns edge_index [2,E] i64 (row0 = src j, row1 = dst i), unit_cell [E,3] f32,
    sqrt(d2) [E] f32, direction [E,3] f32 -- ordered by (dst, src, cell) as the flat
    (index1, index2, cell) enumeration of :116-123,188-191 implies.
    Like the reference (:163) a multi-crystal call uses the max rep over the batch.
    """
    pos = np.ascontiguousarray(pos, dtype=f32)
    cell = np.asarray(cell, dtype=f32).reshape(-1, 3, 3)
    natoms = np.asarray(natoms, dtype=np.int64).reshape(-1)
    r2 = f32(radius * radius)  # python double product cast to the tensor dtype by torch.le, :202
    if reps is None:
        per = [cell_repeats(cell[b], radius, pbc) for b in range(len(natoms))]
        reps = tuple(max(p[k] for p in per) for k in range(3))
    starts = np.concatenate([[0], np.cumsum(natoms)])
    src_l, dst_l, uc_l, d2_l, dir_l = [], [], [], [], []
    for b in range(len(natoms)):
        n = int(natoms[b])
        base = int(starts[b])
        p = pos[base:base + n]
        u, off = cell_offsets(cell[b], reps)
        C = len(u)
        for r0 in range(0, n, chunk_rows):
            r1 = min(n, r0 + chunk_rows)
            p1 = p[r0:r1][:, None, None, :]                                    # [R,1,1,3]
            p2 = (p[None, :, None, :] + off[None, None, :, :]).astype(f32)      # [1,n,C,3]  :193
            d = (p1 - p2).astype(f32)                                          # [R,n,C,3]  :196
            sq = (d * d).astype(f32)
            d2 = ((sq[..., 0] + sq[..., 1]).astype(f32) + sq[..., 2]).astype(f32)  # :197
            mask = (d2 <= r2) & (d2 > f32(0.0001))                            # :202-205
            ri, si, ci = np.nonzero(mask)                                      # row-major == (i1,i2,c)
            dst_l.append(ri.astype(np.int64) + r0 + base)
            src_l.append(si.astype(np.int64) + base)
            uc_l.append(u[ci])
            d2_l.append(d2[ri, si, ci])
            dir_l.append(d[ri, si, ci])
    if not dst_l:
        z = np.zeros((0,), np.int64)
        return np.stack([z, z]), np.zeros((0, 3), f32), np.zeros((0,), f32), np.zeros((0, 3), f32)
    src = np.concatenate(src_l)
    dst = np.concatenate(dst_l)
    uc, d2, direc = np.concatenate(uc_l).astype(f32), np.concatenate(d2_l).astype(f32), np.concatenate(dir_l).astype(f32)
    if max_num_neighbors_threshold is not None:                                    # :215-233
        keep = max_neighbors_mask_oracle(dst, d2, int(natoms.sum()), max_num_neighbors_threshold,
                                         enforce_max_strictly=enforce_max_neighbors_strictly)
        src, dst, uc, d2, direc = src[keep], dst[keep], uc[keep], d2[keep], direc[keep]
    return np.stack([src, dst]), uc, np.sqrt(d2).astype(f32), direc


def max_neighbors_mask_oracle(dst, d2, num_nodes, threshold, degeneracy_tolerance=0.01, enforce_max_strictly=False):
    """Restates get_max_neighbors_mask (dataset/utils.py:240-360) for dst-sorted edges: per atom keep the edges whose
    SQUARED distance is <= (the (threshold+1)-th smallest squared distance of that atom) + tolerance, i.e. degenerate
    neighbours are kept together (:322-329); atoms with <= threshold edges keep everything (their cutoff is inf).
    Strict mode keeps the `threshold` smallest (:316-319; ties broken by edge order here -- torch.sort leaves them
    unspecified). If no atom exceeds the threshold, or threshold <= 0, everything is kept (:283-290)."""
    dst = np.asarray(dst)
    d2 = np.asarray(d2, dtype=f32)
    counts = np.bincount(dst, minlength=num_nodes)
    keep = np.ones(len(dst), dtype=bool)
    if threshold is None or threshold <= 0 or counts.max(initial=0) <= threshold:
        return keep
    starts = np.concatenate([[0], np.cumsum(counts)])
    for n in range(num_nodes):
        lo, hi = starts[n], starts[n + 1]
        if hi - lo <= threshold:
            continue
        row = d2[lo:hi]
        order = np.argsort(row, kind="stable")
        if enforce_max_strictly:
            k = np.zeros(hi - lo, dtype=bool)
            k[order[:threshold]] = True
        else:
            cutoff = f32(row[order[threshold]] + f32(degeneracy_tolerance))
            k = row <= cutoff
        keep[lo:hi] = k
    return keep


def edge_vectors(direction: torch.Tensor):
    """cart_dist = ||v||_2 ; cart_dir = v / max(||v||, 1e-12)  (dataset/figshare_dataset.py:67-68)."""
    return torch.norm(direction, p=2, dim=-1), F.normalize(direction, p=2, dim=-1)


# ----------------------------------------------------------------------------------------
# model  (models/cartnet.py, models/utils.py)
# ----------------------------------------------------------------------------------------
def cosine_cutoff(d: torch.Tensor, upper: float) -> torch.Tensor:
    """models/utils.py:87-91 (cutoff_lower = 0 branch)."""
    c = 0.5 * (torch.cos(d * math.pi / upper) + 1.0)
    return c * (d < upper)


def rbf_params(upper: float, num_rbf: int, dtype=torch.float32):
    """models/utils.py:36-49."""
    start = torch.exp(torch.scalar_tensor(-upper + 0.0, dtype=dtype))
    means = torch.linspace(start, 1, num_rbf, dtype=dtype)
    betas = torch.tensor([(2 / num_rbf * (1 - start)) ** -2] * num_rbf, dtype=dtype)
    return means, betas


def exp_normal_smearing(d, means, betas, upper):
    """models/utils.py:56-61 with cutoff_lower = 0, alpha = 5/upper."""
    alpha = 5.0 / upper
    d = d.unsqueeze(-1)
    return cosine_cutoff(d, upper) * torch.exp(-betas * (torch.exp(alpha * (-d)) - means) ** 2)


class OracleEncoder(nn.Module):
    """models/cartnet.py:75-161."""

    def __init__(self, dim_in, dim_rbf, radius=5.0, invariant=False, temperature=True, atom_types=True):
        super().__init__()
        self.dim_in, self.invariant, self.temperature, self.atom_types = dim_in, invariant, temperature, atom_types
        self.radius = radius
        if atom_types:
            self.embedding = nn.Embedding(119, dim_in * 2)
            nn.init.xavier_uniform_(self.embedding.weight.data)
        elif not temperature:
            self.embedding = nn.Embedding(1, dim_in)
        if temperature:
            self.temperature_proj_atom = nn.Linear(1, dim_in * 2, bias=True)
        elif atom_types:
            self.bias = nn.Parameter(torch.zeros(dim_in * 2))
        if temperature or atom_types:
            self.encoder_atom = nn.Sequential(nn.SiLU(), nn.Linear(dim_in * 2, dim_in), nn.SiLU())
        dim_edge = dim_rbf if invariant else dim_rbf + 3
        self.encoder_edge = nn.Sequential(nn.Linear(dim_edge, dim_in * 2), nn.SiLU(),
                                          nn.Linear(dim_in * 2, dim_in), nn.SiLU())
        self.rbf = nn.Module()
        means, betas = rbf_params(radius, dim_rbf)
        self.rbf.register_buffer("means", means)
        self.rbf.register_buffer("betas", betas)

    def forward(self, batch):
        if self.temperature and self.atom_types:                                   # :144-151
            x = self.embedding(batch.x) + self.temperature_proj_atom(batch.temperature.unsqueeze(-1))[batch.batch]
        elif self.atom_types:
            x = self.embedding(batch.x) + self.bias
        elif self.temperature:
            x = self.temperature_proj_atom(batch.temperature.unsqueeze(-1))[batch.batch]
        else:
            batch.x = self.embedding.weight.repeat(batch.x.shape[0], 1)
        if self.temperature or self.atom_types:                                    # :153-154
            batch.x = self.encoder_atom(x)
        rbf = exp_normal_smearing(batch.cart_dist, self.rbf.means, self.rbf.betas, self.radius)
        if self.invariant:                                                         # :156-159
            batch.edge_attr = self.encoder_edge(rbf)
        else:
            batch.edge_attr = self.encoder_edge(torch.cat([rbf, batch.cart_dir], dim=-1))
        return batch


class OracleLayer(nn.Module):
    """models/cartnet.py:163-274 (PyG propagate unrolled: _i = edge_index[1], _j = edge_index[0])."""

    def __init__(self, dim_in, use_envelope=True, radius=5.0):
        super().__init__()
        self.MLP_aggr = nn.Sequential(nn.Linear(dim_in * 3, dim_in), nn.SiLU(), nn.Linear(dim_in, dim_in))
        self.MLP_gate = nn.Sequential(nn.Linear(dim_in * 3, dim_in), nn.SiLU(), nn.Linear(dim_in, dim_in))
        self.norm = nn.BatchNorm1d(dim_in)
        self.norm2 = nn.BatchNorm1d(dim_in)
        self.use_envelope, self.radius = use_envelope, radius

    def forward(self, batch):
        x, e, ei, dist = batch.x, batch.edge_attr, batch.edge_index, batch.cart_dist
        xi, xj = x.index_select(0, ei[1]), x.index_select(0, ei[0])
        c = torch.cat([xi, xj, e], dim=-1)                                         # :237
        g = torch.sigmoid(self.norm(self.MLP_gate(c)))                             # :237-238
        sig = cosine_cutoff(dist, self.radius).unsqueeze(-1) * g if self.use_envelope else g   # :240-243
        s = self.MLP_aggr(c)                                                       # :256
        m = torch.zeros_like(x).index_add_(0, ei[1], sig * s)                      # :259-260
        batch.x = F.silu(self.norm2(m)) + x                                        # :269,223
        batch.edge_attr = e + sig                                                  # :225
        return batch


class OracleCholeskyHead(nn.Module):
    """models/cartnet.py:276-305."""

    def __init__(self, dim_in):
        super().__init__()
        self.MLP = nn.Sequential(nn.Linear(dim_in, dim_in // 2), nn.SiLU(), nn.Linear(dim_in // 2, 6))

    def forward(self, batch):
        p = self.MLP(batch.x[batch.non_H_mask])
        d = F.softplus(p[:, :3])
        L = torch.zeros(p.size(0), 3, 3, dtype=p.dtype, device=p.device)
        L[:, 0, 0], L[:, 1, 1], L[:, 2, 2] = d[:, 0], d[:, 1], d[:, 2]
        L[:, 0, 1], L[:, 0, 2], L[:, 1, 2] = p[:, 3], p[:, 4], p[:, 5]
        return torch.bmm(L.transpose(1, 2), L), batch.y


class OracleScalarHead(nn.Module):
    """models/cartnet.py:307-327."""

    def __init__(self, dim_in):
        super().__init__()
        self.MLP = nn.Sequential(nn.Linear(dim_in, dim_in // 2), nn.SiLU(), nn.Linear(dim_in // 2, 1))

    def forward(self, batch):
        nb = int(batch.batch.max().item() + 1)
        h = self.MLP(batch.x)
        tot = torch.zeros(nb, 1, dtype=h.dtype).index_add_(0, batch.batch, h)
        cnt = torch.zeros(nb, dtype=h.dtype).index_add_(0, batch.batch, torch.ones_like(batch.batch, dtype=h.dtype))
        batch.x = (tot / cnt.clamp(min=1).unsqueeze(-1)).squeeze(-1)
        return batch.x, batch.y


class OracleCartNet(nn.Module):
    """models/cartnet.py:14-73; identical state-dict keys to the reference."""

    def __init__(self, dim_in, dim_rbf, num_layers, radius=5.0, invariant=False, temperature=True,
                 use_envelope=True, atom_types=True, cholesky=True, layer_radius=5.0):
        super().__init__()
        self.encoder = OracleEncoder(dim_in, dim_rbf, radius, invariant, temperature, atom_types)
        self.layers = nn.Sequential(*[OracleLayer(dim_in, use_envelope, layer_radius) for _ in range(num_layers)])
        self.head = OracleCholeskyHead(dim_in) if cholesky else OracleScalarHead(dim_in)

    def forward(self, batch):
        batch = self.encoder(batch)
        for layer in self.layers:
            batch = layer(batch)
        return self.head(batch)
