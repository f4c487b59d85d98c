"""Device-resident data set and on-device batch assembly -- SURVEY.md 8(f)1.

The reference collates on the host (PyG `DataLoader` -> `Batch.from_data_list`, /root/reference/loader/loader.py:114-124),
ships the batch with `batch.to("cuda:0")` every iteration (train/train.py:169) and rebuilds nothing on the device. Once the
layer runs ~250x faster than eager, that host work is the step. Here the whole data set (graphs included) is uploaded ONCE
as per-field blobs -- sized for HBM: the ADP set, 208 k crystals x ~10.6 k edges x 36 B/edge, is ~80 GB of 180 -- and a
batch is assembled from a list of crystal ids by ONE kernel launch (C-ABI `cartnet_collate`): concatenated node / edge /
target tensors, `edge_index` with cumulative node offsets, the `batch` vector, `non_H_index`, and the int32 CSR views
(dst CSR, src CSR, src permutation) that every layer uses, which are therefore never rebuilt per step. The only host->device
traffic per step is the id list and four offset vectors (a few hundred bytes, pinned, asynchronous); nothing synchronises.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from .batch import CrystalBatch, collate

NODE, EDGE, NONH, GRAPH = 0, 1, 2, 3
COPY, I32_NODE, I32_EDGE, I32_TO_I64_NODE, SLOT_I64 = 0, 1, 2, 3, 4


class DeviceDataset:
    """items: per-crystal dicts / objects as `cartnet_b200.collate` takes them (x, pos, cell, edge_index with LOCAL node
    ids, cart_dist, cart_dir and optionally temperature, non_H_mask, y)."""

    def __init__(self, items: Sequence, device="cuda"):
        self._build(collate(list(items)), device)

    @classmethod
    def from_batch(cls, batch: CrystalBatch, device="cuda") -> "DeviceDataset":
        """From an already collated batch (global node ids in edge_index), e.g. what a PyG InMemoryDataset holds."""
        self = cls.__new__(cls)
        self._build(CrystalBatch(**{k: v for k, v in batch.__dict__.items() if torch.is_tensor(v)}), device)
        return self

    def _build(self, big: CrystalBatch, device):
        self.device = torch.device(device)
        big = big.to(self.device)                                       # one upload of the whole set
        G, N, E = big.num_graphs, big.num_nodes, big.num_edges
        self.num_graphs = G
        natoms = big.natoms.cpu().numpy().astype(np.int64)
        plan = ops.graph_plan(big.edge_index, N)
        if plan.perm_dst is not None:                                   # bring every crystal's edges into dst-sorted order once
            big.edge_index = big.edge_index[:, plan.perm_dst].contiguous()
            big.cart_dist, big.cart_dir = big.cart_dist[plan.perm_dst].contiguous(), big.cart_dir[plan.perm_dst].contiguous()
        node_g = big.batch.to(torch.int32)
        edge_g = node_g[plan.dst32.long()]
        ecount = torch.bincount(edge_g.long(), minlength=G).cpu().numpy().astype(np.int64)      # data-set build time: one sync
        node_ptr = np.concatenate([[0], np.cumsum(natoms)])
        edge_ptr = np.concatenate([[0], np.cumsum(ecount)])
        has_mask = getattr(big, "non_H_mask", None) is not None
        if has_mask:
            nonh_idx = torch.nonzero(big.non_H_mask).squeeze(-1)
            hcount = torch.bincount(node_g[nonh_idx].long(), minlength=G).cpu().numpy().astype(np.int64)
        else:
            nonh_idx, hcount = torch.zeros(0, dtype=torch.int64, device=self.device), np.zeros(G, dtype=np.int64)
        nonh_ptr = np.concatenate([[0], np.cumsum(hcount)])
        graph_ptr = np.arange(G + 1, dtype=np.int64)
        assert max(node_ptr[-1], edge_ptr[-1]) < 2 ** 31, "data set too large for int32 offsets: shard it"
        self.sizes = np.stack([natoms, ecount, hcount, np.ones(G, dtype=np.int64)])          # host copy: batch sizes without a sync
        self.blob_ptr = torch.from_numpy(np.stack([node_ptr, edge_ptr, nonh_ptr, graph_ptr]).astype(np.int32)).to(self.device)
        d_node_ptr = torch.from_numpy(node_ptr.astype(np.int32)).to(self.device)
        d_edge_ptr = torch.from_numpy(edge_ptr.astype(np.int32)).to(self.device)
        # blobs (local ids): what the kernel gathers from
        self.x = big.x.contiguous()
        self.pos = big.pos.contiguous()
        self.cell = big.cell.contiguous()
        self.natoms = big.natoms.contiguous()
        self.mask = big.non_H_mask.contiguous() if has_mask else None
        self.src_local = (plan.src32 - d_node_ptr[edge_g.long()]).contiguous()
        self.dst_local = (plan.dst32 - d_node_ptr[edge_g.long()]).contiguous()
        self.cart_dist, self.cart_dir = big.cart_dist.contiguous(), big.cart_dir.contiguous()
        self.row_start = (plan.row_ptr[:-1] - d_edge_ptr[node_g.long()]).contiguous()           # edges before node n inside its crystal
        self.col_start = (plan.col_ptr[:-1] - d_edge_ptr[node_g.long()]).contiguous()
        # positions [edge_ptr[g], edge_ptr[g+1]) of the src-sorted order belong to crystal g (edges never cross crystals)
        self.perm_src_local = plan.perm_src
        if E > 0:
            pos_g = torch.repeat_interleave(torch.arange(G, device=self.device), torch.from_numpy(ecount).to(self.device), output_size=E)
            self.perm_src_local = (plan.perm_src - d_edge_ptr[pos_g]).contiguous()
        self.nonh_local = (nonh_idx.to(torch.int32) - d_node_ptr[node_g[nonh_idx].long()]).contiguous() if has_mask else None
        self.temperature = big.temperature.contiguous() if getattr(big, "temperature", None) is not None else None
        y = getattr(big, "y", None)
        self.y, self.y_kind = (None, None)
        if y is not None:
            self.y = y.contiguous()
            self.y_kind = NONH if y.dim() == 3 else GRAPH
        self._meta_host = None

    def __len__(self):
        return self.num_graphs

    # ------------------------------------------------------------------------------------------------------------
    def collate(self, ids: Iterable[int]) -> CrystalBatch:
        """Assembles the batch of crystals `ids` on the device (one launch); returns a CrystalBatch whose graph plan is
        already registered, so the model's forward finds the CSR views without building anything."""
        from . import cartnet as CN
        lib = _lib.load()
        ids = np.asarray(list(ids), dtype=np.int64)
        B = int(ids.size)
        sel = self.sizes[:, ids]                                                    # [4, B]
        out_ptr = np.zeros((4, B + 1), dtype=np.int64)
        np.cumsum(sel, axis=1, out=out_ptr[:, 1:])
        N, E, H = int(out_ptr[NODE, -1]), int(out_ptr[EDGE, -1]), int(out_ptr[NONH, -1])
        meta = torch.from_numpy(np.concatenate([ids.astype(np.int32), out_ptr.astype(np.int32).reshape(-1)]))
        meta = meta.pin_memory().to(self.device, non_blocking=True) if self.device.type == "cuda" else meta
        d_ids, d_out_ptr = meta[:B], meta[B:]
        dev = self.device
        i32, i64, f32 = torch.int32, torch.int64, torch.float32
        out = CrystalBatch(
            x=torch.empty(N, dtype=self.x.dtype, device=dev), pos=torch.empty(N, 3, dtype=f32, device=dev),
            batch=torch.empty(N, dtype=i64, device=dev), natoms=torch.empty(B, dtype=i64, device=dev),
            cell=torch.empty(B, 3, 3, dtype=f32, device=dev), edge_index=torch.empty(2, E, dtype=i64, device=dev),
            cart_dist=torch.empty(E, dtype=f32, device=dev), cart_dir=torch.empty(E, 3, dtype=f32, device=dev))
        src32, dst32 = torch.empty(E, dtype=i32, device=dev), torch.empty(E, dtype=i32, device=dev)
        row_ptr, col_ptr = torch.empty(N + 1, dtype=i32, device=dev), torch.empty(N + 1, dtype=i32, device=dev)
        perm_src = torch.empty(E, dtype=i32, device=dev)
        fields = [
            (self.x, out.x, NODE, COPY, self.x.element_size()), (self.pos, out.pos, NODE, COPY, 12),
            (None, out.batch, NODE, SLOT_I64, 8),
            (self.row_start, row_ptr, NODE, I32_EDGE, 4), (self.col_start, col_ptr, NODE, I32_EDGE, 4),
            (self.src_local, out.edge_index[0], EDGE, I32_TO_I64_NODE, 8), (self.dst_local, out.edge_index[1], EDGE, I32_TO_I64_NODE, 8),
            (self.src_local, src32, EDGE, I32_NODE, 4), (self.dst_local, dst32, EDGE, I32_NODE, 4),
            (self.cart_dist, out.cart_dist, EDGE, COPY, 4), (self.cart_dir, out.cart_dir, EDGE, COPY, 12),
            (self.perm_src_local, perm_src, EDGE, I32_EDGE, 4),
            (self.natoms, out.natoms, GRAPH, COPY, 8), (self.cell, out.cell, GRAPH, COPY, 36),
        ]
        if self.mask is not None:
            out.non_H_mask = torch.empty(N, dtype=torch.bool, device=dev)
            out.non_H_index = torch.empty(H, dtype=i64, device=dev)
            fields += [(self.mask, out.non_H_mask, NODE, COPY, 1), (self.nonh_local, out.non_H_index, NONH, I32_TO_I64_NODE, 8)]
        if self.temperature is not None:
            out.temperature = torch.empty(B, dtype=f32, device=dev)
            fields.append((self.temperature, out.temperature, GRAPH, COPY, 4))
        if self.y is not None:
            if self.y_kind == NONH:
                out.y = torch.empty(H, 3, 3, dtype=f32, device=dev)
                fields.append((self.y, out.y, NONH, COPY, 36))
            else:
                out.y = torch.empty(B, dtype=f32, device=dev)
                fields.append((self.y, out.y, GRAPH, COPY, 4))
        table = (_lib.CollateField * len(fields))()
        for k, (src, dst, kind, op, eb) in enumerate(fields):
            table[k].src, table[k].dst = (None if src is None else src.data_ptr()), dst.data_ptr()
            table[k].kind, table[k].op, table[k].elem_bytes = kind, op, eb
        totals = (C.c_int64 * 4)(N, E, H, B)
        st = ops._stream()
        _lib.check(lib.cartnet_collate(table, len(fields), d_ids.data_ptr(), B, self.blob_ptr.data_ptr(), int(self.blob_ptr.shape[1]),
                                       d_out_ptr.data_ptr(), totals, st), "collate")
        _lib.check(lib.cartnet_collate_close_csr(row_ptr.data_ptr(), col_ptr.data_ptr(), N, E, st), "collate_close_csr")
        out.edges_dst_sorted = True
        out._cn_keepalive = meta
        CN.register_plan(out, ops.GraphPlan(N, E, src32, dst32, row_ptr, col_ptr, perm_src, None))
        return out


class DeviceLoader:
    """Iterates a DeviceDataset in batches (the role of the reference's DataLoader(dataset, batch_size, shuffle),
    loader.py:114-124): every batch is assembled on the device, no worker processes, no pinned staging copies."""

    def __init__(self, dataset: DeviceDataset, batch_size: int, shuffle: bool = False, seed: int = 0, drop_last: bool = False):
        self.ds, self.bs, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.rng = np.random.default_rng(seed)

    def __len__(self):
        n = len(self.ds)
        return n // self.bs if self.drop_last else (n + self.bs - 1) // self.bs

    def __iter__(self):
        order = self.rng.permutation(len(self.ds)) if self.shuffle else np.arange(len(self.ds))
        for k in range(len(self)):
            yield self.ds.collate(order[k * self.bs:(k + 1) * self.bs])
