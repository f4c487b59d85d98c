"""Minimal stand-in for the PyG `Batch` the reference feeds to `CartNet.forward`.

The model only reads attributes (x, batch, temperature, edge_index,
cart_dist, cart_dir, non_H_mask, y) and writes `x` / `edge_attr` back, so any object
with those attributes works -- a real `torch_geometric.data.Batch` included. This class
exists because PyG is not installable offline; `collate` restates what
`Batch.from_data_list` / the PyG `DataLoader` do for these fields
(/root/reference/loader/loader.py:114-124): concatenate per-crystal tensors, add the
cumulative node offset to `edge_index`, build the `batch` vector.
"""
from __future__ import annotations

import copy

import torch


class CrystalBatch:
    _FIELDS = ("x", "pos", "cell", "natoms", "batch", "temperature", "edge_index", "cart_dist",
               "cart_dir", "non_H_mask", "y", "edge_attr")

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def to(self, device, non_blocking: bool = False):
        """In place, like `Batch.to` as used at /root/reference/train/train.py:169."""
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    def pin_memory(self):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v) and not v.is_cuda:
                setattr(self, k, v.pin_memory())
        return self

    def clone(self):
        """Deep copy of tensors (the reference clones before forward because forward mutates
        the batch in place, /root/reference/main.py:87)."""
        out = CrystalBatch()
        for k, v in self.__dict__.items():
            setattr(out, k, v.clone() if torch.is_tensor(v) else copy.deepcopy(v))
        return out

    @property
    def num_graphs(self) -> int:
        return int(self.natoms.numel())

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])


def collate(items) -> CrystalBatch:
    """items: list of per-crystal dicts / objects with x, pos, cell[1,3,3] or [3,3], edge_index
    (local node ids), cart_dist, cart_dir and optionally temperature, non_H_mask, y."""
    def get(it, k):
        return it[k] if isinstance(it, dict) else getattr(it, k)

    def has(it, k):
        return (k in it) if isinstance(it, dict) else hasattr(it, k)

    xs, poss, cells, nat, bvec, eis, cds, cdirs = [], [], [], [], [], [], [], []
    temps, masks, ys = [], [], []
    off = 0
    for g, it in enumerate(items):
        x = torch.as_tensor(get(it, "x"))
        n = int(x.shape[0])
        xs.append(x)
        poss.append(torch.as_tensor(get(it, "pos")))
        cells.append(torch.as_tensor(get(it, "cell")).reshape(1, 3, 3))
        nat.append(n)
        bvec.append(torch.full((n,), g, dtype=torch.int64))
        eis.append(torch.as_tensor(get(it, "edge_index")) + off)
        cds.append(torch.as_tensor(get(it, "cart_dist")))
        cdirs.append(torch.as_tensor(get(it, "cart_dir")))
        if has(it, "temperature"):
            temps.append(torch.as_tensor(get(it, "temperature")).reshape(1))
        if has(it, "non_H_mask"):
            masks.append(torch.as_tensor(get(it, "non_H_mask")))
        if has(it, "y"):
            y = torch.as_tensor(get(it, "y"))
            ys.append(y.reshape(1) if y.dim() == 0 else y)
        off += n
    out = CrystalBatch(
        x=torch.cat(xs), pos=torch.cat(poss), cell=torch.cat(cells),
        natoms=torch.tensor(nat, dtype=torch.int64), batch=torch.cat(bvec),
        edge_index=torch.cat(eis, dim=1), cart_dist=torch.cat(cds), cart_dir=torch.cat(cdirs))
    if temps:
        out.temperature = torch.cat(temps)
    if masks:
        out.non_H_mask = torch.cat(masks)
    if ys:
        out.y = torch.cat(ys)
    return out


class DevicePrefetcher:
    """Iterates host batches (pinned memory) and hands out DEVICE batches one step ahead: the host->device copies
    and the per-batch graph plan (int32 CSR views used by every layer) of batch i+1 are issued on a side stream while
    batch i trains, so neither sits on the critical path of the step. This is the device-side counterpart of the
    reference's DataLoader(pin_memory=True) + `batch.to("cuda:0")` (/root/reference/loader/loader.py:114-124,
    train/train.py:169)."""

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._next = None
        self._preload()

    def _preload(self):
        from .cartnet import get_plan
        try:
            hb = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            b = CrystalBatch(**hb.__dict__).to(self.device, non_blocking=True)
            if hasattr(b, "edge_index"):
                get_plan(b)                     # cached by edge_index identity; the layers find it ready
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (b, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        b, ev = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        tensors = list(b.__dict__.values())     # everything below was allocated on the side stream
        if hasattr(b, "edge_index"):
            from .cartnet import get_plan
            tensors += list(get_plan(b).__dict__.values())
        for v in tensors:
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(cur)
        self._preload()
        return b
