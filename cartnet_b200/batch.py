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

    _streams = {}      # one copy stream per device for the process: the caching allocator pools blocks per stream, so a
                       # fresh stream per prefetcher (per epoch) would cudaMalloc its batch buffers again every time

    _pending = None    # weak set of prefetchers whose next batch has not been issued yet (see run_pending)

    def __init__(self, batches, device, eager: bool = False):
        """The host work of a prefetch (20 async copies + the graph plan: ~0.4 ms for an ADP-64 batch, 6 ms for 4 096 JARVIS
        crystals) should sit BEHIND the launches of the step that is about to run, not in front of them. eager=False
        (default): `next()` only hands out the batch that is already in flight; the following one is issued by the first of
        (a) an explicit `prefetch_next()` (best: right after `loss.backward()`), (b) the end of the next `CartNet.forward`
        (automatic, `run_pending`), (c) the next `next()` (on demand, no overlap). eager=True: `next()` issues it before it
        returns, as a plain iterator would."""
        self.eager = bool(eager)
        self.it = iter(batches)
        self.device = torch.device(device)
        key = (self.device.type, self.device.index if self.device.index is not None else torch.cuda.current_device())
        if key not in DevicePrefetcher._streams:
            DevicePrefetcher._streams[key] = torch.cuda.Stream(device=self.device)
        self.stream = DevicePrefetcher._streams[key]
        self._next = None
        self._done = False
        self._preload()

    def _preload(self):
        from .cartnet import get_plan
        try:
            hb = next(self.it)
        except StopIteration:
            self._next = None
            self._done = True
            return
        with torch.cuda.stream(self.stream):
            # a PyG Batch keeps its fields in a store, not in __dict__: go through keys() / getattr
            names = list(hb.keys()) if callable(getattr(hb, "keys", None)) else [k for k in hb.__dict__ if not k.startswith("_")]
            fields = {k: getattr(hb, k) for k in names}
            for k in ("edges_dst_sorted", "non_H_index"):       # optional hints set as plain attributes
                if k not in fields and getattr(hb, k, None) is not None:
                    fields[k] = getattr(hb, k)
            b = CrystalBatch(**fields).to(self.device, non_blocking=True)
            if hasattr(b, "edge_index"):
                get_plan(b)                     # cached by edge_index identity; the layers find it ready
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._next = (b, ev)

    def __iter__(self):
        return self

    def prefetch_next(self):
        """Issue the host->device copies + graph plan of the next batch now (no-op when it is already in flight)."""
        if self._next is None and not self._done:
            self._preload()
        if DevicePrefetcher._pending is not None:
            DevicePrefetcher._pending.discard(self)

    @staticmethod
    def run_pending():
        """Called by CartNet.forward once its kernels are launched: issue the next batch of every prefetcher that is waiting."""
        if DevicePrefetcher._pending:
            for p in list(DevicePrefetcher._pending):
                p.prefetch_next()

    def __next__(self):
        self.prefetch_next()                    # eager=False and nobody prefetched: load on demand
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
        self._next = None
        if self.eager:
            self._preload()
        elif not self._done:
            if DevicePrefetcher._pending is None:
                import weakref
                DevicePrefetcher._pending = weakref.WeakSet()
            DevicePrefetcher._pending.add(self)
        return b


class DeferredScalars:
    """Per-step scalars (loss, metrics) read back WITHOUT stalling the step that produced them: `push(t)` issues an
    asynchronous device->host copy of a 1-element tensor into a pinned ring slot and records an event; `pop()` returns
    the oldest pushed value as a float, waiting only for ITS copy (normally long finished). With `depth` >= 2 the host
    keeps issuing step i+1 while step i runs, instead of draining the GPU at every `loss.item()` the way the reference's
    per-iteration metric reads do (/root/reference/train/train.py:192-199, train/metrics.py:202-206).
    CPU tensors are accepted (the value is copied immediately) so that host logic can be tested without a GPU."""

    def __init__(self, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.depth = depth
        self.buf = torch.zeros(depth, dtype=torch.float32)
        if torch.cuda.is_available():
            self.buf = self.buf.pin_memory()
        self.events = [None] * depth
        self.head = 0          # next slot to write
        self.count = 0         # values pushed and not yet popped

    def push(self, t: torch.Tensor):
        """Queue a scalar; returns the oldest value (float) if the ring was full, else None."""
        out = self.pop() if self.count == self.depth else None
        slot = self.head
        src = t.detach().reshape(1).to(torch.float32)
        self.buf[slot:slot + 1].copy_(src, non_blocking=True)
        if src.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(src.device))
            self.events[slot] = ev
        else:
            self.events[slot] = None
        self.head = (self.head + 1) % self.depth
        self.count += 1
        return out

    def pop(self) -> float:
        if self.count == 0:
            raise IndexError("no pending scalar")
        slot = (self.head - self.count) % self.depth
        ev = self.events[slot]
        if ev is not None:
            ev.synchronize()
        self.count -= 1
        return float(self.buf[slot])

    def drain(self):
        """All pending values, oldest first."""
        return [self.pop() for _ in range(self.count)]
