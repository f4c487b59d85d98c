import inspect
import torch


class MessagePassing(torch.nn.Module):
    """Minimal PyG-2.5 MessagePassing: flow=source_to_target, node_dim=-2.
    kwargs suffixed _i index edge_index[1], _j index edge_index[0]; `index` is
    edge_index[1]; routing is by parameter NAME of message/aggregate/update and
    propagate returns whatever update returns."""

    def __init__(self, *a, **kw):
        super().__init__()

    def _collect(self, fn, edge_index, kwargs, extra):
        out = {}
        for name in list(inspect.signature(fn).parameters):
            if name in extra:
                out[name] = extra[name]
            elif name.endswith("_i"):
                out[name] = kwargs[name[:-2]].index_select(0, edge_index[1])
            elif name.endswith("_j"):
                out[name] = kwargs[name[:-2]].index_select(0, edge_index[0])
            elif name == "index":
                out[name] = edge_index[1]
            elif name in kwargs:
                out[name] = kwargs[name]
        return out

    def propagate(self, edge_index, **kwargs):
        msg = self.message(**self._collect(self.message, edge_index, kwargs, {}))
        first = list(inspect.signature(self.aggregate).parameters)[0]
        agg = self.aggregate(**self._collect(self.aggregate, edge_index, kwargs, {first: msg}))
        first = list(inspect.signature(self.update).parameters)[0]
        return self.update(**self._collect(self.update, edge_index, kwargs, {first: agg}))
