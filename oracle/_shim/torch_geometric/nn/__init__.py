import torch
from . import conv  # noqa: F401


class Linear(torch.nn.Linear):
    """pyg_nn.Linear(in, out, bias=True) == F.linear; identical state-dict keys
    and U(+-1/sqrt(in)) init."""

    def __init__(self, in_channels, out_channels, bias=True, **kw):
        super().__init__(in_channels, out_channels, bias=bias)
