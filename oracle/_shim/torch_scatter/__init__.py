"""TEST-ONLY stand-in for torch_scatter.scatter (sum / mean), see SURVEY.md App. B."""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1
    shape = (dim_size,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    res = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)
    if reduce in ("sum", "add"):
        return res
    if reduce == "mean":
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).scatter_add_(
            0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))
        return res / cnt
    raise NotImplementedError(reduce)


def segment_coo(src, index, out=None, dim_size=None, reduce="sum"):
    """sum of src over sorted `index` (torch_scatter.segment_coo, used at dataset/utils.py:269)."""
    assert reduce == "sum" and src.dim() == 1
    if dim_size is None:
        dim_size = int(index.max()) + 1
    return torch.zeros(int(dim_size), dtype=src.dtype, device=src.device).scatter_add_(0, index, src)


def segment_csr(src, indptr, out=None, reduce="sum"):
    """sum of src over [indptr[i], indptr[i+1]) (torch_scatter.segment_csr, used at dataset/utils.py:280,342)."""
    assert reduce == "sum" and src.dim() == 1
    c = torch.cat([src.new_zeros(1), torch.cumsum(src, 0)])
    return c[indptr[1:]] - c[indptr[:-1]]
