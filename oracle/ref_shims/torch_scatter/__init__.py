"""Stand-in for torch_scatter.scatter (reference call site: models/dmt.py:57)."""
import torch


def scatter(src, index, dim=0, out=None, dim_size=None, reduce='sum'):
    assert dim == 0 and out is None
    if dim_size is None:
        dim_size = int(index.max()) + 1
    shape = (dim_size,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if reduce in ('sum', 'add'):
        return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)
    if reduce == 'mean':
        s = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)
        c = torch.zeros(dim_size, dtype=src.dtype, device=src.device).scatter_add_(
            0, index, torch.ones_like(index, dtype=src.dtype)).clamp(min=1)
        return s / c.view(-1, *([1] * (src.dim() - 1)))
    if reduce == 'max':
        return torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device).scatter_reduce_(
            0, idx, src, 'amax', include_self=True)
    raise ValueError(reduce)
