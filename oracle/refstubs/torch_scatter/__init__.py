import torch


def scatter_sum(src, index, dim=0, dim_size=None, **k):
    n = dim_size if dim_size is not None else int(index.max()) + 1
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add_(0, index, src)


def scatter_mean(src, index, dim=0, dim_size=None, **k):
    n = dim_size if dim_size is not None else int(index.max()) + 1
    s = scatter_sum(src, index, dim, n)
    cnt = torch.zeros(n, dtype=src.dtype).index_add_(0, index, torch.ones(index.numel(), dtype=src.dtype))
    cnt = cnt.clamp_(min=1)
    return s / cnt.view((-1,) + (1,) * (src.dim() - 1))


def scatter_max(src, index, dim=0, dim_size=None, **k):
    n = dim_size if dim_size is not None else int(index.max()) + 1
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    out = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype)
    out = out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    return out, None
