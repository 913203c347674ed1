import torch


def sort_edge_index(edge_index, *a, **k):
    n = int(edge_index.max()) + 1 if edge_index.numel() else 1
    order = torch.argsort(edge_index[0] * n + edge_index[1], stable=True)
    return edge_index[:, order]


def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce="add"):
    n = num_nodes if num_nodes is not None else int(edge_index.max()) + 1
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    key = both[0] * n + both[1]
    uniq, inv = torch.unique(key, return_inverse=True)
    out = torch.stack([uniq // n, uniq % n])
    if edge_attr is None:
        return out
    attr = torch.cat([edge_attr, edge_attr], dim=0)
    red = torch.zeros((uniq.numel(),) + attr.shape[1:], dtype=attr.dtype).index_add_(0, inv, attr)
    return out, red


def degree(index, num_nodes=None, dtype=None):
    n = num_nodes if num_nodes is not None else int(index.max()) + 1
    return torch.zeros(n, dtype=dtype or torch.float).index_add_(
        0, index, torch.ones(index.numel(), dtype=dtype or torch.float))


def subgraph(subset, edge_index, relabel_nodes=False, **k):
    raise NotImplementedError("subgraph stub")
