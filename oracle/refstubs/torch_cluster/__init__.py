import numpy as np
import torch
from scipy.spatial import cKDTree


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target", **kw):
    """k nearest neighbours per node without self loops, grouped by centre.

    source_to_target: row0 = neighbour, row1 = centre; target_to_source: swapped."""
    pts = x.detach().cpu().numpy().astype(np.float64)
    tree = cKDTree(pts)
    _, nn = tree.query(pts, k=k + 1)
    n = len(pts)
    out = np.empty((n, k), dtype=np.int64)
    for r in range(n):
        row = [c for c in nn[r] if c != r][:k]
        out[r] = row
    centre = np.repeat(np.arange(n), k)
    nbr = out.reshape(-1)
    if flow == "source_to_target":
        ei = np.stack([nbr, centre])
    else:
        ei = np.stack([centre, nbr])
    return torch.from_numpy(ei).long()
