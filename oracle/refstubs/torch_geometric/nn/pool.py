import numpy as np
import torch
from scipy.spatial import cKDTree


def knn(x, y, k):
    """For every row of y, the k nearest rows of x.  row0 = index into y, row1 = index into x.

    torch_cluster's CPU path is an fp32 nanoflann KD-tree; this stand-in queries a
    KD-tree over the fp32 values (exactly representable in fp64)."""
    tree = cKDTree(x.detach().cpu().numpy().astype(np.float64))
    _, nn = tree.query(y.detach().cpu().numpy().astype(np.float64), k=k)
    nn = np.asarray(nn).reshape(len(y), k)
    rows = np.repeat(np.arange(len(y)), k)
    return torch.from_numpy(np.stack([rows, nn.reshape(-1)])).long()
