"""Neighbourhood selection with the interface of the reference's Pointcloud/Modules/Selector.py:
`Selection` (CSR of neighbour ids, :41-134) and `Selector` (:136-246).  The SciPy KD-tree of the reference is
replaced by the frozen grid of libngpd; like that tree it is built ONCE from the positions the graph holds
when the Selector is constructed and is never rebuilt, so later queries answer "which construction-time
points are nearest to the current position of point i" (SURVEY.md 3.1)."""
from __future__ import annotations

import torch

from . import _lib
from .Utils import GeneralUtils, TorchUtils


class Selection:
    """i: centre ids [M]; j: neighbour ids [E]; slices: [M+1] row boundaries into j."""

    def __init__(self, i: torch.Tensor, j: torch.Tensor, slices: torch.Tensor, _table: torch.Tensor = None):
        for name, t in (("i", i), ("j", j), ("slices", slices)):
            assert t.dim() == 1, f"Actual size of {name}: {t.size()}"
            assert not t.is_floating_point()
        assert slices[0] == 0, f"Slider start: {slices[0]}"
        assert i.size(0) == slices.size(0) - 1
        assert j.size(0) == slices[-1], f"Data size: {j.size(0)}\nSlices last entry: {slices[-1]}"
        self.i, self.j, self.slices = i, j, slices
        self._table = _table    # int32 [M,k] view of j when every row has k entries (what the kernels read)

    # -- kernel-facing views -------------------------------------------------------------------------
    def uniform_k(self):
        if self._table is not None:
            return self._table.size(1)
        m = len(self)
        if m == 0:
            return None
        lens = self.slices[1:] - self.slices[:-1]
        k = int(lens[0])
        return k if bool((lens == k).all()) else None

    def table(self) -> torch.Tensor:
        """int32 [M,k] neighbour table (rows must have equal length)."""
        if self._table is None:
            k = self.uniform_k()
            assert k is not None, "ragged selection: use csr()"
            self._table = self.j.view(len(self), k).to(torch.int32).contiguous()
        return self._table

    def csr(self):
        return self.j.to(torch.int32).contiguous(), self.slices.to(torch.int32).contiguous()

    # -- reference interface ---------------------------------------------------------------------------
    def __len__(self):
        return self.slices.size(0) - 1

    def __getitem__(self, key: int):
        assert key >= 0 and key <= len(self)
        return self.j[self.slices[key]:self.slices[key + 1]]

    def filter(self, indices: torch.Tensor) -> "Selection":
        """Row subset (Selector.py:85-92)."""
        starts, ends = self.slices[indices], self.slices[indices + 1]
        new_slices = torch.cat([torch.zeros(1, dtype=torch.long, device=self.slices.device), (ends - starts).cumsum(0)])
        if self._table is not None:
            tab = self._table[indices]
            return Selection(self.i[indices], tab.reshape(-1).long(), new_slices, tab.contiguous())
        return Selection(self.i[indices], self.j[TorchUtils.rangeBoundariesToIndices(starts, ends)], new_slices)

    @classmethod
    def fromEdgeIndex(cls, edge_index: torch.Tensor) -> "Selection":
        n = int(edge_index.max()) + 1
        order = torch.argsort(edge_index[0] * n + edge_index[1])
        start, end = edge_index[0][order], edge_index[1][order]
        uniq, counts = start.unique(return_counts=True)
        slices = torch.zeros(uniq.numel() + 1, dtype=torch.long, device=edge_index.device)
        slices[1:] = counts.cumsum(0)
        return cls(uniq, end, slices)

    def getBatchIndex(self) -> torch.Tensor:
        m = self.i.size(0)
        ids = self.i if self.i[-1] == m - 1 else torch.arange(m, device=self.i.device)
        return torch.repeat_interleave(ids, self.slices[1:] - self.slices[:-1])

    def getEdgeIndex(self) -> torch.Tensor:
        start = torch.repeat_interleave(self.i, self.slices[1:] - self.slices[:-1])
        return torch.vstack([start[None], self.j[None]])

    def scatter(self, source: torch.Tensor, reduce: str) -> torch.Tensor:
        """Segmented reduction over rows (Selector.py:127-134)."""
        rows = torch.repeat_interleave(torch.arange(len(self), device=source.device), self.slices[1:] - self.slices[:-1])
        shape = (len(self),) + tuple(source.shape[1:])
        idx = rows.view((-1,) + (1,) * (source.dim() - 1)).expand_as(source)
        if reduce == "add":
            return torch.zeros(shape, dtype=source.dtype, device=source.device).scatter_add_(0, idx, source)
        if reduce == "mean":
            return torch.zeros(shape, dtype=source.dtype, device=source.device).scatter_reduce_(0, idx, source, "mean", include_self=False)
        if reduce == "max":
            out = torch.full(shape, float("-inf"), dtype=source.dtype, device=source.device)
            return out.scatter_reduce_(0, idx, source, "amax"), None
        raise ValueError(reduce)


class Selector:
    def __init__(self, graph):
        GeneralUtils.validateAttributes(graph, ["pos"])
        _lib.require_cuda()
        self.graph = graph
        # frozen copy, like scipy's KDTree(graph.pos.cpu()) at Selector.py:141
        self.tree_pos = graph.pos.detach().to(torch.float32).contiguous().clone()
        self._grids = {}

    def _grid(self, k: int) -> _lib.Grid:
        bucket = 8 if k <= 10 else (16 if k <= 24 else (32 if k <= 48 else 64))
        if bucket not in self._grids:
            self._grids[bucket] = _lib.Grid(self.tree_pos, k_hint=bucket)
        return self._grids[bucket]

    def getPointsInRangeSelectionVectorized(self, radii: torch.Tensor, indices: torch.Tensor = None) -> Selection:
        """Every construction-time point within radii[r] of the CURRENT position of (selected) point r, ascending by
        index -- scipy's query_ball_point on the frozen tree (Selector.py:214-229)."""
        from .Utils import TorchUtils
        TorchUtils.validateIndices(indices)
        pos = self.graph.pos
        N = pos.size(0) if indices is None else indices.size(0)
        assert radii.dim() == 1
        assert radii.size(0) == N, f"Actual: {radii.size(0)}\nExpected: {N}"
        assert radii.is_floating_point()
        flags = 0
        if indices is None:
            query = pos
            flags = _lib.KNN_QUERY_IS_TREE if pos.size(0) == self.tree_pos.size(0) else 0
            indices = torch.arange(N, dtype=torch.long, device=pos.device)
        else:
            query = pos[indices]
        j, off = self._grid(16).ball(query, radii.to(device=pos.device, dtype=torch.float32), flags)
        return Selection(indices, j.long(), off.long())

    def getPointsInRangeSelection(self, radius: float, indices: torch.Tensor = None) -> Selection:
        n = self.graph.pos.size(0) if indices is None else indices.size(0)
        return self.getPointsInRangeSelectionVectorized(torch.full((n,), float(radius), dtype=torch.float32, device=self.graph.pos.device), indices)

    def getKNNSelection(self, k: int, indices: torch.Tensor = None) -> Selection:
        """k nearest construction-time points of the CURRENT position of every (selected) point, self included,
        ascending by distance (Selector.py:235-246)."""
        pos = self.graph.pos
        flags = 0
        if indices is None:
            indices = torch.arange(pos.size(0), dtype=torch.long, device=pos.device)
            query = pos
            flags = _lib.KNN_QUERY_IS_TREE if pos.size(0) == self.tree_pos.size(0) else 0
        else:
            if not torch.is_tensor(indices) or indices.is_floating_point():
                raise ValueError("indices should contain integer values and not floating point values.")
            query = pos[indices]
        table = self._grid(k).knn(query, k, flags)
        slices = torch.arange(table.size(0) + 1, device=table.device, dtype=torch.long) * k
        return Selection(indices, table.reshape(-1).long(), slices, table)
