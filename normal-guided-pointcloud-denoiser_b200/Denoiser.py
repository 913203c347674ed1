"""Class-specific position updates with the interface of the reference's Pointcloud/Modules/Denoiser.py
(:18-231): every step takes a `Selection` of the rows to move and returns their new positions."""
from __future__ import annotations

import torch

from . import _lib
from .Selector import Selection


class Denoiser:
    def __init__(self, graph):
        assert hasattr(graph, "pos") and graph.pos is not None
        assert graph.pos.dim() == 2
        assert graph.pos.size(1) == 3
        self.graph = graph

    def _run(self, kind: int, selection: Selection, n: torch.Tensor, d, alpha: float, edge_vectors: torch.Tensor = None):
        pos = _lib.dev(self.graph.pos, torch.float32, "graph.pos")
        assert n.dim() == 2
        assert pos.size(0) == n.size(0)
        assert n.size(1) == 3
        nrm = _lib.dev(n, torch.float32, "n")
        m = len(selection)
        out = torch.empty((m, 3), dtype=torch.float32, device=pos.device)
        if m == 0:
            return out
        rows = selection.i.to(torch.int32).contiguous()
        k = selection.uniform_k()
        if k is not None:
            idx, off = selection.table(), None
        else:
            idx, off = selection.csr()
            k = 0
        cd = None
        lib = _lib.load()
        if kind == _lib.STEP_FLAT:
            cd = torch.empty(4, dtype=torch.float32, device=pos.device)
            _lib.check(lib.ngpd_center_delta(_lib.ptr(pos), _lib.ptr(idx), idx.numel(), _lib.ptr(cd), _lib.stream()), "ngpd_center_delta")
        edge = _lib.dev(edge_vectors, torch.float32, "edge_vectors") if edge_vectors is not None else None
        _lib.check(lib.ngpd_update(kind, _lib.ptr(pos), _lib.ptr(nrm), _lib.ptr(edge), _lib.ptr(idx), _lib.ptr(off), _lib.ptr(rows), m, k,
                                   float(alpha), float(d), _lib.ptr(cd), _lib.ptr(out), _lib.stream()), "ngpd_update")
        return out

    def corner_step(self, selection: Selection, n: torch.Tensor, d: float, alpha: float = 0.1):
        return self._run(_lib.STEP_CORNER, selection, n, d, alpha)

    def edge_step(self, selection: Selection, n: torch.Tensor, edge_vectors: torch.Tensor, d: float, alpha: float = 0.1):
        return self._run(_lib.STEP_EDGE, selection, n, d, alpha, edge_vectors)

    def flat_step(self, selection: Selection, n: torch.Tensor, d: float, alpha: float = 0.1):
        return self._run(_lib.STEP_FLAT, selection, n, d, alpha)

    def feature_step(self, selection: Selection, n: torch.Tensor, d: float, alpha: float = 0.1):
        return self._run(_lib.STEP_FEATURE, selection, n, d, alpha)

    def dummy_step(self, selection: Selection, n: torch.Tensor, d: float, alpha: float = 0.1):
        assert n.dim() == 2
        assert self.graph.pos.size(0) == n.size(0)
        assert n.size(1) == 3
        return self.graph.pos[selection.i].clone()
