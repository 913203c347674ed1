"""Voting-tensor decompositions with the interface of the reference's Pointcloud/Modules/Decompositionor.py:
`Decomposition` (:25-127) and `Decompositionor.getBetterFilteredNVT` (:278-300)."""
from __future__ import annotations

import torch

from . import _lib
from .Selector import Selection
from .Utils import GeneralUtils


class Decomposition:
    def __init__(self, eigval: torch.Tensor, eigvec: torch.Tensor):
        assert eigval.dim() == 2
        assert eigval.size(1) == 3
        assert eigval.is_floating_point()
        assert eigvec.dim() == 3
        assert eigvec.size(1) == 3 and eigvec.size(2) == 3
        assert eigvec.is_floating_point()
        assert eigval.size(0) == eigvec.size(0)
        self.eigval, self.eigvec = eigval, eigvec

    def __len__(self):
        return self.eigval.size(0)

    def getNVTFeatures(self):
        """planarity, linearity, sphericity of the ascending eigenvalues (:57-63)."""
        ev = _lib.dev(self.eigval, torch.float32, "eigval")
        m = ev.size(0)
        lab = torch.empty(m, dtype=torch.uint8, device=ev.device)
        feat = torch.empty((m, 3), dtype=torch.float32, device=ev.device)
        _lib.check(_lib.load().ngpd_classify(_lib.ptr(ev), m, 1.0, _lib.ptr(lab), _lib.ptr(feat), _lib.stream()), "ngpd_classify")
        return feat[:, 0], feat[:, 1], feat[:, 2]

    def getClasses(self, scale: float = 0.2) -> torch.Tensor:
        """argmax(scale*planarity, linearity, sphericity): 0 flat, 1 edge, 2 corner (:65-69)."""
        ev = _lib.dev(self.eigval, torch.float32, "eigval")
        m = ev.size(0)
        lab = torch.empty(m, dtype=torch.uint8, device=ev.device)
        _lib.check(_lib.load().ngpd_classify(_lib.ptr(ev), m, float(scale), _lib.ptr(lab), None, _lib.stream()), "ngpd_classify")
        return lab.long()

    def getVUFeatures(self, tau: float) -> torch.Tensor:
        return (self.eigval < tau).sum(dim=1) % 3

    def getBetterVUFeatures(self, mean_graph_edge_length: float, k: int = 6) -> torch.Tensor:
        """(:87-90) the same count with the threshold scaled to the cloud: tau = 16 / k * l^2."""
        tau = 16.0 / k * mean_graph_edge_length ** 2
        return (self.eigval < tau).sum(dim=1) % 3

    def getVUSmoothedNormals(self, n: torch.Tensor, tau: float = 0.3, d: float = 3) -> torch.Tensor:
        """Eigen-space smoothing exactly as the reference evaluates it (:92-106), eigenvector signs included."""
        ev = _lib.dev(self.eigval, torch.float32, "eigval")
        vec = _lib.dev(self.eigvec, torch.float32, "eigvec")
        n = _lib.dev(n, torch.float32, "n")
        assert n.size(0) == ev.size(0)
        out = torch.empty_like(n)
        _lib.check(_lib.load().ngpd_smooth_normals(_lib.ptr(ev), _lib.ptr(vec), _lib.ptr(n), ev.size(0), float(tau), float(d),
                                                   _lib.ptr(out), _lib.stream()), "ngpd_smooth_normals")
        return out


class Decompositionor:
    def __init__(self, graph):
        GeneralUtils.validateAttributes(graph, ["pos"])
        self.graph = graph

    def getBetterFilteredNVT(self, selection: Selection, _n: torch.Tensor, rho: float = 0.9) -> Decomposition:
        """Normal voting tensor over the neighbours whose normal is within the angle band of the connecting
        direction; rows whose filter rejects everybody use all neighbours; LAPACK-order eigen-decomposition."""
        pos = _lib.dev(self.graph.pos, torch.float32, "graph.pos")
        nrm = _lib.dev(_n, torch.float32, "n")
        m = len(selection)
        rows = selection.i.to(torch.int32).contiguous()
        k = selection.uniform_k()
        if k is not None:
            idx, off = selection.table(), None
        else:
            idx, off = selection.csr()
            k = 0
        eigval = torch.empty((m, 3), dtype=torch.float32, device=pos.device)
        eigvec = torch.empty((m, 3, 3), dtype=torch.float32, device=pos.device)
        _lib.check(_lib.load().ngpd_nvt(_lib.ptr(pos), _lib.ptr(nrm), _lib.ptr(idx), _lib.ptr(off), _lib.ptr(rows), m, k,
                                        _lib.acos_threshold(rho), _lib.ptr(eigval), _lib.ptr(eigvec), None, None, _lib.stream()),
                   "ngpd_nvt")
        return Decomposition(eigval, eigvec)

    def _normal_filtered(self, which: str, selection: Selection, _n: torch.Tensor, rho: float) -> Decomposition:
        pos = _lib.dev(self.graph.pos, torch.float32, "graph.pos")
        nrm = _lib.dev(_n, torch.float32, "n")
        m = len(selection)
        rows = selection.i.to(torch.int32).contiguous()
        k = selection.uniform_k()
        if k is not None:
            idx, off = selection.table(), None
        else:
            idx, off = selection.csr()
            k = 0
        eigval = torch.empty((m, 3), dtype=torch.float32, device=pos.device)
        eigvec = torch.empty((m, 3, 3), dtype=torch.float32, device=pos.device)
        lib, x_le = _lib.load(), _lib.acos_threshold_le(rho)
        if which == "nvt":
            _lib.check(lib.ngpd_nvt_normal(_lib.ptr(nrm), _lib.ptr(idx), _lib.ptr(off), _lib.ptr(rows), m, k, x_le, _lib.ptr(eigval),
                                           _lib.ptr(eigvec), None, None, _lib.stream()), "ngpd_nvt_normal")
        else:
            _lib.check(lib.ngpd_pvt_normal(_lib.ptr(pos), _lib.ptr(nrm), _lib.ptr(idx), _lib.ptr(off), _lib.ptr(rows), m, k, x_le,
                                           _lib.ptr(eigval), _lib.ptr(eigvec), None, None, _lib.stream()), "ngpd_pvt_normal")
        return Decomposition(eigval, eigvec)

    def getNormalFilteredNVT(self, selection: Selection, _n: torch.Tensor, rho: float = 0.9) -> Decomposition:
        """Yadav-2018 normal voting tensor: neighbours whose normal is within rho of the centre's vote (Decompositionor.py:260-276)."""
        return self._normal_filtered("nvt", selection, _n, rho)

    def getNormalFilteredPVT(self, selection: Selection, _n: torch.Tensor, rho: float = 0.9) -> Decomposition:
        """Yadav-2018 point voting tensor: covariance of the neighbours whose normal is within rho of the centre's,
        about their own mean, with the reference's fall-backs (Decompositionor.py:172-211)."""
        return self._normal_filtered("pvt", selection, _n, rho)

