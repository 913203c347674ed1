"""Graph construction, PCA normals and normal orientation with the interface of the reference's
Pointcloud/Modules/GraphBuilder.py (:40-209)."""
from __future__ import annotations

import math

import torch

from . import _lib
from .Object import Pointcloud
from .Utils import GeneralUtils, TorchUtils


class Graph:
    """Attribute bag standing in for torch_geometric.data.Data: pos, n, edge_index, edge_attr, gt, gt_n."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self) -> int:
        return self.pos.size(0)

    @property
    def num_edges(self) -> int:
        return self.edge_index.size(1)


class GraphBuilder:
    def __init__(self, pointcloud: Pointcloud):
        GeneralUtils.validateAttributes(pointcloud, ["v"])
        _lib.require_cuda()
        if not pointcloud.v.is_cuda:
            pointcloud.v = pointcloud.v.cuda()
            if pointcloud.n is not None:
                pointcloud.n = pointcloud.n.cuda()
        self.device = pointcloud.v.device
        self.pointcloud = pointcloud
        self.graph = Graph(pos=pointcloud.v)
        if pointcloud.hasNormals():
            self.graph.n = pointcloud.n

    def getKNNEdgeIndex(self, k: int = 12) -> torch.Tensor:
        """kNN graph of the CURRENT positions without self loops, row 0 = centre, row 1 = neighbour, grouped by
        centre, neighbours ascending by distance (torch_cluster.knn_graph(..., flow="target_to_source"), :60-63)."""
        g = self.graph
        GeneralUtils.validateAttributes(g, ["pos"])
        pos = _lib.dev(g.pos, torch.float32, "graph.pos")
        table = _lib.Grid(pos, k_hint=k).knn(pos, k, _lib.KNN_SKIP_SELF | _lib.KNN_QUERY_IS_TREE)
        centre = torch.arange(pos.size(0), device=pos.device).repeat_interleave(k)
        return torch.stack([centre, table.reshape(-1).long()])

    def getPVTDecompositionWithKNN(self, edge_index: torch.Tensor) -> torch.Tensor:
        """Eigenvectors (columns, ascending eigenvalue) of each point's neighbour covariance (:99-111)."""
        w, vec = self._pca(edge_index)
        return vec

    def _pca(self, edge_index: torch.Tensor):
        g = self.graph
        GeneralUtils.validateAttributes(g, ["pos"])
        k = int(TorchUtils.validateKNNEdgeIndex(edge_index))
        pos = _lib.dev(g.pos, torch.float32, "graph.pos")
        n = pos.size(0)
        table = edge_index[1].view(-1, k).to(torch.int32).contiguous()
        eigval = torch.empty((n, 3), dtype=torch.float32, device=pos.device)
        vec = torch.empty((n, 3, 3), dtype=torch.float32, device=pos.device)
        _lib.check(_lib.load().ngpd_pca_normals(_lib.ptr(pos), _lib.ptr(table), None, n, k, None, _lib.ptr(eigval), _lib.ptr(vec),
                                                _lib.stream()), "ngpd_pca_normals")
        return eigval, vec

    def setPVTNormals(self, edge_index: torch.Tensor) -> None:
        g = self.graph
        k = int(TorchUtils.validateKNNEdgeIndex(edge_index))
        pos = _lib.dev(g.pos, torch.float32, "graph.pos")
        n = pos.size(0)
        table = edge_index[1].view(-1, k).to(torch.int32).contiguous()
        normals = torch.empty((n, 3), dtype=torch.float32, device=pos.device)
        _lib.check(_lib.load().ngpd_pca_normals(_lib.ptr(pos), _lib.ptr(table), None, n, k, _lib.ptr(normals), None, None, _lib.stream()),
                   "ngpd_pca_normals")
        g.n = normals

    def setAndFlipNormals(self, flip: bool = True) -> None:
        g = self.graph
        GeneralUtils.validateAttributes(g, ["edge_index"])
        self.setPVTNormals(g.edge_index)
        if flip:
            self.flipNormals()

    # ---- orientation (:129-209): edge cost 1-|ni.nj|, minimum spanning tree, propagate from the top-most point
    # flipping a child when n_parent.n_child < cos(7pi/12).  On the GPU (Boruvka rounds + breadth-first frontier,
    # csrc/orient.cu); there is no host version in this package (the SciPy cross-check lives in oracle/).
    def calculateEdgeCost(self) -> None:
        g = self.graph
        GeneralUtils.validateAttributes(g, ["edge_index", "n"])
        nn = g.n[g.edge_index]
        g.edge_attr = 1 - (nn[0] * nn[1]).sum(dim=-1).abs_()

    def flipNormals(self) -> None:
        import ctypes
        g = self.graph
        GeneralUtils.validateAttributes(g, ["pos", "edge_index", "n"])
        self.calculateEdgeCost()                       # the reference leaves graph.edge_attr behind (:131)
        pos = _lib.dev(g.pos, torch.float32, "graph.pos")
        nrm = _lib.dev(g.n, torch.float32, "graph.n").clone()
        src = g.edge_index[0].to(torch.int32).contiguous()
        dst = g.edge_index[1].to(torch.int32).contiguous()
        info = (ctypes.c_int32 * 3)()
        _lib.check(_lib.load().ngpd_orient_normals(_lib.ptr(pos), _lib.ptr(nrm), pos.size(0), _lib.ptr(src), _lib.ptr(dst), src.numel(),
                                                   math.cos(7.0 / 12.0 * math.pi), info, _lib.stream()), "ngpd_orient_normals")
        self.orientation_info = {"components": info[0], "levels": info[1], "rounds": info[2]}
        g.n = nrm
