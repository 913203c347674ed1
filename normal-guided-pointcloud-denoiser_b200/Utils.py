"""Metrics and small helpers with the names and argument meaning of the reference's
Pointcloud/Modules/Utils.py (`TorchUtils`, `GeneralUtils`); the nearest-neighbour work runs in libngpd."""
from __future__ import annotations

import torch

from . import _lib


class GeneralUtils:
    @classmethod
    def validateAttributes(cls, obj, attrs) -> None:
        # Utils.py:55-66: ValueError naming the first missing attribute
        for a in attrs:
            if getattr(obj, a, None) is None:
                raise ValueError(f"Object does not have attribute '{a}'.")


def _check_cloud(t: torch.Tensor):
    assert t.dim() == 2
    assert t.size(1) == 3


def _nearest(tree_pos: torch.Tensor, query: torch.Tensor) -> torch.Tensor:
    """fp32 squared distance from every query to its nearest row of tree_pos (one grid build + one 1-NN pass)."""
    grid = _lib.Grid(tree_pos, k_hint=4)
    return grid.nn_sqdist(query)


class TorchUtils:
    @classmethod
    def validateIndices(cls, indices) -> None:
        assert indices is None or (not indices.is_floating_point() and indices.dim() == 1)

    @classmethod
    def validateEdgeIndex(cls, edge_index: torch.Tensor) -> None:
        assert not edge_index.is_floating_point()
        assert edge_index.dim() == 2
        assert edge_index.size(0) == 2

    @classmethod
    def validateKNNEdgeIndex(cls, edge_index: torch.Tensor):
        """k of a regular kNN edge list (every centre has the same number of edges), Utils.py:220-225."""
        cls.validateEdgeIndex(edge_index)
        counts = torch.bincount(edge_index[0])
        counts = counts[counts > 0].unique()
        assert counts.numel() == 1
        return counts[0]

    @classmethod
    def ChamferDistance(cls, pos0: torch.Tensor, pos1: torch.Tensor) -> torch.Tensor:
        """Utils.py:253-265: un-reduced squared distances, first every pos1 point to its nearest pos0 point,
        then every pos0 point to its nearest pos1 point.  Callers take .mean()."""
        _check_cloud(pos0); _check_cloud(pos1)
        return torch.cat([_nearest(pos0, pos1), _nearest(pos1, pos0)], dim=0)

    @classmethod
    def SingleChamferDistance(cls, gt: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        """sCD.  PostProcessing.ipynb#c9 uses TorchUtils.SingleChamferDistance but the reference checkout does not
        define it; defined here as the first half of ChamferDistance(gt, pos): each evaluated point to its nearest
        ground-truth point (the direction PaperDistance uses, Utils.py:292-293)."""
        _check_cloud(gt); _check_cloud(pos)
        return _nearest(gt, pos)

    @classmethod
    def HausdorffDistance(cls, pos0: torch.Tensor, pos1: torch.Tensor) -> torch.Tensor:
        """Utils.py:267-279: as ChamferDistance with Euclidean norms; callers take .max()."""
        return cls.ChamferDistance(pos0, pos1).sqrt_()

    @classmethod
    def PaperDistance(cls, gt: torch.Tensor, noisy: torch.Tensor) -> torch.Tensor:
        """Utils.py:281-295: distance of every noisy point to its nearest gt point over the gt bbox diagonal."""
        _check_cloud(gt); _check_cloud(noisy)
        diag = (gt.max(dim=0).values - gt.min(dim=0).values).norm(dim=0)
        return _nearest(gt, noisy).sqrt_() / diag

    @classmethod
    def averageEdgeLength(cls, pos: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        """Utils.py:298-299."""
        return (pos[edge_index[1]] - pos[edge_index[0]]).norm(dim=1).mean(dim=0)

    @classmethod
    def pointcloudRadius(cls, pos: torch.Tensor) -> torch.Tensor:
        """Utils.py:302-303."""
        return (pos - pos.mean(dim=0, keepdim=True)).norm(dim=1).max(dim=0).values

    @classmethod
    def rangeBoundariesToIndices(cls, starts: torch.Tensor, ends: torch.Tensor) -> torch.Tensor:
        """Utils.py:311-327: concatenation of arange(starts[i], ends[i]); empty or reversed ranges are skipped."""
        assert starts.dtype == torch.long
        lens = (ends - starts).clamp_(min=0)
        total = int(lens.sum())
        if total == 0:
            return torch.empty(0, dtype=torch.long, device=starts.device)
        first = torch.cumsum(lens, 0) - lens                      # output offset of each range
        owner = torch.repeat_interleave(torch.arange(lens.numel(), device=starts.device), lens)
        return starts[owner] + (torch.arange(total, device=starts.device) - first[owner])
