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


# The metrics are called in loops against a target that never changes (the ground truth of denoiseUntilMinimumError,
# Processor.py:157-176, is evaluated every iteration): the index over an unchanged tensor is kept instead of being rebuilt per
# call as the reference's KD-tree is (Utils.py:260-261).  Keyed on storage address, shape and the tensor's version counter, so
# an in-place edit rebuilds it; the entry holds the tensor, so the address cannot be recycled while it is cached.
_GRID_CACHE: dict = {}
_GRID_CACHE_MAX_POINTS = 8_000_000
_GRID_CACHE_ENTRIES = 2


def _grid_for(tree_pos: torch.Tensor) -> "_lib.Grid":
    if tree_pos.size(0) > _GRID_CACHE_MAX_POINTS:
        return _lib.Grid(tree_pos, k_hint=4)
    key = (tree_pos.data_ptr(), tuple(tree_pos.shape), tree_pos._version, str(tree_pos.device), tree_pos.dtype)
    hit = _GRID_CACHE.pop(key, None)
    if hit is None:
        hit = (_lib.Grid(tree_pos, k_hint=4), tree_pos)
    _GRID_CACHE[key] = hit                                       # most recently used last
    while len(_GRID_CACHE) > _GRID_CACHE_ENTRIES:
        _GRID_CACHE.pop(next(iter(_GRID_CACHE)))
    return hit[0]


def _nearest(tree_pos: torch.Tensor, query: torch.Tensor) -> torch.Tensor:
    """fp32 squared distance from every query to its nearest row of tree_pos (one 1-NN pass; the index is cached)."""
    return _grid_for(tree_pos).nn_sqdist(query)


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
    def clearIndexCache(cls) -> None:
        _GRID_CACHE.clear()

    @classmethod
    def chamferSummary(cls, pos0: torch.Tensor, pos1: torch.Tensor) -> dict:
        """What callers reduce the metrics to, without materialising the per-point vectors (ngpd_nn_sqdist_reduce: the block
        reduction is fused into the nearest-neighbour kernel): {"chamfer": ChamferDistance(pos0, pos1).mean(),
        "single_chamfer": SingleChamferDistance(pos0, pos1).mean(), "hausdorff": HausdorffDistance(pos0, pos1).max(),
        "mean_distance": mean distance of pos1 to pos0} as 0-d fp64 device tensors (no host synchronisation)."""
        _check_cloud(pos0); _check_cloud(pos1)
        a = _grid_for(pos0).nn_reduce(pos1)                      # every pos1 point -> nearest pos0 point
        b = _grid_for(pos1).nn_reduce(pos0)
        return {"chamfer": (a[0] + b[0]) / (a[3] + b[3]), "single_chamfer": a[0] / a[3],
                "hausdorff": torch.maximum(a[2], b[2]).sqrt(), "mean_distance": a[1] / a[3]}

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
