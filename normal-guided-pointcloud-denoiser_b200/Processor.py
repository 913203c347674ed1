"""Orchestration with the interface of the reference's Pointcloud/Modules/Processor.py (:24-199).
`denoise` and `denoiseUntilMinimumError` run on the fused tree-order session of libngpd; the step-by-step
public operators (Selector / Decompositionor / Denoiser) give the same results and remain available for
callers that compose their own loops, as the reference notebooks do."""
from __future__ import annotations

import math
from typing import Callable

import torch

from . import _lib
from .Decompositionor import Decompositionor
from .Denoiser import Denoiser
from .GraphBuilder import GraphBuilder
from .Noise import Noise
from .Object import Pointcloud
from .Selector import Selector
from .Utils import TorchUtils


class Processor:
    def __init__(self, pointcloud: Pointcloud):
        self.pointcloud = pointcloud
        self.graphBuilder = GraphBuilder(pointcloud)
        graph = self.graphBuilder.graph
        self.graph = graph
        self.selector = Selector(graph)
        self.noise = Noise(graph)
        self.denoiser = Denoiser(graph)
        self.decompositionor = Decompositionor(graph)
        self._session = None

    # ---- fused path -----------------------------------------------------------------------------------
    def _get_session(self) -> _lib.Session:
        if self._session is None:
            self._session = _lib.Session(self.selector.tree_pos, k_hint=16)
        return self._session

    def denoiseOurs(self, iterations: int = 2, original_pos: torch.Tensor | None = None, strategy=None, alphas=(1.0, 0.2, 1.0)):
        """The thesis' final strategy as the notebook runs it (PostProcessing.ipynb#c9, row "Ours"; not a method of the
        reference's Processor): per iteration flat_step for class 0 and feature_step for classes 1 and 2 with d * 20000, all
        from one snapshot of the positions, then the displacement clamp |x_new - x_original| < d.  One fused session step per
        iteration (NGPD_STEP_SNAPSHOT_CLASSES + clamp_radius).  strategy: Denoiser step per class, default flat/feature/feature."""
        g = self.graph
        sess = self._get_session()
        sess.set_state(g.pos, g.n)
        sess.set_original(g.pos if original_pos is None else original_pos)
        d = 2.0 * self._mean_edge_length(6)
        dn = self.denoiser
        funcs = strategy if strategy is not None else (dn.flat_step, dn.feature_step, dn.feature_step)
        kinds = [self._kind_of(f) for f in funcs]
        assert None not in kinds, "denoiseOurs: strategy entries must be Denoiser steps"
        params = _lib.make_params(16, 8, None, 0.3, 3.0, 0.2, kinds, alphas, d * 20000.0, _lib.STEP_SNAPSHOT_CLASSES, d)
        for _ in range(iterations):
            sess.step(params)
        pos, nrm, _ = sess.get_state(False)
        sess.set_original(None)
        g.pos.copy_(pos)
        g.n = nrm

    def _kind_of(self, func) -> int | None:
        d = self.denoiser
        table = {d.flat_step: _lib.STEP_FLAT, d.edge_step: _lib.STEP_EDGE, d.feature_step: _lib.STEP_FEATURE,
                 d.corner_step: _lib.STEP_CORNER, d.dummy_step: _lib.STEP_NONE}
        return table.get(func)

    def _mean_edge_length(self, k: int) -> float:
        s, c = self._get_session().mean_edge_length_parts(k)
        return float(torch.tensor(s / c, dtype=torch.float32))

    # ---- reference interface ----------------------------------------------------------------------------
    def getVUDecomposition(self):
        """(:83-99) Yadav-2018 feature decomposition at r = 2 x mean 6-NN edge length, rho = 0.95."""
        g = self.graph
        g.edge_index = self.graphBuilder.getKNNEdgeIndex(6)
        r = 2 * TorchUtils.averageEdgeLength(g.pos, g.edge_index)
        selection = self.selector.getPointsInRangeSelection(float(r))
        nvt = self.decompositionor.getNormalFilteredNVT(selection, g.n, rho=0.95)
        filtered_normals = nvt.getVUSmoothedNormals(g.n, tau=0.3, d=3)
        return self.decompositionor.getNormalFilteredPVT(selection, filtered_normals, rho=0.95)

    def getMartinFeatureDecomposition(self, r: float, rho: float = 0.9):
        """(:101-108) radius selection -> normal-filtered NVT -> eigen-space smoothing -> normal-filtered PVT."""
        n = self.graph.n
        selection = self.selector.getPointsInRangeSelection(float(r))
        nvt = self.decompositionor.getNormalFilteredNVT(selection, n, rho)
        filtered_normals = nvt.getVUSmoothedNormals(n)
        decomposition = self.decompositionor.getNormalFilteredPVT(selection, filtered_normals, rho)
        return decomposition, filtered_normals

    def getMyFeatureDecomposition(self, N: int = 2 ** 4, angle: float = None):
        """(:110-117) NVT on the current normals -> eigen-space smoothing -> NVT on the smoothed normals."""
        angle = angle if angle is not None else math.pi * 5 / 12
        n = self.graph.n
        selection = self.selector.getKNNSelection(N)
        nvt = self.decompositionor.getBetterFilteredNVT(selection, n, angle)
        filtered_normals = nvt.getVUSmoothedNormals(n)
        decomposition = self.decompositionor.getBetterFilteredNVT(selection, filtered_normals, angle)
        return decomposition, filtered_normals

    def denoise(self):
        """(:119-139) d = 2 * mean 6-NN edge length; two iterations; flat / edge / feature steps with
        alpha (1, .2, 1); positions are updated in place, graph.n becomes the smoothed normals."""
        self._denoise_fused(16, 8, 2)

    def _denoise_fused(self, k_feature: int, k_update: int, iterations: int):
        """The loop of `denoise` on the fused session with the neighbourhood sizes as parameters (the reference hard-codes
        16 / 8 / 2, :113, :126, :123; BASELINE configs[2] asks for k = 32)."""
        g = self.graph
        sess = self._get_session()
        sess.set_state(g.pos, g.n)
        d = 2.0 * self._mean_edge_length(6)
        params = _lib.make_params(k_feature, k_update, None, 0.3, 3.0, 0.2, (_lib.STEP_FLAT, _lib.STEP_EDGE, _lib.STEP_FEATURE),
                                  (1.0, 0.2, 1.0), d)
        for _ in range(iterations):
            sess.step(params)
        pos, nrm, _ = sess.get_state(False)
        g.pos.copy_(pos)
        g.n = nrm

    def denoiseUntilMinimumError(self, gt_pos: torch.Tensor, strategy: dict, k: int = 7, alpha: list = [0.02, 0.02, 0.1],
                                 d: float = 200, error_funcs: list[Callable] = [TorchUtils.PaperDistance]):
        """(:141-185) iterate while mean(error_funcs[0](gt, pos)) decreases.  Returns (positions, errors,
        iteration count) with the reference's bookkeeping: the returned positions are those of the LAST
        iteration when at least two ran (the reference's previous_pos aliases graph.pos, :174-175), the noisy input
        otherwise; graph.pos / graph.n are reset to the noisy input on exit.  A class holding exactly one point is
        updated normally (the reference raises IndexError there, :163)."""
        g = self.graph
        noisy_pos, noisy_n = g.pos.clone(), g.n.clone()
        kinds = [_lib.STEP_NONE] * 3
        # the fused session moves the classes in the order 0, 1, 2 from rows that are prefixes of the 16-NN rows: a strategy
        # dict in another order (the reference walks strategy.items(), and the classes move in place one after the other) or a
        # wider update neighbourhood goes through the public operators instead
        fused = list(strategy) == sorted(strategy) and 1 <= k <= 16
        for key, func in strategy.items():
            kind = self._kind_of(func)
            if kind is None or key not in (0, 1, 2):
                fused = False
            else:
                kinds[key] = kind
        i = 0
        previous_pos = current_pos = noisy_pos
        previous_error = [f(gt_pos, g.pos) + 200 for f in error_funcs]
        current_error = [f(gt_pos, g.pos) for f in error_funcs]
        if fused:
            sess = self._get_session()
            sess.set_state(g.pos, g.n)
            params = _lib.make_params(16, k, None, 0.3, 3.0, 0.2, kinds, [alpha[q] if q < len(alpha) else 0.0 for q in range(3)], float(d))
        while current_error[0].mean(dim=0) < previous_error[0].mean(dim=0):
            if fused:
                sess.step(params)
                pos, f_n, _ = sess.get_state(False)
                g.pos.copy_(pos)
            else:
                f_n = self._generic_iteration(strategy, k, alpha, d)
            error = [f(gt_pos, g.pos) for f in error_funcs]
            previous_error, current_error = current_error, error
            previous_pos, current_pos = current_pos, g.pos
            g.n = f_n
            i += 1
        print(f"Stopped cause new error was {current_error[0].mean(dim=0):.2E} compaired to previous error {previous_error[0].mean(dim=0):.2E}")
        g.pos = noisy_pos
        g.n = noisy_n
        return previous_pos, previous_error, i - 1

    def _generic_iteration(self, strategy: dict, k: int, alpha, d):
        """One iteration through the public operators, for strategies holding callables that are not Denoiser steps."""
        g = self.graph
        decomposition, f_n = self.getMyFeatureDecomposition()
        edge_vectors = decomposition.eigvec[..., 0]
        classes = decomposition.getClasses()
        selection = self.selector.getKNNSelection(k)
        for key, func in strategy.items():
            indices = (classes == key).nonzero().flatten()
            if indices.size(0) == 0:
                continue
            if func == self.denoiser.edge_step:
                new_pos = func(selection.filter(indices), f_n, edge_vectors, d, alpha[key])
            else:
                new_pos = func(selection.filter(indices), f_n, d, alpha[key])
            g.pos[indices] = new_pos
        return f_n

    def preprocessPointcloud(self, k: int = 12, noise_level: float = 0.3):
        """(:187-199) kNN graph + PCA normals on the clean cloud, Gaussian noise along the normal of
        sigma = noise_level * mean edge length, then normals of the noisy cloud over the clean cloud's graph,
        oriented by the spanning tree."""
        g = self.graph
        gb = self.graphBuilder
        g.edge_index = gb.getKNNEdgeIndex(k)
        gb.setAndFlipNormals(flip=False)
        l = TorchUtils.averageEdgeLength(g.pos, g.edge_index)
        self.noise.generateNoise(noise_level, l, keepNormals=False)
        gb.setAndFlipNormals(flip=True)
