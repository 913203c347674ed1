"""Synthetic noise with the interface of the reference's Pointcloud/Modules/Noise.py (:23-88).  Random numbers
are drawn with torch's CPU generator, as the reference does, so that a seeded run perturbs a cloud identically."""
from __future__ import annotations

import torch


class Noise:
    def __init__(self, graph):
        self.graph = graph
        self.noise_level = self.noise_type = self.noise_direction = None

    def getGT(self):
        g = self.graph
        pos = g.gt if getattr(g, "gt", None) is not None else g.pos
        nrm = g.gt_n if getattr(g, "gt_n", None) is not None else getattr(g, "n", None)
        return pos, nrm

    def generateNoise(self, noise_level, mean_edge_length: float, noise_type: int = 0, noise_direction: int = 0,
                      keepNormals: bool = False):
        """Gaussian (type 0) or impulsive (type 1) noise of sigma = mean_edge_length*noise_level along the normal
        (direction 0) or isotropic (direction 1), added to the ground-truth positions (:33-59)."""
        for name, val in (("noise_level", noise_level), ("noise_type", noise_type), ("noise_direction", noise_direction)):
            if not 0 <= val <= 1:
                raise ValueError(f"{name} is {val}, but should be a number between 0 and 1!")
        self.noise_level, self.noise_type, self.noise_direction = noise_level, noise_type, noise_direction
        g = self.graph
        gt, _ = self.getGT()
        n = g.num_nodes
        sigma = float(mean_edge_length) * noise_level
        draws = torch.normal(torch.zeros((n, 3), dtype=torch.float), torch.full((n, 3), sigma, dtype=torch.float)).to(gt.device)
        offset = draws if noise_direction == 1 else g.n * draws[:, 0, None]
        if noise_type == 1:
            quiet = torch.randperm(n)[:int(n * (1 - noise_level))].to(gt.device)
            offset[quiet] = 0
        self.setNoise(gt + offset, keepNormals)

    def setNoise(self, noise: torch.Tensor, keepNormals: bool = False):
        g = self.graph
        g.gt, g.gt_n = self.getGT()
        g.pos = noise
        if not keepNormals and hasattr(g, "n"):
            delattr(g, "n")

    def resetNoise(self):
        g = self.graph
        if getattr(g, "gt", None) is None:
            raise ValueError("Can't reset noise if noise has never been applied")
        g.pos = g.gt
        self.noise_level = self.noise_type = self.noise_direction = None
