"""Synthetic workloads of BASELINE.json (configs 3-5): a noisy CAD-like surface made of the faces of the cube
[-1,1]^3 (60 % of the points, area-uniform) and a torus R=.6, r=.25 centred on a cube edge (40 %), with analytic
normals, shuffled.  Generated with torch on whichever device is asked for (input generation is plumbing, not the
measured path)."""
from __future__ import annotations

import math

import torch


def creased_surface(n: int, seed: int = 1234, device="cuda"):
    """(positions [n,3] fp32, unit normals [n,3] fp32) of the clean surface, shuffled."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    nc = int(n * 0.6)
    nt = n - nc
    face = torch.randint(0, 6, (nc,), generator=gen, device=device)
    uv = torch.rand((nc, 2), generator=gen, device=device) * 2 - 1
    axis = face // 2
    sign = (face % 2).float() * 2 - 1
    cube = torch.empty((nc, 3), device=device)
    cn = torch.zeros((nc, 3), device=device)
    for a in range(3):
        m = axis == a
        o = [c for c in range(3) if c != a]
        cube[m, a] = sign[m]
        cube[m, o[0]] = uv[m, 0]
        cube[m, o[1]] = uv[m, 1]
        cn[m, a] = sign[m]
    u = torch.rand(nt, generator=gen, device=device) * (2 * math.pi)
    v = torch.rand(nt, generator=gen, device=device) * (2 * math.pi)
    R, r = 0.6, 0.25
    tor = torch.stack([(R + r * v.cos()) * u.cos() + 1.0, (R + r * v.cos()) * u.sin() + 1.0, r * v.sin()], 1)
    tn = torch.stack([v.cos() * u.cos(), v.cos() * u.sin(), v.sin()], 1)
    pos = torch.cat([cube, tor])
    nrm = torch.cat([cn, tn])
    perm = torch.randperm(n, generator=gen, device=device)
    return pos[perm].contiguous(), nrm[perm].contiguous()


def add_noise(pos: torch.Tensor, sigma: float, seed: int = 99):
    """isotropic Gaussian displacement of standard deviation sigma (config 4: random direction)"""
    gen = torch.Generator(device=pos.device)
    gen.manual_seed(seed)
    return pos + torch.randn(pos.shape, generator=gen, device=pos.device) * sigma


def expected_spacing(n: int) -> float:
    """rough mean nearest-neighbour spacing of creased_surface(n): 1/sqrt(mean density)"""
    area = 24.0 + 4 * math.pi ** 2 * 0.6 * 0.25
    return math.sqrt(area / n)
