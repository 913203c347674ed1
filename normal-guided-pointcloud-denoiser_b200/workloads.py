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


# ---- the same surfaces in independent chunks: any rank can generate any part of the cloud (no replicated input) -------------
CHUNK = 1 << 20


def _chunk_gen(seed: int, chunk: int, device):
    gen = torch.Generator(device=device)
    gen.manual_seed((seed * 1000003 + chunk * 7919 + 12345) & 0x7FFFFFFFFFFFFFFF)
    return gen


def surface_chunk(kind: str, n: int, chunk: int, seed: int = 1234, device="cuda", cubes: int = 17):
    """Points [chunk * CHUNK, min(n, (chunk + 1) * CHUNK)) of the cloud: (clean positions, unit normals, global ids).
    kind "creased": the cube + torus surface of creased_surface (60 % / 40 % inside every chunk, shuffled inside the chunk).
    kind "cubes": a cubes^3 lattice of small cubes (edge 1.6 / cubes, pitch 2 / cubes) -- about a fifth of the points lie
    within a k = 16 neighbourhood of a crease or a corner, which exercises edge_step / feature_step (the thesis' subject)."""
    lo, hi = chunk * CHUNK, min(n, (chunk + 1) * CHUNK)
    m = hi - lo
    gen = _chunk_gen(seed, chunk, device)
    if kind == "creased":
        nc = int(m * 0.6)
        nt = m - nc
        face = torch.randint(0, 6, (nc,), generator=gen, device=device)
        uv = torch.rand((nc, 2), generator=gen, device=device) * 2 - 1
        axis, sign = face // 2, (face % 2).float() * 2 - 1
        cube = torch.empty((nc, 3), device=device)
        cn = torch.zeros((nc, 3), device=device)
        for a in range(3):
            msk = axis == a
            o = [c for c in range(3) if c != a]
            cube[msk, a] = sign[msk]
            cube[msk, o[0]] = uv[msk, 0]
            cube[msk, o[1]] = uv[msk, 1]
            cn[msk, a] = sign[msk]
        u = torch.rand(nt, generator=gen, device=device) * (2 * math.pi)
        v = torch.rand(nt, generator=gen, device=device) * (2 * math.pi)
        R, r = 0.6, 0.25
        tor = torch.stack([(R + r * v.cos()) * u.cos() + 1.0, (R + r * v.cos()) * u.sin() + 1.0, r * v.sin()], 1)
        tn = torch.stack([v.cos() * u.cos(), v.cos() * u.sin(), v.sin()], 1)
        pos, nrm = torch.cat([cube, tor]), torch.cat([cn, tn])
    elif kind == "cubes":
        cell = torch.randint(0, cubes ** 3, (m,), generator=gen, device=device)
        face = torch.randint(0, 6, (m,), generator=gen, device=device)
        uv = torch.rand((m, 2), generator=gen, device=device) * 2 - 1
        axis, sign = face // 2, (face % 2).float() * 2 - 1
        local = torch.empty((m, 3), device=device)
        nrm = torch.zeros((m, 3), device=device)
        for a in range(3):
            msk = axis == a
            o = [c for c in range(3) if c != a]
            local[msk, a] = sign[msk]
            local[msk, o[0]] = uv[msk, 0]
            local[msk, o[1]] = uv[msk, 1]
            nrm[msk, a] = sign[msk]
        pitch = 2.0 / cubes
        centre = torch.stack([cell % cubes, (cell // cubes) % cubes, cell // (cubes * cubes)], 1).float() * pitch - 1.0 + pitch / 2
        pos = centre + local * (0.4 * pitch)
    else:
        raise ValueError(kind)
    perm = torch.randperm(m, generator=gen, device=device)
    return pos[perm].contiguous(), nrm[perm].contiguous(), (torch.arange(lo, hi, device=device, dtype=torch.long))


def noise_chunk(pos: torch.Tensor, sigma: float, chunk: int, seed: int = 99):
    gen = _chunk_gen(seed, chunk, pos.device)
    return pos + torch.randn(pos.shape, generator=gen, device=pos.device) * sigma


def chunks_of(n: int, rank: int = 0, world: int = 1) -> range:
    """the contiguous block of chunks rank `rank` of `world` generates"""
    total = (n + CHUNK - 1) // CHUNK
    return range(rank * total // world, (rank + 1) * total // world)


def surface_area(kind: str, cubes: int = 17):
    """[(fraction of the points, area)] of the surface's parts (uniform density inside a part)"""
    if kind == "creased":
        return [(0.6, 24.0), (0.4, 4 * math.pi ** 2 * 0.6 * 0.25)]
    edge = 0.8 * 2.0 / cubes
    return [(1.0, cubes ** 3 * 6 * edge * edge)]


def mean_knn_distance(kind: str, n: int, k: int = 6, cubes: int = 17) -> float:
    """Expected mean edge length of the k-NN rows INCLUDING the zero self edge (Processor.py:120 averages over the row the
    KD-tree returns, self first) for a Poisson sample of the surface: E[d_j] = Gamma(j + 1/2) / (Gamma(j) sqrt(pi rho)).
    Used instead of a measured value so that every rank derives the same noise level without seeing the whole cloud."""
    c = sum(math.gamma(j + 0.5) / math.gamma(j) for j in range(1, k)) / k / math.sqrt(math.pi)
    return sum(f * c / math.sqrt(f * n / a) for f, a in surface_area(kind, cubes))
