"""Deterministic synthetic workloads C1..C5 of BASELINE.json (generators fixed in SURVEY.md §8(d)).

All tensors are fp32, generated on the CPU with a seeded `torch.Generator` (so the CPU oracle and
the CUDA path see bit-identical inputs) and are POST-activation parameters: `svec = exp(raw)`,
`alpha = sigmoid(raw)`, `color = sigmoid(raw)` as the reference renderer would hand them to the
rasterizer (gs/gaussian_splatting.py:113-123, utils/activations.py:38-47).

C2's "Point-E init" follows `point_e_intialize` (utils/initialize.py:110-167: 4096 seed points, the
non-duplicating padding branch :124-135, centre, scale to max-norm 1 x mean_std, `facex` rotation,
svec 0.02, alpha 0.8, identity qvec -- conf/base.yaml:27-35).  The live Point-E sample needs network
weights, and reference assets are not copied into this repo, so the 4096 seed points are a
procedural stand-in (a lumpy quadruped made of ellipsoid shells) with the same statistics.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from .camera import CameraInfo, orbit_c2w

SH_C0 = 0.28209479177387814  # gs/sh_renderer.py:34 sh_base


@dataclass
class Scene:
    name: str
    mean: torch.Tensor  # [N,3]
    qvec: torch.Tensor  # [N,4] (w,x,y,z), not necessarily unit
    svec: torch.Tensor  # [N,3] post-exp
    alpha: torch.Tensor  # [N]   post-sigmoid
    color: torch.Tensor  # [N,3] post-sigmoid RGB (RGB path)
    sh: Optional[torch.Tensor]  # [N,3,C*C] (SH path) or None
    C: int  # SH "degree+1" (reference template parameter); 1 for deg 0
    cams: List[CameraInfo] = field(default_factory=list)
    c2ws: List[torch.Tensor] = field(default_factory=list)
    seed: int = 0

    @property
    def N(self):
        return self.mean.shape[0]

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)
        return Scene(self.name, mv(self.mean), mv(self.qvec), mv(self.svec), mv(self.alpha), mv(self.color),
                     mv(self.sh), self.C, self.cams, [c.to(device) for c in self.c2ws], self.seed)


def _log_uniform(g, n, lo, hi):
    u = torch.rand(n, 3, generator=g)
    return torch.exp(math.log(lo) + u * (math.log(hi) - math.log(lo)))


def _uniform_ball(g, n, r):
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    rad = r * torch.rand(n, 1, generator=g) ** (1.0 / 3.0)
    return d * rad


def init_sh_coeffs(rgb: torch.Tensor, C: int, g: torch.Generator, noise=0.1) -> torch.Tensor:
    """gs/sh_renderer.py:38-43: sh[:,:,0] = logit(rgb)/0.2820948; higher bands 0.1*randn so that
    view dependence is exercised."""
    sh = torch.zeros(rgb.shape[0], 3, C * C)
    rgbc = rgb.clamp(1e-4, 1 - 1e-4)
    sh[:, :, 0] = torch.log(rgbc / (1 - rgbc)) / SH_C0
    if C > 1:
        sh[:, :, 1:] = noise * torch.randn(rgb.shape[0], 3, C * C - 1, generator=g)
    return sh.contiguous()


def _random_scene(name, seed, N, mean, svec_lo, svec_hi, C):
    g = torch.Generator().manual_seed(seed)
    mean = mean(g)
    svec = _log_uniform(g, N, svec_lo, svec_hi)
    qvec = torch.randn(N, 4, generator=g)
    qvec = qvec / qvec.norm(dim=-1, keepdim=True)
    alpha = 0.05 + 0.9 * torch.rand(N, generator=g)
    rgb = torch.rand(N, 3, generator=g)
    sh = init_sh_coeffs(rgb, C, g)
    return Scene(name, mean.contiguous(), qvec.contiguous(), svec.contiguous(), alpha.contiguous(), rgb.contiguous(),
                 sh, C, seed=seed)


def _procedural_quadruped(g, n=4096):
    """4096 surface points + colours: body, head, four legs, tail as ellipsoid shells in +-0.5."""
    parts = [  # centre, radii, colour, weight
        ((0.0, 0.0, 0.05), (0.32, 0.16, 0.15), (0.80, 0.55, 0.25), 0.40),
        ((0.36, 0.0, 0.22), (0.13, 0.11, 0.12), (0.85, 0.60, 0.30), 0.16),
        ((0.22, 0.09, -0.22), (0.05, 0.05, 0.16), (0.95, 0.90, 0.85), 0.09),
        ((0.22, -0.09, -0.22), (0.05, 0.05, 0.16), (0.95, 0.90, 0.85), 0.09),
        ((-0.22, 0.09, -0.22), (0.05, 0.05, 0.16), (0.95, 0.90, 0.85), 0.09),
        ((-0.22, -0.09, -0.22), (0.05, 0.05, 0.16), (0.95, 0.90, 0.85), 0.09),
        ((-0.38, 0.0, 0.16), (0.10, 0.03, 0.03), (0.70, 0.45, 0.20), 0.08),
    ]
    counts = [int(round(p[3] * n)) for p in parts]
    counts[0] += n - sum(counts)
    xyz, rgb = [], []
    for (c, r, col, _), k in zip(parts, counts):
        d = torch.randn(k, 3, generator=g)
        d = d / d.norm(dim=-1, keepdim=True)
        xyz.append(d * torch.tensor(r) + torch.tensor(c))
        rgb.append((torch.tensor(col) + 0.05 * torch.randn(k, 3, generator=g)).clamp(0.02, 0.98))
    return torch.cat(xyz), torch.cat(rgb)


def _cam(reso, focal, dist, elev, azim):
    return CameraInfo.from_reso(reso, focal, 0.01, 100.0), orbit_c2w(dist, elev, azim)


def make_scene(cfg: str, N: Optional[int] = None, reso: Optional[int] = None, svec_scale: float = 1.0) -> Scene:
    """cfg in {"c1","c2","c3","c4","c5"}; N / reso override the BASELINE sizes (for small parity cases)."""
    cfg = cfg.lower()
    if cfg == "c1":  # 10k random, 256^2, deg 0
        N = N or 10_000
        sc = _random_scene("c1", 0, N, lambda g: 0.5 * torch.randn(N, 3, generator=g), 0.005, 0.05, 1)
        cam, c2w = _cam(reso or 256, 1.0, 2.5, 15.0, 30.0)
        sc.cams, sc.c2ws = [cam], [c2w]
    elif cfg == "c2":  # 100k Point-E-init, 512^2, deg 2
        N = N or 100_000
        g = torch.Generator().manual_seed(1)
        n_seed = min(4096, N)
        xyz, rgb = _procedural_quadruped(g, 4096)
        xyz, rgb = xyz[:n_seed], rgb[:n_seed]
        if N > n_seed:  # utils/initialize.py:124-135 (non-duplicating branch)
            xyz = torch.cat([xyz, torch.randn(N - n_seed, 3, generator=g) * 0.8])
            rgb = torch.cat([rgb, torch.rand(N - n_seed, 3, generator=g)])
        xyz = xyz - xyz.mean(dim=0, keepdim=True)
        xyz = xyz / (xyz.norm(dim=-1).max() + 1e-5) * 0.8
        x, y, z = xyz.chunk(3, dim=-1)
        xyz = torch.cat([-y, x, z], dim=-1)  # facex
        C = 3
        sc = Scene("c2", xyz.contiguous(), torch.tensor([[1.0, 0, 0, 0]]).repeat(N, 1).contiguous(),
                   torch.full((N, 3), 0.02), torch.full((N,), 0.8), rgb.contiguous(), init_sh_coeffs(rgb, C, g), C,
                   seed=1)
        cam, c2w = _cam(reso or 512, 1.0, 2.5, 20.0, 45.0)
        sc.cams, sc.c2ws = [cam], [c2w]
    elif cfg == "c3":  # 1M, 1024^2, deg 3
        N = N or 1_000_000
        sc = _random_scene("c3", 2, N, lambda g: _uniform_ball(g, N, 1.0), 0.002, 0.02, 4)
        cam, c2w = _cam(reso or 1024, 1.1, 2.5, 15.0, 30.0)
        sc.cams, sc.c2ws = [cam], [c2w]
    elif cfg == "c4":  # 500k, 8 orbit views 800^2, deg 3
        N = N or 500_000
        sc = _random_scene("c4", 3, N, lambda g: _uniform_ball(g, N, 1.0), 0.002, 0.02, 4)
        g = torch.Generator().manual_seed(3003)
        for k in range(8):
            elev = -20.0 + 80.0 * torch.rand(1, generator=g).item()
            focal = 0.75 + 0.6 * torch.rand(1, generator=g).item()
            cam, c2w = _cam(reso or 800, focal, 2.5, elev, -180.0 + 45.0 * k)
            sc.cams.append(cam)
            sc.c2ws.append(c2w)
    elif cfg == "c5":  # 2M post-densify, 1600^2, deg 3
        N = N or 2_000_000
        sc = _random_scene("c5", 4, N, lambda g: _uniform_ball(g, N, 1.0), 0.001, 0.01, 4)
        cam, c2w = _cam(reso or 1600, 1.1, 2.5, 15.0, 30.0)
        sc.cams, sc.c2ws = [cam], [c2w]
    else:
        raise ValueError(f"unknown scene config {cfg!r}")
    if svec_scale != 1.0:  # C5 tile-occupancy sweep
        sc.svec = (sc.svec * svec_scale).contiguous()
    return sc


def mock_two_gaussians() -> Scene:
    """The reference's only deterministic fixture: `MockRenderer` (gs/debug.py:52-65) -- two
    Gaussians near the origin seen from (1,0,0), camera fx=961.22 fy=963.09 cx=648.38 cy=420.12
    1297x840 (gs/debug.py:384-393), up=+z."""
    mean = torch.tensor([[0.0, 0.0, 0.0], [0.1, 0.07, 0.0]])
    svec = torch.full((2, 3), 0.02)
    qvec = torch.tensor([[1.0, 0, 0, 0], [1.0, 0, 0, 0]])
    alpha = torch.tensor([0.8, 0.6])
    color = torch.tensor([[1.0, 0.2, 0.1], [0.1, 0.3, 1.0]])
    g = torch.Generator().manual_seed(7)
    sc = Scene("mock2", mean, qvec, svec, alpha, color, init_sh_coeffs(color, 2, g), 2, seed=7)
    from .camera import get_c2w_from_up_and_look_at
    import numpy as np

    c2w = torch.from_numpy(get_c2w_from_up_and_look_at(np.array([0, 0, 1.0]), np.zeros(3), np.array([1.0, 0, 0])))
    sc.cams = [CameraInfo(961.22, 963.09, 648.38, 420.12, 1297, 840, 0.01, 100.0)]
    sc.c2ws = [c2w]
    return sc
