"""Background models of the renderer (`gs/backgrounds.py`, `GaussianSplattingRenderer.setup_bg`
gs/gaussian_splatting.py:207-218) -- the producers of the per-pixel `bg [H,W,3]` that the composite adds as `T * bg`
inside the kernel (`render_view(bg=...)`, gradient `nan_to_num(g_rgb * T)` back into the module).

Host-side mirrors with the reference's semantics, draw for draw:
  * `Background.forward(dirs)`: `get_bg` unless random augmentation is on AND the module is training, in which case
    `random.random() < random_aug_prob` decides between `get_bg` and one uniform random colour for the whole image
    (:23-35);
  * `FixedBackground`   (:38-49)  a parameter initialised from `cfg.color`, augmentation forced off;
  * `RandomBackground`  (:52-72)  training: `torch.rand(3)` per call scaled into `cfg.range`; eval: black;
  * `ConstBackground`   (:75-85)  "learned_const": a learnable colour, `cfg.initial_color`.
`MLPBackground` (:88-118) needs tinycudann's fully fused MLP + SH encoding (absent here, outside the rasterizer path) and
raises.  Random numbers come from the same sources as the reference's (`random.random()`, the default torch CPU
generator), so a seeded run draws the same colours.

Multi-GPU note (SURVEY §8(e) lists RandomBackground's RNG next to the split noise): a view's background colour only
enters that view's gradient, and gradients are all-reduced, so replicas stay identical whatever each rank draws; only a
LEARNED background (`learned_const`) holds a parameter outside the Gaussian arena, whose gradient the caller must
all-reduce like any other replicated parameter.
"""
from __future__ import annotations

import random

import torch
import torch.nn as nn


def _get(cfg, key, default=None):
    if cfg is None:
        return default
    if hasattr(cfg, "get"):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def _expand(color: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """einops `repeat(color, "c -> h w c", h=H, w=W)`"""
    H, W = dirs.shape[:2]
    return color.reshape(1, 1, -1).expand(H, W, color.shape[-1])


class Background(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.device = _get(cfg, "device", "cuda")
        self.random_aug = bool(_get(cfg, "random_aug", False))
        self.random_aug_prob = float(_get(cfg, "random_aug_prob", 0.0))

    def get_bg(self, dirs):
        raise NotImplementedError

    def forward(self, dirs):
        if not self.random_aug or (not self.training):
            return self.get_bg(dirs)
        if random.random() < self.random_aug_prob:
            return self.get_bg(dirs)
        return _expand(torch.rand(3).to(dirs), dirs)


class FixedBackground(Background):
    def __init__(self, cfg):
        super().__init__(cfg)
        self.bg_color = nn.Parameter(torch.tensor(_get(cfg, "color"), dtype=torch.float32))
        self.random_aug = False  # (:45-46)
        self.random_aug_prob = 0.0

    def get_bg(self, dirs):
        return _expand(self.bg_color, dirs)


class RandomBackground(Background):
    def __init__(self, cfg):
        super().__init__(cfg)
        self.range = list(_get(cfg, "range", [0.0, 1.0]))

    def get_bg(self, dirs):
        color = torch.rand(3) if self.training else torch.zeros(3)
        return _expand(color.to(dirs) * (self.range[1] - self.range[0]) + self.range[0], dirs)


class ConstBackground(Background):
    def __init__(self, cfg):
        super().__init__(cfg)
        # (the reference hands the raw list to nn.Parameter, which only works for a tensor: `initial_color` as a tensor)
        self.initial_color = torch.as_tensor(_get(cfg, "initial_color", [0.5, 0.5, 0.5]), dtype=torch.float32)
        self.bg_color = nn.Parameter(self.initial_color.clone())

    def get_bg(self, dirs):
        return _expand(self.bg_color, dirs)


def make_background(cfg) -> Background:
    """`setup_bg(cfg)` (gs/gaussian_splatting.py:207-218): cfg.type in {"random", "learned_const", "mlp", "fixed"}"""
    kind = _get(cfg, "type")
    if kind == "random":
        return RandomBackground(cfg)
    if kind == "learned_const":
        return ConstBackground(cfg)
    if kind == "fixed":
        return FixedBackground(cfg)
    if kind == "mlp":
        raise NotImplementedError("Background type mlp needs tinycudann (FullyFusedMLP + SH encoding): outside the "
                                  "rasterizer path; pass any callable(rays_d[H,W,3]) -> [H,W,3] as `background`")
    raise NotImplementedError(f"Background type {kind} not implemented")
