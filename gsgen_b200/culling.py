"""`gs.culling.tile_culling_aabb_count` (reference gs/culling.py:8-37) as one CUDA kernel + scan.

Same signature and return value `(N_with_dub: int, aabb_topleft[N,2] int32, aabb_bottomright[N,2] int32)`;
integer results are bit-identical to the torch op sequence of the reference on the same inputs
(separate fp32 multiply / add, truncating cast, clamp, floor division).
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import torch

from . import _lib
from ._lib import c_f32, c_u32, fptr, iptr


@torch.no_grad()
def tile_culling_aabb_count(mean, cov, tile_size: int, camera_info, D: float) -> Tuple[int, torch.Tensor, torch.Tensor]:
    mean = mean.detach().contiguous()
    cov = cov.detach().contiguous()
    N = mean.shape[0]
    dev = mean.device
    tl = torch.empty(N, 2, dtype=torch.int32, device=dev)
    br = torch.empty(N, 2, dtype=torch.int32, device=dev)
    total = ctypes.c_int64(0)
    _lib.check(_lib.lib().gsb200_tile_culling_aabb_count(
        _lib.ctx(dev), fptr(mean, "mean"), fptr(cov, "cov"), c_u32(N), c_u32(tile_size), c_f32(camera_info.fx),
        c_f32(camera_info.fy), c_f32(camera_info.cx), c_f32(camera_info.cy), c_u32(camera_info.w),
        c_u32(camera_info.h), c_f32(D), iptr(tl), iptr(br), ctypes.byref(total), _lib.stream_ptr(dev)))
    return int(total.value), tl, br
