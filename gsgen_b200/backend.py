"""Drop-in for the reference's `_gs` pybind module (gs/src/bindings.cpp:5-82, gs/backend.py:52-65).

`from gsgen_b200.backend import _backend` gives an object with the reference's 23 function names.
The hot-path ones take the same positional arguments (tensors + scalars, outputs mutated in place,
return None -- gs/src/render.h:3-155) and forward `tensor.data_ptr()`s plus the current CUDA stream
to libgsb200.so.  The legacy / debug names (CSR-offset composites, tile-major
binning, experimental SH backwards, host debug checker) exist as attributes so that `gs/debug.py` /
`gs/benchmarks.py` import, and raise NotImplementedError when called.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import c_f32, c_u32, fptr, iptr, ptr


def _dev(t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("expected a CUDA tensor (gsgen_b200 has no CPU path)")
    return t.device


class _Backend:
    """The `_gs` function table."""

    # ---- K1 -------------------------------------------------------------------------------------
    @staticmethod
    def culling_gaussian_bsphere(mean, qvec, svec, normal, pts, mask, thresh):
        """render.h:3-5.  mask: bool[N], written in place."""
        dev = _dev(mean)
        if mask.dtype != torch.bool:
            raise RuntimeError("mask must be a bool tensor")
        L = _lib.lib()
        _lib.check(L.gsb200_culling_gaussian_bsphere(
            fptr(mean, "mean"), fptr(qvec, "qvec"), fptr(svec, "svec"), fptr(normal, "normal"), fptr(pts, "pts"),
            ptr(mask, torch.bool, "mask"), c_u32(mean.size(0)), c_f32(thresh), _lib.stream_ptr(dev)))

    # ---- K2-K4 ----------------------------------------------------------------------------------
    @staticmethod
    def tile_culling_aabb_start_end(aabb_topleft, aabb_bottomright, gaussian_ids, start, end, depth, n_tiles_h,
                                    n_tiles_w):
        """render.h:65-68.  gaussian_ids / start / end written in place."""
        dev = _dev(aabb_topleft)
        L = _lib.lib()
        _lib.check(L.gsb200_tile_culling_aabb_start_end(
            _lib.ctx(dev), iptr(aabb_topleft, "aabb_topleft"), iptr(aabb_bottomright, "aabb_bottomright"),
            iptr(gaussian_ids, "gaussian_ids"), iptr(start, "start"), iptr(end, "end"), fptr(depth, "depth"),
            c_u32(aabb_topleft.size(0)), c_u32(gaussian_ids.size(0)), c_u32(n_tiles_h), c_u32(n_tiles_w),
            _lib.stream_ptr(dev)))

    # ---- K5 / K12 -------------------------------------------------------------------------------
    @staticmethod
    def tile_based_vol_rendering_start_end_with_T(mean, cov, color, alpha, start, end, gaussian_ids, out, topleft,
                                                  tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W,
                                                  thresh, T):
        """render.h:151-155."""
        dev = _dev(mean)
        L = _lib.lib()
        _lib.check(L.gsb200_tile_based_vol_rendering_start_end_with_T(
            _lib.ctx(dev), fptr(mean, "mean"), fptr(cov, "cov"), fptr(color, "color"), fptr(alpha, "alpha"),
            iptr(start, "start"), iptr(end, "end"), iptr(gaussian_ids, "gaussian_ids"), fptr(out, "out"),
            fptr(topleft, "topleft"), c_u32(mean.size(0)), c_u32(gaussian_ids.size(0)), c_u32(tile_size),
            c_u32(n_tiles_h), c_u32(n_tiles_w), c_f32(pixel_size_x), c_f32(pixel_size_y), c_u32(H), c_u32(W),
            c_f32(thresh), fptr(T, "T"), _lib.stream_ptr(dev)))

    @staticmethod
    def tile_based_vol_rendering_start_end(mean, cov, color, alpha, start, end, gaussian_ids, out, topleft, tile_size,
                                           n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh):
        """render.h:70-76 (K5 without the T output)."""
        _Backend.tile_based_vol_rendering_start_end_with_T(
            mean, cov, color, alpha, start, end, gaussian_ids, out, topleft, tile_size, n_tiles_h, n_tiles_w,
            pixel_size_x, pixel_size_y, H, W, thresh, None)

    # ---- K6 -------------------------------------------------------------------------------------
    @staticmethod
    def tile_based_vol_rendering_backward_start_end(mean, cov, color, alpha, start, end, gaussian_ids, out, grad_mean,
                                                    grad_cov, grad_color, grad_alpha, grad_out, topleft, tile_size,
                                                    n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh):
        """render.h:78-84.  grad_* accumulated in place."""
        dev = _dev(mean)
        L = _lib.lib()
        _lib.check(L.gsb200_tile_based_vol_rendering_backward_start_end(
            _lib.ctx(dev), fptr(mean, "mean"), fptr(cov, "cov"), fptr(color, "color"), fptr(alpha, "alpha"),
            iptr(start, "start"), iptr(end, "end"), iptr(gaussian_ids, "gaussian_ids"), fptr(out, "out"),
            fptr(grad_mean, "grad_mean"), fptr(grad_cov, "grad_cov"), fptr(grad_color, "grad_color"),
            fptr(grad_alpha, "grad_alpha"), fptr(grad_out, "grad_out"), fptr(topleft, "topleft"),
            c_u32(mean.size(0)), c_u32(gaussian_ids.size(0)), c_u32(tile_size), c_u32(n_tiles_h), c_u32(n_tiles_w),
            c_f32(pixel_size_x), c_f32(pixel_size_y), c_u32(H), c_u32(W), c_f32(thresh), _lib.stream_ptr(dev)))

    # ---- K7 / K8 --------------------------------------------------------------------------------
    @staticmethod
    def tile_based_vol_rendering_scalar(mean, cov, scalar, alpha, start, end, gaussian_ids, out, topleft, tile_size,
                                        n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh, T):
        """render.h:134-141."""
        dev = _dev(mean)
        L = _lib.lib()
        _lib.check(L.gsb200_tile_based_vol_rendering_scalar(
            _lib.ctx(dev), fptr(mean, "mean"), fptr(cov, "cov"), fptr(scalar, "scalar"), fptr(alpha, "alpha"),
            iptr(start, "start"), iptr(end, "end"), iptr(gaussian_ids, "gaussian_ids"), fptr(out, "out"),
            fptr(topleft, "topleft"), c_u32(mean.size(0)), c_u32(gaussian_ids.size(0)), c_u32(tile_size),
            c_u32(n_tiles_h), c_u32(n_tiles_w), c_f32(pixel_size_x), c_f32(pixel_size_y), c_u32(H), c_u32(W),
            c_f32(thresh), fptr(T, "T"), _lib.stream_ptr(dev)))

    @staticmethod
    def tile_based_vol_rendering_scalar_backward(mean, cov, scalar, alpha, start, end, gaussian_ids, out, grad_mean,
                                                 grad_cov, grad_scalar, grad_alpha, grad_out, topleft, tile_size,
                                                 n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh):
        """render.h:143-149."""
        dev = _dev(mean)
        L = _lib.lib()
        _lib.check(L.gsb200_tile_based_vol_rendering_scalar_backward(
            _lib.ctx(dev), fptr(mean, "mean"), fptr(cov, "cov"), fptr(scalar, "scalar"), fptr(alpha, "alpha"),
            iptr(start, "start"), iptr(end, "end"), iptr(gaussian_ids, "gaussian_ids"), fptr(out, "out"),
            fptr(grad_mean, "grad_mean"), fptr(grad_cov, "grad_cov"), fptr(grad_scalar, "grad_scalar"),
            fptr(grad_alpha, "grad_alpha"), fptr(grad_out, "grad_out"), fptr(topleft, "topleft"),
            c_u32(mean.size(0)), c_u32(gaussian_ids.size(0)), c_u32(tile_size), c_u32(n_tiles_h), c_u32(n_tiles_w),
            c_f32(pixel_size_x), c_f32(pixel_size_y), c_u32(H), c_u32(W), c_f32(thresh), _lib.stream_ptr(dev)))

    # ---- K9-K11 ---------------------------------------------------------------------------------
    @staticmethod
    def _sh_fwd(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w, tile_size, n_tiles_h,
                n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh, bg_rgb):
        dev = _dev(mean)
        if c2w.numel() < 9:
            raise RuntimeError("c2w must hold at least 9 floats")
        L = _lib.lib()
        _lib.check(L.gsb200_tile_based_vol_rendering_sh(
            _lib.ctx(dev), fptr(mean, "mean"), fptr(cov, "cov"), fptr(sh_coeffs, "sh_coeffs"), fptr(alpha, "alpha"),
            iptr(start, "start"), iptr(end, "end"), iptr(gaussian_ids, "gaussian_ids"), fptr(out, "out"),
            fptr(topleft, "topleft"), fptr(c2w, "c2w"), c_u32(mean.size(0)), c_u32(gaussian_ids.size(0)),
            c_u32(tile_size), c_u32(n_tiles_h), c_u32(n_tiles_w), c_f32(pixel_size_x), c_f32(pixel_size_y), c_u32(H),
            c_u32(W), c_u32(C), c_f32(thresh), fptr(bg_rgb, "bg_rgb"), _lib.stream_ptr(dev)))

    @staticmethod
    def _sh_bwd(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov, grad_sh_coeffs,
                grad_alpha, grad_out, topleft, c2w, tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W,
                C, thresh, bg_rgb):
        dev = _dev(mean)
        L = _lib.lib()
        _lib.check(L.gsb200_tile_based_vol_rendering_backward_sh(
            _lib.ctx(dev), fptr(mean, "mean"), fptr(cov, "cov"), fptr(sh_coeffs, "sh_coeffs"), fptr(alpha, "alpha"),
            iptr(start, "start"), iptr(end, "end"), iptr(gaussian_ids, "gaussian_ids"), fptr(out, "out"),
            fptr(grad_mean, "grad_mean"), fptr(grad_cov, "grad_cov"), fptr(grad_sh_coeffs, "grad_sh_coeffs"),
            fptr(grad_alpha, "grad_alpha"), fptr(grad_out, "grad_out"), fptr(topleft, "topleft"), fptr(c2w, "c2w"),
            c_u32(mean.size(0)), c_u32(gaussian_ids.size(0)), c_u32(tile_size), c_u32(n_tiles_h), c_u32(n_tiles_w),
            c_f32(pixel_size_x), c_f32(pixel_size_y), c_u32(H), c_u32(W), c_u32(C), c_f32(thresh),
            fptr(bg_rgb, "bg_rgb"), _lib.stream_ptr(dev)))

    @staticmethod
    def tile_based_vol_rendering_sh(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w,
                                    tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh):
        """render.h:86-93."""
        _Backend._sh_fwd(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w, tile_size,
                         n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh, None)

    @staticmethod
    def tile_based_vol_rendering_backward_sh(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, grad_mean,
                                             grad_cov, grad_sh_coeffs, grad_alpha, grad_out, topleft, c2w, tile_size,
                                             n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh):
        """render.h:95-101."""
        _Backend._sh_bwd(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov,
                         grad_sh_coeffs, grad_alpha, grad_out, topleft, c2w, tile_size, n_tiles_h, n_tiles_w,
                         pixel_size_x, pixel_size_y, H, W, C, thresh, None)

    @staticmethod
    def tile_based_vol_rendering_sh_with_bg(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w,
                                            tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C,
                                            thresh, bg_rgb):
        """render.h:119-124."""
        _Backend._sh_fwd(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w, tile_size,
                         n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh, bg_rgb)

    @staticmethod
    def tile_based_vol_rendering_backward_sh_with_bg(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out,
                                                     grad_mean, grad_cov, grad_sh_coeffs, grad_alpha, grad_out,
                                                     topleft, c2w, tile_size, n_tiles_h, n_tiles_w, pixel_size_x,
                                                     pixel_size_y, H, W, C, thresh, bg_rgb):
        """render.h:126-132."""
        _Backend._sh_bwd(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov,
                         grad_sh_coeffs, grad_alpha, grad_out, topleft, c2w, tile_size, n_tiles_h, n_tiles_w,
                         pixel_size_x, pixel_size_y, H, W, C, thresh, bg_rgb)


def _legacy(name):
    def f(*args, **kwargs):
        raise NotImplementedError(
            f"_gs.{name} is a legacy / debug entry point of the reference that the production path never calls "
            "(SURVEY.md §2.2 K13-K16); gsgen_b200 implements the start/end hot path only")

    f.__name__ = name
    return staticmethod(f)


for _n in ("count_num_gaussians_each_tile", "count_num_gaussians_each_tile_bcircle", "prepare_image_sort",
           "image_sort", "tile_based_vol_rendering", "tile_based_vol_rendering_v1", "tile_based_vol_rendering_v2",
           "tile_based_vol_rendering_backward", "tile_culling_aabb", "tile_based_vol_rendering_backward_sh_v1",
           "tile_based_vol_rendering_backward_sh_warp_reduce", "debug_check_tiledepth"):
    setattr(_Backend, _n, _legacy(_n))

_backend = _Backend()

REFERENCE_NAMES = [
    "culling_gaussian_bsphere", "count_num_gaussians_each_tile", "count_num_gaussians_each_tile_bcircle",
    "prepare_image_sort", "image_sort", "tile_based_vol_rendering", "tile_based_vol_rendering_backward",
    "debug_check_tiledepth", "tile_culling_aabb", "tile_based_vol_rendering_v1", "tile_based_vol_rendering_v2",
    "tile_culling_aabb_start_end", "tile_based_vol_rendering_start_end",
    "tile_based_vol_rendering_backward_start_end", "tile_based_vol_rendering_sh",
    "tile_based_vol_rendering_backward_sh", "tile_based_vol_rendering_backward_sh_v1",
    "tile_based_vol_rendering_backward_sh_warp_reduce", "tile_based_vol_rendering_sh_with_bg",
    "tile_based_vol_rendering_backward_sh_with_bg", "tile_based_vol_rendering_scalar",
    "tile_based_vol_rendering_scalar_backward", "tile_based_vol_rendering_start_end_with_T",
]
