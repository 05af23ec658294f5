"""The differentiable-op layer: same names and call signatures as the reference's `gs/renderer.py`
(`project_gaussians` :391-421, `render_with_T` / `render_scalar` / `render_sh` / `render_sh_bg` /
`render_start_end` = `.apply` of the autograd Functions :424-1291, `step_check` :27-31), with the
bodies calling libgsb200.so through `gsgen_b200.backend._backend`.

`project_gaussians` here is ONE fused kernel forward and ONE backward (the reference runs ~15 + ~30
torch kernels materialising [N,3,3] temporaries); the composite Functions keep the reference's
save-for-backward protocol (the saved `out` includes the background term, renderer.py:1182-1192).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import c_u32, fptr
from .backend import _backend


def step_check(step, step_size, run_at_zero=False) -> bool:
    """gs/renderer.py:27-31."""
    if step_size == 0:
        return False
    return (run_at_zero or step != 0) and step % step_size == 0


def _c2w12(c2w: torch.Tensor):
    h = c2w.detach().to("cpu", torch.float32)
    if h.shape[0] < 3 or h.shape[1] < 4:
        raise RuntimeError("c2w must be [3,4] (or [4,4])")
    vals = h[:3, :4].contiguous().view(-1).tolist()
    return (ctypes.c_float * 12)(*vals)


class _ProjectGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, qvec, svec, c2w, detach_depth):
        mean, qvec, svec = mean.contiguous(), qvec.contiguous(), svec.contiguous()
        N = mean.shape[0]
        dev = mean.device
        mean2d = torch.empty(N, 2, device=dev, dtype=torch.float32)
        cov2d = torch.empty(N, 2, 2, device=dev, dtype=torch.float32)
        JW = torch.empty(N, 3, 3, device=dev, dtype=torch.float32)
        depth = torch.empty(N, 1, device=dev, dtype=torch.float32)
        h = _c2w12(c2w)
        _lib.check(_lib.lib().gsb200_project_gaussians_forward(
            fptr(mean, "mean"), fptr(qvec, "qvec"), fptr(svec, "svec"), h, c_u32(N), fptr(mean2d), fptr(cov2d),
            fptr(JW), fptr(depth), _lib.stream_ptr(dev)))
        ctx.save_for_backward(mean, qvec, svec)
        ctx.h_c2w = h
        ctx.detach_depth = bool(detach_depth)
        ctx.mark_non_differentiable(JW)
        return mean2d, cov2d, JW, depth

    @staticmethod
    def backward(ctx, g_mean2d, g_cov2d, g_JW, g_depth):
        mean, qvec, svec = ctx.saved_tensors
        N = mean.shape[0]
        dev = mean.device
        gm, gq, gs = torch.empty_like(mean), torch.empty_like(qvec), torch.empty_like(svec)
        g_mean2d = None if g_mean2d is None else g_mean2d.contiguous()
        g_cov2d = None if g_cov2d is None else g_cov2d.contiguous()
        g_depth = None if g_depth is None else g_depth.contiguous()
        _lib.check(_lib.lib().gsb200_project_gaussians_backward(
            fptr(mean), fptr(qvec), fptr(svec), ctx.h_c2w, c_u32(N), ctypes.c_int(1 if ctx.detach_depth else 0),
            fptr(g_mean2d), fptr(g_cov2d), fptr(g_depth), fptr(gm), fptr(gq), fptr(gs), _lib.stream_ptr(dev)))
        return gm, gq, gs, None, None


def project_gaussians(mean, qvec, svec, c2w, detach_depth: bool = False):
    """gs/renderer.py:391-421 -> (mean2d[N,2], cov2d[N,2,2], JW[N,3,3], depth[N,1])."""
    return _ProjectGaussians.apply(mean, qvec, svec, c2w, detach_depth)


class _render_with_T(torch.autograd.Function):
    """gs/renderer.py:1135-1283."""

    @staticmethod
    def forward(ctx, mean, cov, scalar, alpha, start, end, gaussian_ids, topleft, tile_size, n_tiles_h, n_tiles_w,
                pixel_size_x, pixel_size_y, H, W, thresh, bg):
        out = torch.zeros([H, W, 3], dtype=torch.float32, device=mean.device)
        T = torch.ones_like(out[..., :1])
        mean, cov, scalar, alpha = mean.contiguous(), cov.contiguous(), scalar.contiguous(), alpha.contiguous()
        _backend.tile_based_vol_rendering_start_end_with_T(
            mean, cov, scalar, alpha, start, end, gaussian_ids, out, topleft, tile_size, n_tiles_h, n_tiles_w,
            pixel_size_x, pixel_size_y, H, W, thresh, T)
        out = out + T * bg
        ctx.save_for_backward(mean, cov, scalar, alpha, start, end, gaussian_ids, out, topleft, T)
        ctx.const = [tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh]
        return out

    @staticmethod
    def backward(ctx, grad):
        mean, cov, color, alpha, start, end, gaussian_ids, out, topleft, T = ctx.saved_tensors
        grad = grad.contiguous()
        grad_mean, grad_cov = torch.zeros_like(mean), torch.zeros_like(cov)
        grad_color, grad_alpha = torch.zeros_like(color), torch.zeros_like(alpha)
        _backend.tile_based_vol_rendering_backward_start_end(
            mean, cov, color, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov, grad_color, grad_alpha, grad,
            topleft, *ctx.const)
        return (grad_mean, grad_cov, grad_color, grad_alpha, None, None, None, None, None, None, None, None, None,
                None, None, None, torch.nan_to_num(grad * T))


class _render_start_end(torch.autograd.Function):
    """gs/renderer.py:541-671 (no T output, no background)."""

    @staticmethod
    def forward(ctx, mean, cov, color, alpha, start, end, gaussian_ids, topleft, tile_size, n_tiles_h, n_tiles_w,
                pixel_size_x, pixel_size_y, H, W, thresh):
        out = torch.zeros([H * W * 3], dtype=torch.float32, device=mean.device)
        mean, cov, color, alpha = mean.contiguous(), cov.contiguous(), color.contiguous(), alpha.contiguous()
        _backend.tile_based_vol_rendering_start_end(
            mean, cov, color, alpha, start, end, gaussian_ids, out, topleft, tile_size, n_tiles_h, n_tiles_w,
            pixel_size_x, pixel_size_y, H, W, thresh)
        ctx.save_for_backward(mean, cov, color, alpha, start, end, gaussian_ids, out, topleft)
        ctx.const = [tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh]
        return out

    @staticmethod
    def backward(ctx, grad):
        mean, cov, color, alpha, start, end, gaussian_ids, out, topleft = ctx.saved_tensors
        grad = grad.contiguous()
        grad_mean, grad_cov = torch.zeros_like(mean), torch.zeros_like(cov)
        grad_color, grad_alpha = torch.zeros_like(color), torch.zeros_like(alpha)
        _backend.tile_based_vol_rendering_backward_start_end(
            mean, cov, color, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov, grad_color, grad_alpha, grad,
            topleft, *ctx.const)
        return (grad_mean, grad_cov, grad_color, grad_alpha) + (None,) * 12


class _render_scalar(torch.autograd.Function):
    """gs/renderer.py:999-1132."""

    @staticmethod
    def forward(ctx, mean, cov, scalar, alpha, start, end, gaussian_ids, topleft, tile_size, n_tiles_h, n_tiles_w,
                pixel_size_x, pixel_size_y, H, W, thresh, T):
        out = torch.zeros([H * W], dtype=torch.float32, device=mean.device)
        mean, cov, scalar, alpha = mean.contiguous(), cov.contiguous(), scalar.contiguous(), alpha.contiguous()
        _backend.tile_based_vol_rendering_scalar(
            mean, cov, scalar, alpha, start, end, gaussian_ids, out, topleft, tile_size, n_tiles_h, n_tiles_w,
            pixel_size_x, pixel_size_y, H, W, thresh, T)
        ctx.save_for_backward(mean, cov, scalar, alpha, start, end, gaussian_ids, out, topleft)
        ctx.const = [tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, thresh]
        return out

    @staticmethod
    def backward(ctx, grad):
        mean, cov, scalar, alpha, start, end, gaussian_ids, out, topleft = ctx.saved_tensors
        grad = grad.contiguous()
        grad_mean, grad_cov = torch.zeros_like(mean), torch.zeros_like(cov)
        grad_scalar, grad_alpha = torch.zeros_like(scalar), torch.zeros_like(alpha)
        _backend.tile_based_vol_rendering_scalar_backward(
            mean, cov, scalar, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov, grad_scalar, grad_alpha,
            grad, topleft, *ctx.const)
        return (grad_mean, grad_cov, grad_scalar, grad_alpha) + (None,) * 13


class _render_sh(torch.autograd.Function):
    """gs/renderer.py:674-830."""

    @staticmethod
    def forward(ctx, mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, topleft, c2w, tile_size, n_tiles_h,
                n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh):
        out = torch.zeros([H * W * 3], dtype=torch.float32, device=mean.device)
        mean, cov, sh_coeffs, alpha = mean.contiguous(), cov.contiguous(), sh_coeffs.contiguous(), alpha.contiguous()
        c2w = c2w.contiguous()
        _backend.tile_based_vol_rendering_sh(
            mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w, tile_size, n_tiles_h, n_tiles_w,
            pixel_size_x, pixel_size_y, H, W, C, thresh)
        ctx.save_for_backward(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w)
        ctx.const = [tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh]
        return out

    @staticmethod
    def backward(ctx, grad):
        mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w = ctx.saved_tensors
        grad = grad.contiguous()
        grad_mean, grad_cov = torch.zeros_like(mean), torch.zeros_like(cov)
        grad_sh, grad_alpha = torch.zeros_like(sh_coeffs), torch.zeros_like(alpha)
        _backend.tile_based_vol_rendering_backward_sh(
            mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov, grad_sh, grad_alpha, grad,
            topleft, c2w, *ctx.const)
        return (grad_mean, grad_cov, grad_sh, grad_alpha) + (None,) * 14


class _render_sh_bg(torch.autograd.Function):
    """gs/renderer.py:833-996 (constant background colour blended in-kernel)."""

    @staticmethod
    def forward(ctx, mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, topleft, c2w, tile_size, n_tiles_h,
                n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh, bg_rgb):
        out = torch.zeros([H * W * 3], dtype=torch.float32, device=mean.device)
        mean, cov, sh_coeffs, alpha = mean.contiguous(), cov.contiguous(), sh_coeffs.contiguous(), alpha.contiguous()
        c2w, bg_rgb = c2w.contiguous(), bg_rgb.contiguous()
        _backend.tile_based_vol_rendering_sh_with_bg(
            mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w, tile_size, n_tiles_h, n_tiles_w,
            pixel_size_x, pixel_size_y, H, W, C, thresh, bg_rgb)
        ctx.save_for_backward(mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w, bg_rgb)
        ctx.const = [tile_size, n_tiles_h, n_tiles_w, pixel_size_x, pixel_size_y, H, W, C, thresh]
        return out

    @staticmethod
    def backward(ctx, grad):
        mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, topleft, c2w, bg_rgb = ctx.saved_tensors
        grad = grad.contiguous()
        grad_mean, grad_cov = torch.zeros_like(mean), torch.zeros_like(cov)
        grad_sh, grad_alpha = torch.zeros_like(sh_coeffs), torch.zeros_like(alpha)
        _backend.tile_based_vol_rendering_backward_sh_with_bg(
            mean, cov, sh_coeffs, alpha, start, end, gaussian_ids, out, grad_mean, grad_cov, grad_sh, grad_alpha, grad,
            topleft, c2w, *ctx.const, bg_rgb)
        return (grad_mean, grad_cov, grad_sh, grad_alpha) + (None,) * 15


def render(*args, **kwargs):
    """gs/renderer.py:424-538 `_render` drives the legacy CSR-offset kernels (K13), which the production
    path never calls; use render_start_end / render_with_T."""
    raise NotImplementedError("legacy `render` (CSR offsets) is out of scope; use render_start_end / render_with_T")


render_start_end = _render_start_end.apply
render_sh = _render_sh.apply
render_sh_bg = _render_sh_bg.apply
render_scalar = _render_scalar.apply
render_with_T = _render_with_T.apply
