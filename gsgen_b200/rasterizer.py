"""The whole view in one autograd node: what `GaussianSplattingRenderer.render_one`
(gs/gaussian_splatting.py:1198-1421) and `SHRenderer.forward` (gs/sh_renderer.py:227-361) orchestrate
from ~60 torch kernels, 5 extension calls, 2 host syncs and 5 cudaMallocs per view becomes
`gsb200_render_forward` / `gsb200_render_backward` (include/gsb200.h Part 3):

    cull + project + AABB + count  ->  scan  ->  key emit  ->  radix sort  ->  ranges  ->
    composite (RGB + depth + opacity + depth^2 in ONE walk, or SH)            [1 host sync: N_with_dub]
    composite backward (all channels in one walk)  ->  fused projection backward

Inputs are the POST-activation parameters, exactly what `render_one` hands to the rasterizer stage
(`self.mean / qvec / svec / color / alpha`), or -- with `raw_params=True` -- the RAW leaves
(`svec_before_activation`, `alpha_before_activation`, `color_before_activation`, gs/gaussian_splatting.py:113-123)
with the shipped activations exp / sigmoid / sigmoid (conf/renderer/base.yaml:14-16) evaluated inside the front-end
kernel and their chain rule inside the projection backward: three fewer elementwise passes over the
parameters in each direction, and the gradients land directly on the optimizer's leaves.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import Gsb200Camera, Gsb200ViewGrads, Gsb200ViewIn, Gsb200ViewOut, fptr, ptr


def make_camera_struct(c2w: torch.Tensor, camera_info, *, frustum_radius=6.0, tile_radius=6.0, T_thresh=1e-4,
                       skip_frustum_culling=False, depth_detach=True) -> Gsb200Camera:
    """Host-side view description.  The frustum planes come from `CameraInfo.get_frustum`
    (utils/camera.py:260-294) evaluated on the CPU copy of c2w."""
    h = c2w.detach().to("cpu", torch.float32)[:3, :4].contiguous()  # a CUDA c2w costs one D2H sync: pass the host pose
    cam = Gsb200Camera()
    cam.c2w = (ctypes.c_float * 12)(*h.view(-1).tolist())
    cam.fx, cam.fy, cam.cx, cam.cy = camera_info.fx, camera_info.fy, camera_info.cx, camera_info.cy
    cam.W, cam.H = int(camera_info.w), int(camera_info.h)
    normals, pts = camera_info.get_frustum(h)
    cam.frustum_normals = (ctypes.c_float * 18)(*normals.reshape(-1).tolist())
    cam.frustum_pts = (ctypes.c_float * 18)(*pts.reshape(-1).tolist())
    cam.frustum_radius, cam.tile_radius, cam.T_thresh = frustum_radius, tile_radius, T_thresh
    cam.skip_frustum_culling = 1 if skip_frustum_culling else 0
    cam.depth_detach = 1 if depth_detach else 0
    return cam


class _RenderView(torch.autograd.Function):
    """Differentiable w.r.t. mean, qvec, svec, alpha, color | sh, bg.  Returns
    (rgb[H,W,3], depth[H,W,1], opacity[H,W,1], z2[H,W,1], T[H,W,1], mean2d[N,2]); mean2d carries the
    densification gradient (gaussian_splatting.py:1246-1250, :464-469)."""

    @staticmethod
    def forward(ctx, mean, qvec, svec, alpha, color, sh, bg, cam: Gsb200Camera, C, sh_c2w9, bg_rgb, rgb_only, slot,
                aux, grad_sink, act=0):
        dev = mean.device
        N = mean.shape[0]
        H, W = cam.H, cam.W
        mean, qvec, svec, alpha = mean.contiguous(), qvec.contiguous(), svec.contiguous(), alpha.contiguous()
        is_sh = sh is not None
        vin = Gsb200ViewIn()
        vin.N = N
        vin.mean, vin.qvec, vin.svec = fptr(mean, "mean"), fptr(qvec, "qvec"), fptr(svec, "svec")
        vin.alpha = fptr(alpha, "alpha")
        vin.act = int(act)
        if is_sh:
            sh = sh.contiguous()
            if sh.shape[1:] != (3, C * C):
                raise RuntimeError(f"sh must be [N,3,{C * C}] for C={C}")
            vin.sh, vin.color, vin.C = fptr(sh, "sh"), None, C
            vin.sh_c2w9 = (ctypes.c_float * 9)(*sh_c2w9)
            if bg_rgb is not None:
                bg_rgb = bg_rgb.contiguous()
            vin.bg, vin.bg_rgb = None, fptr(bg_rgb, "bg_rgb")
        else:
            color = color.contiguous()
            vin.color, vin.sh, vin.C = fptr(color, "color"), None, 1
            if bg is not None:
                bg = bg.contiguous()
            vin.bg, vin.bg_rgb = fptr(bg, "bg"), None
        extras = (not is_sh) and (not rgb_only)
        f32 = dict(device=dev, dtype=torch.float32)
        rgb = torch.empty(H, W, 3, **f32)
        T = torch.empty(H, W, 1, **f32)
        depth = torch.empty(H, W, 1, **f32) if extras else None
        opacity = torch.empty(H, W, 1, **f32) if extras else None
        z2 = torch.empty(H, W, 1, **f32) if extras else None
        mean2d = torch.empty(N, 2, **f32)
        cov2d = torch.empty(N, 2, 2, **f32)
        depthg = torch.empty(N, 1, **f32)
        mask = torch.empty(N, device=dev, dtype=torch.bool)
        radii = torch.empty(N, **f32)
        ndup = ctypes.c_int64(0)
        gen = ctypes.c_int64(0)
        vout = Gsb200ViewOut()
        vout.rgb, vout.T = fptr(rgb), fptr(T)
        vout.depth, vout.opacity, vout.z2 = fptr(depth), fptr(opacity), fptr(z2)
        vout.mean2d, vout.cov2d, vout.depthg = fptr(mean2d), fptr(cov2d), fptr(depthg)
        vout.mask, vout.radii2d = ptr(mask, torch.bool), fptr(radii)
        vout.h_num_dup = ctypes.pointer(ndup)
        vout.h_generation = ctypes.pointer(gen)
        c = _lib.ctx(dev, slot)
        _lib.check(_lib.lib().gsb200_render_forward(c, ctypes.byref(cam), ctypes.byref(vin), ctypes.byref(vout),
                                                    _lib.stream_ptr(dev)))
        ctx.save_for_backward(mean, qvec, svec, alpha, color, sh, bg, bg_rgb, rgb, depth, opacity, z2, T, mask)
        ctx.cam, ctx.C, ctx.sh_c2w9, ctx.slot, ctx.extras, ctx.aux = cam, C, sh_c2w9, slot, extras, aux
        ctx.grad_sink, ctx.act = grad_sink, int(act)
        ctx.generation = int(gen.value)  # the context holds THIS view until another forward runs on the same slot
        if aux is not None:  # N_with_dub is None in asynchronous-count mode (view_stats() reads it later)
            aux.update(mask=mask, cov2d=cov2d, depth=depthg, radii2d=radii,
                       N_with_dub=int(ndup.value) if ndup.value >= 0 else None, slot=slot)
        ctx.mark_non_differentiable(T)
        ctx.set_materialize_grads(False)  # unused outputs arrive as None instead of freshly zero-filled tensors
        zero = None
        return (rgb, depth if extras else zero, opacity if extras else zero, z2 if extras else zero, T, mean2d)

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_opacity, g_z2, g_T, g_mean2d_in):
        (mean, qvec, svec, alpha, color, sh, bg, bg_rgb, rgb, depth, opacity, z2, T, mask) = ctx.saved_tensors
        dev = mean.device
        N = mean.shape[0]
        is_sh = sh is not None
        cam = ctx.cam
        vin = Gsb200ViewIn()
        vin.N = N
        vin.mean, vin.qvec, vin.svec, vin.alpha = fptr(mean), fptr(qvec), fptr(svec), fptr(alpha)
        vin.act = ctx.act
        if is_sh:
            vin.sh, vin.color, vin.C = fptr(sh), None, ctx.C
            vin.sh_c2w9 = (ctypes.c_float * 9)(*ctx.sh_c2w9)
            vin.bg, vin.bg_rgb = None, fptr(bg_rgb)
        else:
            vin.color, vin.sh, vin.C = fptr(color), None, 1
            vin.bg, vin.bg_rgb = fptr(bg), None
        g = Gsb200ViewGrads()
        cont = lambda t: None if t is None else t.contiguous()
        g_rgb = cont(g_rgb)
        g.g_rgb, g.rgb = fptr(g_rgb), fptr(rgb)
        if ctx.extras:
            g_depth, g_opacity, g_z2 = cont(g_depth), cont(g_opacity), cont(g_z2)
            g.g_depth, g.depth = fptr(g_depth), fptr(depth)
            g.g_opacity, g.opacity = fptr(g_opacity), fptr(opacity)
            g.g_z2, g.z2 = fptr(g_z2), fptr(z2)
        g.T, g.mask = fptr(T), ptr(mask, torch.bool)
        sink = ctx.grad_sink
        if sink is not None:  # accumulate straight into the caller's buffers (views of the flat all-reduce operand)
            gm, gq, gs, ga = sink["mean"], sink["qvec"], sink["svec"], sink["alpha"]
            gcol = sink["color"] if not is_sh else None
            gsh = sink["sh"] if is_sh else None
            g.accumulate = 1
        else:
            gm, gq, gs = torch.empty_like(mean), torch.empty_like(qvec), torch.empty_like(svec)
            ga = torch.empty_like(alpha)
            gcol = torch.empty_like(color) if not is_sh else None
            gsh = torch.zeros_like(sh) if is_sh else None
            g.accumulate = 0
        gm2 = torch.empty(N, 2, device=dev, dtype=torch.float32)
        need_bg = (not is_sh) and bg is not None and ctx.needs_input_grad[6]
        gbg = torch.empty_like(bg) if need_bg else None
        g.g_mean, g.g_qvec, g.g_svec, g.g_alpha = fptr(gm), fptr(gq), fptr(gs), fptr(ga)
        g.g_color, g.g_sh, g.g_mean2d, g.g_bg = fptr(gcol), fptr(gsh), fptr(gm2), fptr(gbg)
        g.generation = ctx.generation  # fails loudly if a later forward re-used this view's context slot
        g.touched = ptr(sink["touched"], torch.uint8, "touched") if (sink is not None and "touched" in sink) else None
        c = _lib.ctx(dev, ctx.slot)
        _lib.check(_lib.lib().gsb200_render_backward(c, ctypes.byref(cam), ctypes.byref(vin), ctypes.byref(g),
                                                     _lib.stream_ptr(dev)))
        if ctx.aux is not None:  # what mean_2d.grad holds in the reference (retain_grad, :1247)
            ctx.aux["mean2d_grad"] = gm2
        if sink is not None:
            return (None,) * 6 + (gbg,) + (None,) * 9
        return (gm, gq, gs, ga, gcol, gsh, gbg) + (None,) * 9


def render_view(mean, qvec, svec, alpha, c2w, camera_info, *, color=None, sh=None, C: int = 1, bg=None, bg_rgb=None,
                rgb_only: bool = False, sh_c2w=None, frustum_radius=6.0, tile_radius=6.0, T_thresh=1e-4,
                skip_frustum_culling=False, depth_detach=True, slot: int = 0, grad_sink=None,
                raw_params: bool = False, async_count: bool = False):
    """One view through the fused path.  Returns the dict `render_one` returns
    ({"rgb","depth","opacity","z_var"} (+"T")) plus "aux" (mask, mean2d, cov2d, depth, radii2d, N_with_dub).

    grad_sink: optional dict name -> fp32 tensor (mean, qvec, svec, alpha, color | sh) shaped like the parameters;
                the backward then ADDS the view's gradients into these tensors (e.g. views of one flat buffer that is
                all-reduced once per step) and autograd receives None for them -- no zero-fill / add pass per view.

    raw_params: svec / alpha / color are the raw leaves (`*_before_activation`); exp / sigmoid / sigmoid run inside the
                kernels and the returned (or sunk) gradients are w.r.t. the raw leaves.  SH coefficients have no
                activation in the reference (sh_renderer.py:38-43), so only svec / alpha are affected on the SH path.

    slot:       index of the library context (scratch arena + the view's saved binning / splat records) to use.  A
                context holds ONE view between its forward and its backward: every view that is in flight at the same
                time (a batch rendered before one loss.backward()) needs its own slot; the backward raises if its slot
                was overwritten.
    async_count: no host wait in forward or backward (GSB200_OPT_ASYNC_COUNT, include/gsb200.h): the tile sort covers a
                capacity learnt from the slot's earlier views, aux["N_with_dub"] is None, and the exact count is polled
                later.  A view that did not fit raises `_lib.TileListOverflow` from the first forward / backward on the
                slot that sees its count -- usually its own backward -- or from `view_stats(device, slot)`, which
                waits for it: render that view again.

    color given -> RGB path (render_with_T + 3x render_scalar semantics, per-pixel bg[H,W,3]);
    sh given    -> SH path  (render_sh / render_sh_bg semantics, constant bg_rgb[3]); `sh_c2w` is the tensor the
                   reference passes as `c2w` to the SH kernels, whose first nine floats are read as rotation rows
                   (A.7 / A.9-10) -- default: the same [3,4] c2w, like sh_renderer.py:324.
    """
    if (color is None) == (sh is None):
        raise RuntimeError("give exactly one of color / sh")
    cam = make_camera_struct(c2w, camera_info, frustum_radius=frustum_radius, tile_radius=tile_radius,
                             T_thresh=T_thresh, skip_frustum_culling=skip_frustum_culling, depth_detach=depth_detach)
    sh_c2w9 = None
    if sh is not None:
        src = c2w if sh_c2w is None else sh_c2w
        sh_c2w9 = src.detach().to("cpu", torch.float32).contiguous().view(-1)[:9].tolist()
    aux = {}
    _lib.set_option(mean.device, slot, _lib.OPT_ASYNC_COUNT, 1 if async_count else 0)
    act = 0
    if raw_params:
        act = _lib.ACT_SVEC_EXP | _lib.ACT_ALPHA_SIGMOID | (_lib.ACT_COLOR_SIGMOID if sh is None else 0)
    rgb, depth, opacity, z2, T, mean2d = _RenderView.apply(mean, qvec, svec, alpha, color, sh, bg, cam, C, sh_c2w9,
                                                           bg_rgb, rgb_only or sh is not None, slot, aux, grad_sink,
                                                           act)
    out = {"rgb": rgb, "T": T}
    if sh is None and not rgb_only:
        out.update(depth=depth, opacity=opacity, z_var=z2 - depth * depth)
    aux["mean2d"] = mean2d
    out["aux"] = aux
    return out


def view_stats(device, slot: int = 0):
    """(N_with_dub, N_visible, longest tile list) of the last forward on the slot's context (synchronises)."""
    h = (ctypes.c_int64 * 3)()
    dev = torch.device(device)
    _lib.check(_lib.lib().gsb200_view_stats(_lib.ctx(dev, slot), h, _lib.stream_ptr(dev)))
    return int(h[0]), int(h[1]), int(h[2])
