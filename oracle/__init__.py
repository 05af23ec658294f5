"""CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement of the reference hot path (gsgen3d/gsgen, `gs/`): frustum cull -> mask gather ->
EWA projection -> AABB tile count -> key sort -> per-tile front-to-back composite (RGB / scalar /
SH / SH+bg) forward and backward.  The composites are plain C (`oracle.c`, loaded with ctypes); the
per-Gaussian torch ops of the reference (projection, AABB count, frustum) are restated with the
same torch ops on CPU so that autograd provides their backward exactly as in the reference.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs
may import this package.  `gsgen_b200/` (the product) never does.

Third-party arithmetic absent from /root/reference: kornia==0.6.0 (requirements.txt:43),
`kornia.geometry.conversions.quaternion_to_rotation_matrix(q, QuaternionCoeffOrder.WXYZ)`, called
from utils/transforms.py:37.  kornia is not installed here and not vendored; `quat_to_rotmat`
restates its published algorithm (normalise with eps 1e-12, then the standard unit-quaternion
matrix).  No reference artefact pins that boundary ("parity unpinned" for non-unit quaternions).
pytorch3d (unpinned, optional import at utils/ops.py:7-14; not installed here, not vendored):
`pytorch3d.ops.knn_points(p1, p2, K)` behind `nearest_neighbor` / `K_nearest_neighbors` (utils/ops.py:103-134).
`knn_points` below restates its published contract by brute force (squared Euclidean distance, ascending, int64
indices); the order of equal distances is not part of that contract and is fixed here to the smaller index.
No reference artefact pins that boundary either ("parity unpinned" for tie order); everything the reference
itself computes AROUND the search (distance_to_gaussian_surface, densify_by_compatnes_with_idx, the penalties)
is pinned by executing its own code over this search (tests/golden/make_compatness_golden.py).
Everything else is pinned: the torch stages by vectors the reference's OWN functions produced
(tests/golden/make_pergaussian_golden.py executes project_gaussians / tile_culling_aabb_count /
CameraInfo unmodified; tests/test_pergaussian_golden_cpu.py), everything downstream of
(mean2d, cov2d, depth) by tests/golden/g*.npz from the real reference `_gs` extension.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

c_f = ctypes.POINTER(ctypes.c_float)
c_i = ctypes.POINTER(ctypes.c_int32)
c_u8 = ctypes.POINTER(ctypes.c_uint8)
c_i64 = ctypes.POINTER(ctypes.c_int64)


class OrcStats(ctypes.Structure):
    _fields_ = [("pairs_evaluated", ctypes.c_int64), ("pairs_blended", ctypes.c_int64), ("d_eff", ctypes.c_int64)]


def build(force: bool = False) -> str:
    """gcc-compile oracle.c -> liboracle.so (a few seconds)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src, "-o", _LIB_PATH, "-lm"]
        )
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_bin.restype = ctypes.c_int64
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu", (t.dtype, t.device)
    return ctypes.cast(t.data_ptr(), c_f)


def _i(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), c_i)


def _fn(t):
    return None if t is None else _f(t)


# ------------------------------------------------------------------------------------------------
# camera helpers (utils/camera.py:219-346, data/__init__.py:14-29)
# ------------------------------------------------------------------------------------------------
@dataclass
class Cam:
    fx: float
    fy: float
    cx: float
    cy: float
    w: int
    h: int
    near: float = 0.01
    far: float = 100.0


def look_at_c2w(up, look_at, pos) -> torch.Tensor:
    """data/__init__.py:14-29 get_c2w_from_up_and_look_at."""
    up = np.asarray(up, dtype=np.float64)
    look_at = np.asarray(look_at, dtype=np.float64)
    pos = np.asarray(pos, dtype=np.float64)
    up = up / np.linalg.norm(up)
    z = look_at - pos
    z = z / np.linalg.norm(z)
    y = -up
    x = np.cross(y, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    c2w = np.zeros([3, 4], dtype=np.float32)
    c2w[:3, 0] = x
    c2w[:3, 1] = y
    c2w[:3, 2] = z
    c2w[:3, 3] = pos
    return torch.from_numpy(c2w)


def get_frustum(cam: Cam, c2w: torch.Tensor):
    """utils/camera.py:260-294 CameraInfo.get_frustum -> (normals[6,3], pts[6,3])."""
    yfov = 2 * np.arctan(cam.h / (2 * cam.fy))
    aspect = cam.w / cam.h
    up = -c2w[:, 1]
    right = c2w[:, 0]
    lookat = c2w[:, 2]
    t = c2w[:, 3]
    half_vside = cam.far * np.tan(yfov * 0.5)
    half_hside = half_vside * aspect
    near_point = cam.near * lookat
    far_point = cam.far * lookat
    left_normal = torch.linalg.cross(far_point - half_hside * right, up)
    right_normal = torch.linalg.cross(up, far_point + half_hside * right)
    up_normal = torch.linalg.cross(far_point + half_vside * up, right)
    down_normal = torch.linalg.cross(right, far_point - half_vside * up)
    pts = torch.stack([near_point + t, far_point + t, t, t, t, t], dim=0)
    normals = torch.stack([lookat, -lookat, left_normal, right_normal, up_normal, down_normal], dim=0)
    normals = F.normalize(normals, dim=-1)
    return normals.contiguous().float(), pts.contiguous().float()


def get_rays_d(cam: Cam, c2w: torch.Tensor) -> torch.Tensor:
    """utils/camera.py:327-346 -> [H,W,3] un-normalised world ray directions."""
    xp = (torch.arange(0, cam.w, dtype=torch.float32) - cam.cx) / cam.fx
    yp = (torch.arange(0, cam.h, dtype=torch.float32) - cam.cy) / cam.fy
    xp, yp = torch.meshgrid(xp, yp, indexing="ij")
    xyz = torch.stack([xp.reshape(-1), yp.reshape(-1), torch.ones(cam.w * cam.h)], dim=-1)
    return torch.einsum("ij,bj->bi", c2w[:3, :3], xyz).reshape(cam.w, cam.h, 3).transpose(0, 1)


# ------------------------------------------------------------------------------------------------
# A.2 cull (culling.h:11-20)
# ------------------------------------------------------------------------------------------------
def cull_bsphere(mean, svec, normals, pts, thresh=6.0) -> torch.Tensor:
    N = mean.shape[0]
    mask = torch.zeros(N, dtype=torch.uint8)
    lib().orc_cull_bsphere(
        N, _f(mean.contiguous()), _f(svec.contiguous()), _f(normals), _f(pts),
        ctypes.cast(mask.data_ptr(), c_u8), ctypes.c_float(thresh),
    )
    return mask.bool()


# ------------------------------------------------------------------------------------------------
# A.3 projection (gs/renderer.py:366-421, utils/transforms.py:34-46, kornia 0.6.0)
# ------------------------------------------------------------------------------------------------
def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """kornia 0.6.0 quaternion_to_rotation_matrix(order=WXYZ): normalise (eps 1e-12), then
    R = [[1-(tyy+tzz), txy-twz, txz+twy], [txy+twz, 1-(txx+tzz), tyz-twx], [txz-twy, tyz+twx, 1-(txx+tyy)]]."""
    qn = F.normalize(q, p=2.0, dim=-1, eps=1e-12)
    w, x, y, z = torch.chunk(qn, 4, dim=-1)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = torch.tensor(1.0)
    m = torch.stack(
        (one - (tyy + tzz), txy - twz, txz + twy, txy + twz, one - (txx + tzz), tyz - twx,
         txz - twy, tyz + twx, one - (txx + tyy)), dim=-1)
    return m.view(-1, 3, 3)


@torch.no_grad()
def jacobian(u):
    """gs/renderer.py:366-378 (no_grad: J is a constant in backward)."""
    l = torch.norm(u, dim=-1)
    J = torch.zeros(u.size(0), 3, 3).to(u)
    J[..., 0, 0] = 1.0 / u[..., 2]
    J[..., 2, 0] = u[..., 0] / l
    J[..., 1, 1] = 1.0 / u[..., 2]
    J[..., 2, 1] = u[..., 1] / l
    J[..., 0, 2] = -u[..., 0] / u[..., 2] / u[..., 2]
    J[..., 1, 2] = -u[..., 1] / u[..., 2] / u[..., 2]
    J[..., 2, 2] = u[..., 2] / l
    return J


def project_pts(pts, c2w):
    """gs/renderer.py:381-388."""
    d = -c2w[..., :3, 3]
    W = torch.transpose(c2w[..., :3, :3], -1, -2)
    return torch.einsum("ij,bj->bi", W, pts + d)


def project_gaussians(mean, qvec, svec, c2w, detach_depth=True):
    """gs/renderer.py:391-421 -> (mean2d[N,2], cov2d[N,2,2], JW[N,3,3], depth[N,1])."""
    pm = project_pts(mean, c2w)
    rotmat = svec.unsqueeze(-2) * quat_to_rotmat(qvec)  # utils/transforms.py:40 (columns scaled)
    sigma = rotmat @ torch.transpose(rotmat, -1, -2)
    W = torch.transpose(c2w[:3, :3], -1, -2)
    J = jacobian(pm)
    JW = torch.einsum("bij,jk->bik", J, W)
    cov = torch.bmm(torch.bmm(JW, sigma), torch.transpose(JW, -1, -2))[..., :2, :2].contiguous()
    depth = pm[..., 2:].clone().contiguous()
    if detach_depth:
        m2 = pm[..., :2].contiguous() / depth.detach()
    else:
        m2 = pm[..., :2].contiguous() / depth
    return m2, cov, JW, depth


# ------------------------------------------------------------------------------------------------
# A.4 AABB + count (gs/culling.py:8-37, utils/camera.py:301-314)
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def tile_culling_aabb_count(mean2d, cov2d, tile_size, cam: Cam, D: float):
    aabb_x = torch.sqrt(D * cov2d[:, 0, 0])
    aabb_y = torch.sqrt(D * cov2d[:, 1, 1])
    side = torch.stack([aabb_x, aabb_y], dim=-1)
    tl = mean2d - side
    br = mean2d + side

    def to_px(p):
        p = p.clone()
        p[:, 0] = p[:, 0] * cam.fx + cam.cx
        p[:, 1] = p[:, 1] * cam.fy + cam.cy
        return p.to(torch.int32)

    tl, br = to_px(tl), to_px(br)
    tl[..., 0].clamp_(min=0, max=cam.w - 1)
    tl[..., 1].clamp_(min=0, max=cam.h - 1)
    br[..., 0].clamp_(min=0, max=cam.w - 1)
    br[..., 1].clamp_(min=0, max=cam.h - 1)
    tl = torch.div(tl, tile_size, rounding_mode="floor")
    br = torch.div(br, tile_size, rounding_mode="floor")
    n = int(torch.prod(br - tl + 1, dim=-1).sum().item())
    return n, tl.contiguous(), br.contiguous()


# ------------------------------------------------------------------------------------------------
# A.5 bin (aabb_culling.h:192-260)
# ------------------------------------------------------------------------------------------------
def tile_culling_aabb_start_end(aabb_tl, aabb_br, depth, n_tiles_h, n_tiles_w, D, return_keys=False):
    N = aabb_tl.shape[0]
    ids = torch.zeros(max(D, 1), dtype=torch.int32)
    start = torch.empty(n_tiles_h * n_tiles_w, dtype=torch.int32)
    end = torch.empty(n_tiles_h * n_tiles_w, dtype=torch.int32)
    keys = torch.zeros(max(D, 1), dtype=torch.int64) if return_keys else None
    n = lib().orc_bin(
        N, _i(aabb_tl.contiguous()), _i(aabb_br.contiguous()), _f(depth.reshape(-1).contiguous()),
        n_tiles_h, n_tiles_w, ctypes.c_int64(D), _i(ids), _i(start), _i(end),
        ctypes.cast(keys.data_ptr(), c_i64) if return_keys else None,
    )
    assert n == D, f"duplicate count mismatch {n} != {D}"  # aabb_culling.h:228 host assert
    ids = ids[:D]
    if return_keys:
        return ids, start, end, keys[:D]
    return ids, start, end


# ------------------------------------------------------------------------------------------------
# A.6 composites (C) -- raw functional forms
# ------------------------------------------------------------------------------------------------
def _geom(cfg):
    return (cfg["n_tiles_h"], cfg["n_tiles_w"], ctypes.c_float(cfg["psx"]), ctypes.c_float(cfg["psy"]),
            cfg["H"], cfg["W"], ctypes.c_float(cfg["thresh"]))


def composite_rgb_fwd(mean2d, cov2d, color, alpha, start, end, ids, topleft, cfg, want_margin=False):
    """-> out[H,W,3], T[H,W], stats dict (, margin[H,W]).  vol_render.h:994-1062."""
    H, W = cfg["H"], cfg["W"]
    out = torch.zeros(H, W, 3)
    T = torch.ones(H, W)
    margin = torch.full((H, W), 1e30) if want_margin else None
    st = OrcStats()
    lib().orc_composite_rgb_fwd(
        _f(mean2d), _f(cov2d.reshape(-1, 4)), _f(color), _f(alpha.reshape(-1)), _i(start), _i(end), _i(ids),
        _f(out), _f(T), _f(topleft), cfg["n_tiles_h"], cfg["n_tiles_w"], ctypes.c_float(cfg["psx"]),
        ctypes.c_float(cfg["psy"]), H, W, ctypes.c_float(cfg["thresh"]), _fn(margin), ctypes.byref(st))
    stats = dict(pairs_evaluated=st.pairs_evaluated, pairs_blended=st.pairs_blended, d_eff=st.d_eff)
    return (out, T, stats, margin) if want_margin else (out, T, stats)


def composite_rgb_bwd(mean2d, cov2d, color, alpha, start, end, ids, final, gout, topleft, cfg):
    N = mean2d.shape[0]
    gm, gc = torch.zeros(N, 2), torch.zeros(N, 4)
    gcol, ga = torch.zeros(N, 3), torch.zeros(N)
    lib().orc_composite_rgb_bwd(
        N, _f(mean2d), _f(cov2d.reshape(-1, 4)), _f(color), _f(alpha.reshape(-1)), _i(start), _i(end), _i(ids),
        _f(final.contiguous()), _f(gm), _f(gc), _f(gcol), _f(ga), _f(gout.contiguous()), _f(topleft), *_geom(cfg))
    return gm, gc.view(N, 2, 2), gcol, ga


def composite_scalar_fwd(mean2d, cov2d, scalar, alpha, start, end, ids, topleft, cfg):
    H, W = cfg["H"], cfg["W"]
    out, T = torch.zeros(H, W), torch.ones(H, W)
    lib().orc_composite_scalar_fwd(
        _f(mean2d), _f(cov2d.reshape(-1, 4)), _f(scalar.reshape(-1)), _f(alpha.reshape(-1)), _i(start), _i(end),
        _i(ids), _f(out), _f(T), _f(topleft), *_geom(cfg))
    return out, T


def composite_scalar_bwd(mean2d, cov2d, scalar, alpha, start, end, ids, final, gout, topleft, cfg):
    N = mean2d.shape[0]
    gm, gc = torch.zeros(N, 2), torch.zeros(N, 4)
    gs, ga = torch.zeros(scalar.reshape(-1).shape[0]), torch.zeros(N)
    lib().orc_composite_scalar_bwd(
        N, _f(mean2d), _f(cov2d.reshape(-1, 4)), _f(scalar.reshape(-1)), _f(alpha.reshape(-1)), _i(start), _i(end),
        _i(ids), _f(final.contiguous()), _f(gm), _f(gc), _f(gs), _f(ga), _f(gout.contiguous()), _f(topleft),
        *_geom(cfg))
    return gm, gc.view(N, 2, 2), gs, ga


def composite_sh_fwd(mean2d, cov2d, sh, alpha, start, end, ids, topleft, c2w9, C, cfg, bg_rgb=None,
                     want_margin=False):
    """-> out[H,W,3], T[H,W], stats (, margin).  vol_render_sh.h:171-248 / vol_render_bg.h:12-110."""
    H, W = cfg["H"], cfg["W"]
    out, T = torch.zeros(H, W, 3), torch.ones(H, W)
    margin = torch.full((H, W), 1e30) if want_margin else None
    st = OrcStats()
    c2w9 = c2w9.reshape(-1)[:9].contiguous()
    lib().orc_composite_sh_fwd(
        _f(mean2d), _f(cov2d.reshape(-1, 4)), _f(sh.contiguous()), _f(alpha.reshape(-1)), _i(start), _i(end),
        _i(ids), _f(out), _f(T), _f(topleft), _f(c2w9), cfg["n_tiles_h"], cfg["n_tiles_w"],
        ctypes.c_float(cfg["psx"]), ctypes.c_float(cfg["psy"]), H, W, C, ctypes.c_float(cfg["thresh"]),
        _fn(bg_rgb), _fn(margin), ctypes.byref(st))
    stats = dict(pairs_evaluated=st.pairs_evaluated, pairs_blended=st.pairs_blended, d_eff=st.d_eff)
    return (out, T, stats, margin) if want_margin else (out, T, stats)


def composite_sh_fwd_exact(mean2d, cov2d, sh, alpha, start, end, ids, topleft, c2w9, C, cfg, bg_rgb=None):
    """ARBITER, not a restatement: the same composite in fp64 real arithmetic on the fp32 inputs (oracle.c).
    -> out[H,W,3] f64, T[H,W] f64, margin[H,W] f64."""
    H, W = cfg["H"], cfg["W"]
    out = torch.zeros(H, W, 3, dtype=torch.float64)
    T = torch.ones(H, W, dtype=torch.float64)
    margin = torch.full((H, W), 1e30, dtype=torch.float64)
    c_d = ctypes.POINTER(ctypes.c_double)
    dp = lambda t: ctypes.cast(t.data_ptr(), c_d)
    c2w9 = c2w9.reshape(-1)[:9].contiguous()
    lib().orc_composite_sh_fwd_exact(
        _f(mean2d), _f(cov2d.reshape(-1, 4)), _f(sh.contiguous()), _f(alpha.reshape(-1)), _i(start), _i(end),
        _i(ids), dp(out), dp(T), _f(topleft), _f(c2w9), cfg["n_tiles_h"], cfg["n_tiles_w"],
        ctypes.c_float(cfg["psx"]), ctypes.c_float(cfg["psy"]), H, W, C, ctypes.c_float(cfg["thresh"]),
        _fn(bg_rgb), dp(margin))
    return out, T, margin


def composite_sh_bwd(mean2d, cov2d, sh, alpha, start, end, ids, final, gout, topleft, c2w9, C, cfg):
    N = mean2d.shape[0]
    gm, gc = torch.zeros(N, 2), torch.zeros(N, 4)
    gsh, ga = torch.zeros(N, 3, C * C), torch.zeros(N)
    c2w9 = c2w9.reshape(-1)[:9].contiguous()
    lib().orc_composite_sh_bwd(
        N, _f(mean2d), _f(cov2d.reshape(-1, 4)), _f(sh.contiguous()), _f(alpha.reshape(-1)), _i(start), _i(end),
        _i(ids), _f(final.contiguous()), _f(gm), _f(gc), _f(gsh), _f(ga), _f(gout.contiguous()), _f(topleft),
        _f(c2w9), cfg["n_tiles_h"], cfg["n_tiles_w"], ctypes.c_float(cfg["psx"]), ctypes.c_float(cfg["psy"]),
        cfg["H"], cfg["W"], C, ctypes.c_float(cfg["thresh"]))
    return gm, gc.view(N, 2, 2), gsh, ga


# ------------------------------------------------------------------------------------------------
# autograd wrappers mirroring gs/renderer.py:_render_with_T / _render_scalar / _render_sh[_bg]
# ------------------------------------------------------------------------------------------------
class _RenderWithT(torch.autograd.Function):
    """gs/renderer.py:1135-1283."""

    @staticmethod
    def forward(ctx, mean, cov, color, alpha, start, end, ids, topleft, cfg, bg):
        out, T, stats = composite_rgb_fwd(mean.detach(), cov.detach().contiguous(), color.detach().contiguous(),
                                          alpha.detach().contiguous(), start, end, ids, topleft, cfg)
        T = T.unsqueeze(-1)
        out = out + T * bg  # renderer.py:1182
        ctx.save_for_backward(mean, cov, color, alpha, start, end, ids, out, topleft, T)
        ctx.cfg = cfg
        ctx.stats = stats
        return out

    @staticmethod
    def backward(ctx, grad):
        mean, cov, color, alpha, start, end, ids, out, topleft, T = ctx.saved_tensors
        gm, gc, gcol, ga = composite_rgb_bwd(mean.detach(), cov.detach().contiguous(), color.detach().contiguous(),
                                             alpha.detach().contiguous(), start, end, ids, out,
                                             grad.contiguous(), topleft, ctx.cfg)
        return gm, gc, gcol, ga.view_as(alpha), None, None, None, None, None, torch.nan_to_num(grad * T)


class _RenderScalar(torch.autograd.Function):
    """gs/renderer.py:999-1132."""

    @staticmethod
    def forward(ctx, mean, cov, scalar, alpha, start, end, ids, topleft, cfg):
        out, T = composite_scalar_fwd(mean.detach(), cov.detach().contiguous(), scalar.detach().contiguous(),
                                      alpha.detach().contiguous(), start, end, ids, topleft, cfg)
        ctx.save_for_backward(mean, cov, scalar, alpha, start, end, ids, out, topleft)
        ctx.cfg = cfg
        return out

    @staticmethod
    def backward(ctx, grad):
        mean, cov, scalar, alpha, start, end, ids, out, topleft = ctx.saved_tensors
        gm, gc, gs, ga = composite_scalar_bwd(mean.detach(), cov.detach().contiguous(), scalar.detach().contiguous(),
                                              alpha.detach().contiguous(), start, end, ids, out,
                                              grad.contiguous(), topleft, ctx.cfg)
        return gm, gc, gs.view_as(scalar), ga.view_as(alpha), None, None, None, None, None


class _RenderSH(torch.autograd.Function):
    """gs/renderer.py:674-830 (_render_sh) and :833-996 (_render_sh_bg when bg_rgb is given)."""

    @staticmethod
    def forward(ctx, mean, cov, sh, alpha, start, end, ids, topleft, c2w, C, cfg, bg_rgb):
        out, T, stats = composite_sh_fwd(mean.detach(), cov.detach().contiguous(), sh.detach().contiguous(),
                                         alpha.detach().contiguous(), start, end, ids, topleft, c2w, C, cfg, bg_rgb)
        ctx.save_for_backward(mean, cov, sh, alpha, start, end, ids, out, topleft, c2w)
        ctx.cfg, ctx.C, ctx.stats = cfg, C, stats
        return out

    @staticmethod
    def backward(ctx, grad):
        mean, cov, sh, alpha, start, end, ids, out, topleft, c2w = ctx.saved_tensors
        gm, gc, gsh, ga = composite_sh_bwd(mean.detach(), cov.detach().contiguous(), sh.detach().contiguous(),
                                           alpha.detach().contiguous(), start, end, ids, out, grad.contiguous(),
                                           topleft, c2w, ctx.C, ctx.cfg)
        return gm, gc, gsh, ga.view_as(alpha), None, None, None, None, None, None, None, None


render_with_T = _RenderWithT.apply
render_scalar = _RenderScalar.apply
render_sh = _RenderSH.apply


# ------------------------------------------------------------------------------------------------
# whole view, as GaussianSplattingRenderer.render_one (gs/gaussian_splatting.py:1198-1421) and
# SHRenderer.forward (gs/sh_renderer.py:227-361) orchestrate it
# ------------------------------------------------------------------------------------------------
def view_cfg(cam: Cam, tile_size=16, thresh=1e-4):
    H, W = cam.h, cam.w
    th = H // tile_size + (H % tile_size > 0)
    tw = W // tile_size + (W % tile_size > 0)
    return dict(H=H, W=W, n_tiles_h=th, n_tiles_w=tw, psx=1.0 / cam.fx, psy=1.0 / cam.fy, thresh=thresh,
                tile_size=tile_size)


def render_view(mean, qvec, svec, alpha, c2w, cam: Cam, *, color=None, sh=None, C=None, bg=None, bg_rgb=None,
                rgb_only=True, depth_detach=True, frustum_radius=6.0, tile_radius=6.0, thresh=1e-4,
                sh_c2w=None):
    """Oracle for one view.  Inputs are POST-activation parameters (leaf tensors may require grad).

    color given  -> RGB path  (render_one: render_with_T + optional 3x render_scalar)
    sh given     -> SH path   (SHRenderer.forward: render_sh / render_sh_bg)
    Returns dict(rgb, [depth, opacity, z_var], aux=dict(mask, mean2d, cov2d, depth, D, ids, start, end, stats)).
    """
    assert tile_radius is not None
    cfg = view_cfg(cam, 16, thresh)
    normals, pts = get_frustum(cam, c2w)
    mask = cull_bsphere(mean.detach(), svec.detach(), normals, pts, frustum_radius)
    m, q, s, a = mean[mask].contiguous(), qvec[mask].contiguous(), svec[mask].contiguous(), alpha[mask].contiguous()
    mean2d, cov2d, JW, depth = project_gaussians(m, q, s, c2w, depth_detach)
    if mean2d.requires_grad:
        mean2d.retain_grad()
    D, tl, br = tile_culling_aabb_count(mean2d.detach(), cov2d.detach(), 16, cam, tile_radius)
    ids, start, end = tile_culling_aabb_start_end(tl, br, depth.detach(), cfg["n_tiles_h"], cfg["n_tiles_w"], D)
    topleft = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=torch.float32)
    out = {}
    H, W = cam.h, cam.w
    if color is not None:
        col = color[mask].contiguous()
        if bg is None:
            bg = torch.zeros(H, W, 3)
        rgb = render_with_T(mean2d, cov2d, col, a, start, end, ids, topleft, cfg, bg).view(H, W, 3)
        out["rgb"] = rgb
        if not rgb_only:
            dep = render_scalar(mean2d, cov2d, depth, a, start, end, ids, topleft, cfg).reshape(H, W, 1)
            ones = torch.ones_like(mean.detach()[..., 0])  # A.9-16: unmasked-size scalar
            opa = render_scalar(mean2d, cov2d, ones, a, start, end, ids, topleft, cfg).reshape(H, W, 1)
            z2 = render_scalar(mean2d, cov2d, depth * depth, a, start, end, ids, topleft, cfg).reshape(H, W, 1)
            out.update(depth=dep, opacity=opa, z_var=z2 - dep * dep)
    else:
        shm = sh[mask][..., : C * C].contiguous()
        c2w_sh = c2w if sh_c2w is None else sh_c2w
        rgb = render_sh(mean2d, cov2d, shm, a, start, end, ids, topleft, c2w_sh, C, cfg, bg_rgb).view(H, W, 3)
        out["rgb"] = rgb
    out["aux"] = dict(mask=mask, mean2d=mean2d, cov2d=cov2d, depth=depth, D=D, ids=ids, start=start, end=end,
                      aabb_tl=tl, aabb_br=br, topleft=topleft, cfg=cfg)
    return out


def knn_points(query, points, K: int, return_dist: bool = True):
    """pytorch3d.ops.knn_points(query, points, K) restated by brute force (see the module docstring): for each query
    the K points with the smallest squared distance `((p - q)**2).sum(-1)` (fp32), ascending, ties by smaller index.
    Returns (dist2 [Q,K] fp32, idx [Q,K] int64); `query=None` = the points query themselves.  Slots beyond the number
    of points hold (+inf, -1).  Same signature as gsgen_b200.knn.knn_points (it is that function's checker)."""
    pts = points.detach().to(torch.float32).cpu().numpy()
    q = pts if query is None else query.detach().to(torch.float32).cpu().numpy()
    n, nq = pts.shape[0], q.shape[0]
    idx = np.full((nq, K), -1, dtype=np.int64)
    d2 = np.full((nq, K), np.inf, dtype=np.float32)
    ar = np.arange(n)
    for j in range(nq):
        d = pts - q[j]
        d = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        if n > 4 * K:  # partial selection first: everything not larger than the K-th smallest value, ties included
            kth = np.partition(d, K - 1)[K - 1]
            cand = ar[d <= kth]
        else:
            cand = ar
        order = cand[np.lexsort((cand, d[cand]))][:K]
        idx[j, : order.shape[0]] = order
        d2[j, : order.shape[0]] = d[order]
    dev = points.device
    return (torch.from_numpy(d2).to(dev) if return_dist else None), torch.from_numpy(idx).to(dev)
