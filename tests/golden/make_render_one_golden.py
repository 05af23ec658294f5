"""Golden vectors for the ORCHESTRATION of one view from the UNMODIFIED reference code (TEST INFRASTRUCTURE; dev
container only -- needs /root/reference):

    gs/gaussian_splatting.py   GaussianSplattingRenderer.render_one, .get_with_overrides
    gs/sh_renderer.py          SHRenderer.forward, .get_with_overrides
    gs/renderer.py             _render_with_T, _render_scalar, _render_sh, _render_sh_bg (autograd Functions),
                               project_gaussians, jacobian, project_pts
    gs/culling.py              tile_culling_aabb_count
    utils/camera.py            CameraInfo
    utils/transforms.py        qsvec2rotmat_batched

compiled out of the files with `ast` as they are and run on CPU tensors.  The `_backend` they call is an adapter
over the CPU oracle's C kernels (same names, same in-place contract as `_gs`), so what this fixture pins is exactly
what the oracle's `render_view` RESTATES: the order of operations of render_one, which tensors are masked, the
unmasked-size `ones` payload of the opacity pass, `out + T * bg`, `z_var`, the side effects on `max_radii2d` and the
densification lists, and the autograd wiring of the Functions.  (The kernels themselves are pinned separately against
the real `_gs` extension, tests/golden/g*.npz.)  Output: tests/golden/render_one_ref.npz, tests/golden/sh_forward_ref.npz.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def compile_defs(path, names, ns, cls=None):
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    found = {}
    for node in body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            for sub in ast.walk(node):
                if isinstance(sub, ast.FunctionDef):
                    sub.returns = None
                    for a in sub.args.args + sub.args.kwonlyargs:
                        a.annotation = None
                    sub.decorator_list = [d for d in sub.decorator_list
                                          if not (isinstance(d, ast.Name) and d.id == "lineprofiler")]
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            found[node.name] = ns[node.name]
    missing = [n for n in names if n not in found]
    assert not missing, missing
    return found


class OracleBackend:
    """`_gs` on CPU tensors: the oracle's C kernels behind the reference's op names and in-place contract."""

    def __init__(self, oracle):
        self.o = oracle

    @staticmethod
    def _cfg(th, tw, psx, psy, H, W, thresh):
        return dict(H=H, W=W, n_tiles_h=th, n_tiles_w=tw, psx=psx, psy=psy, thresh=thresh, tile_size=16)

    def culling_gaussian_bsphere(self, mean, qvec, svec, normal, pts, mask, thresh):
        mask.copy_(self.o.cull_bsphere(mean.detach(), svec.detach(), normal, pts, thresh))

    def tile_culling_aabb_start_end(self, tl, br, ids, start, end, depth, th, tw):
        i, s, e = self.o.tile_culling_aabb_start_end(tl, br, depth.detach(), th, tw, ids.shape[0])
        ids.copy_(i); start.copy_(s); end.copy_(e)

    def tile_based_vol_rendering_start_end_with_T(self, mean, cov, color, alpha, start, end, ids, out, topleft, tile,
                                                  th, tw, psx, psy, H, W, thresh, T):
        o, t, _ = self.o.composite_rgb_fwd(mean.detach(), cov.detach(), color.detach().contiguous(), alpha.detach(),
                                           start, end, ids, topleft, self._cfg(th, tw, psx, psy, H, W, thresh))
        out.copy_(o.view_as(out)); T.copy_(t.view_as(T))

    def tile_based_vol_rendering_backward_start_end(self, mean, cov, color, alpha, start, end, ids, out, gm, gc, gcol,
                                                    ga, grad, topleft, tile, th, tw, psx, psy, H, W, thresh):
        a, b, c, d = self.o.composite_rgb_bwd(mean.detach(), cov.detach(), color.detach().contiguous(), alpha.detach(),
                                              start, end, ids, out.detach(), grad, topleft,
                                              self._cfg(th, tw, psx, psy, H, W, thresh))
        gm.add_(a); gc.add_(b.view_as(gc)); gcol.add_(c); ga.add_(d.view_as(ga))

    def tile_based_vol_rendering_scalar(self, mean, cov, scalar, alpha, start, end, ids, out, topleft, tile, th, tw,
                                        psx, psy, H, W, thresh, T):
        o, t = self.o.composite_scalar_fwd(mean.detach(), cov.detach(), scalar.detach().contiguous(), alpha.detach(),
                                           start, end, ids, topleft, self._cfg(th, tw, psx, psy, H, W, thresh))
        out.copy_(o.view_as(out)); T.copy_(t.view_as(T))

    def tile_based_vol_rendering_scalar_backward(self, mean, cov, scalar, alpha, start, end, ids, out, gm, gc, gs, ga,
                                                 grad, topleft, tile, th, tw, psx, psy, H, W, thresh):
        a, b, c, d = self.o.composite_scalar_bwd(mean.detach(), cov.detach(), scalar.detach().contiguous(),
                                                 alpha.detach(), start, end, ids, out.detach().reshape(H, W),
                                                 grad.reshape(H, W), topleft, self._cfg(th, tw, psx, psy, H, W, thresh))
        gm.add_(a); gc.add_(b.view_as(gc)); gs.add_(c.view_as(gs)); ga.add_(d.view_as(ga))


    # ---- SH ops (render.h:86-132) ----
    def _sh_fwd(self, mean, cov, sh, alpha, start, end, ids, out, topleft, c2w, th, tw, psx, psy, H, W, C, thresh, bg):
        o, _, _ = self.o.composite_sh_fwd(mean.detach(), cov.detach(), sh.detach(), alpha.detach(), start, end, ids,
                                          topleft, c2w, C, self._cfg(th, tw, psx, psy, H, W, thresh), bg_rgb=bg)
        out.copy_(o.reshape(out.shape))

    def _sh_bwd(self, mean, cov, sh, alpha, start, end, ids, out, gm, gc, gsh, ga, grad, topleft, c2w, th, tw, psx, psy,
                H, W, C, thresh):
        a, b, c, d = self.o.composite_sh_bwd(mean.detach(), cov.detach(), sh.detach(), alpha.detach(), start, end, ids,
                                             out.detach().reshape(H, W, 3), grad.reshape(H, W, 3), topleft, c2w, C,
                                             self._cfg(th, tw, psx, psy, H, W, thresh))
        gm.add_(a); gc.add_(b.view_as(gc)); gsh.add_(c); ga.add_(d.view_as(ga))

    def tile_based_vol_rendering_sh(self, mean, cov, sh, alpha, start, end, ids, out, topleft, c2w, tile, th, tw, psx,
                                    psy, H, W, C, thresh):
        self._sh_fwd(mean, cov, sh, alpha, start, end, ids, out, topleft, c2w, th, tw, psx, psy, H, W, C, thresh, None)

    def tile_based_vol_rendering_sh_with_bg(self, mean, cov, sh, alpha, start, end, ids, out, topleft, c2w, tile, th,
                                            tw, psx, psy, H, W, C, thresh, bg_rgb):
        self._sh_fwd(mean, cov, sh, alpha, start, end, ids, out, topleft, c2w, th, tw, psx, psy, H, W, C, thresh, bg_rgb)

    def tile_based_vol_rendering_backward_sh(self, mean, cov, sh, alpha, start, end, ids, out, gm, gc, gsh, ga, grad,
                                             topleft, c2w, tile, th, tw, psx, psy, H, W, C, thresh):
        self._sh_bwd(mean, cov, sh, alpha, start, end, ids, out, gm, gc, gsh, ga, grad, topleft, c2w, th, tw, psx, psy,
                     H, W, C, thresh)

    def tile_based_vol_rendering_backward_sh_with_bg(self, mean, cov, sh, alpha, start, end, ids, out, gm, gc, gsh, ga,
                                                     grad, topleft, c2w, tile, th, tw, psx, psy, H, W, C, thresh,
                                                     bg_rgb):
        self._sh_bwd(mean, cov, sh, alpha, start, end, ids, out, gm, gc, gsh, ga, grad, topleft, c2w, th, tw, psx, psy,
                     H, W, C, thresh)


class TorchProxy:
    """`torch` as _render_sh sees it: `torch.cuda.profiler.cudart()` must not initialise CUDA on a CPU box."""
    _rt = type("RT", (), {"cudaProfilerStart": staticmethod(lambda: 0), "cudaProfilerStop": staticmethod(lambda: 0)})()
    cuda = type("Cuda", (), {"profiler": type("P", (), {"cudart": staticmethod(lambda: TorchProxy._rt)})()})()

    def __getattr__(self, name):
        return getattr(torch, name)


def main():
    import oracle
    from gsgen_b200.scenes import make_scene

    oracle.build()
    noop = lambda *a, **k: None
    ns = {"torch": torch, "np": np, "F": torch.nn.functional, "_backend": OracleBackend(oracle),
          "tic": noop, "toc": noop, "print_info": noop,
          "console": type("C", (), {"print": staticmethod(noop)})(),
          "QuaternionCoeffOrder": type("Q", (), {"WXYZ": "wxyz"}),
          "quaternion_to_rotation_matrix": lambda q, order: oracle.quat_to_rotmat(q)}
    compile_defs(f"{REF}/utils/transforms.py", ["qsvec2rotmat_batched"], ns)
    compile_defs(f"{REF}/gs/renderer.py", ["jacobian", "project_pts", "project_gaussians", "_render_with_T",
                                           "_render_scalar"], ns)
    ns["render_with_T"], ns["render_scalar"] = ns["_render_with_T"].apply, ns["_render_scalar"].apply
    compile_defs(f"{REF}/utils/camera.py", ["CameraInfo"], ns)
    compile_defs(f"{REF}/gs/culling.py", ["tile_culling_aabb_count"], ns)
    methods = compile_defs(f"{REF}/gs/gaussian_splatting.py", ["render_one", "get_with_overrides"], ns,
                           cls="GaussianSplattingRenderer")
    Host = type("Host", (), dict(methods))

    out = {}
    for tag, cfg, N, reso, scale in (("a", "c1", 1000, 96, 1.0), ("b", "c3", 1500, 80, 4.0)):
        sc = make_scene(cfg, N=N, reso=reso)
        sc.svec = (sc.svec * scale).contiguous()
        cam, c2w = sc.cams[0], sc.c2ws[0]
        rcam = ns["CameraInfo"](cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h, cam.near_plane, cam.far_plane)
        g = torch.Generator().manual_seed(31)
        H, W = cam.h, cam.w
        bg = torch.rand(H, W, 3, generator=g).requires_grad_()
        h = Host()
        h.N, h.device = N, "cpu"
        leaves = {k: v.clone().requires_grad_() for k, v in
                  dict(mean=sc.mean, qvec=sc.qvec, svec=sc.svec, color=sc.color, alpha=sc.alpha).items()}
        for k, v in leaves.items():
            setattr(h, k, v)
        h.skip_frustum_culling, h.frustum_culling_radius, h.tile_culling_radius = False, 6.0, 6.0
        h.tile_size, h.T_thresh, h.depth_detach = 16, 1e-4, True
        h.training, h.densify_enabled = True, True
        h.max_radii2d = torch.zeros(N)
        h.mean_2ds, h.masks = [], []
        h.cfg = types.SimpleNamespace(debug=False)
        h.bg = lambda rays_d: bg
        res = h.render_one(c2w, rcam, use_bg=True, rgb_only=False, return_T=True)
        w = {k: torch.randn(res[k].shape, generator=g) for k in ("rgb", "depth", "opacity", "z_var")}
        sum((res[k] * w[k]).sum() for k in w).backward()
        out.update({f"{tag}_{k}": v.detach() for k, v in res.items()})
        out.update({f"{tag}_w_{k}": v for k, v in w.items()})
        out.update({f"{tag}_in_{k}": v.detach() for k, v in leaves.items()})
        out.update({f"{tag}_grad_{k}": v.grad.clone() for k, v in leaves.items()})
        out.update({f"{tag}_in_bg": bg.detach(), f"{tag}_grad_bg": bg.grad.clone(), f"{tag}_c2w": c2w,
                    f"{tag}_cam": torch.tensor([cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h, cam.near_plane,
                                                cam.far_plane], dtype=torch.float64),
                    f"{tag}_max_radii2d": h.max_radii2d.clone(), f"{tag}_mask": h.masks[0].clone(),
                    f"{tag}_mean2d_grad": h.mean_2ds[0].grad.clone(),
                    f"{tag}_N_with_dub": torch.tensor([int(h.total_dub_gaussians)])})
        print(tag, cfg, "N", N, "visible", int(h.masks[0].sum()), "D", int(h.total_dub_gaussians))
    path = os.path.join(ROOT, "tests", "golden", "render_one_ref.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print("wrote", os.path.getsize(path), "bytes")

    # ---- SHRenderer.forward (gs/sh_renderer.py:227-361) with _render_sh / _render_sh_bg (gs/renderer.py:674-997)
    ns_sh = dict(ns, torch=TorchProxy())
    compile_defs(f"{REF}/gs/renderer.py", ["_render_sh", "_render_sh_bg"], ns_sh)
    ns_sh["render_sh"], ns_sh["render_sh_bg"] = ns_sh["_render_sh"].apply, ns_sh["_render_sh_bg"].apply
    sh_methods = compile_defs(f"{REF}/gs/sh_renderer.py", ["forward", "get_with_overrides"], ns_sh, cls="SHRenderer")
    HostSH = type("HostSH", (), dict(sh_methods))
    out = {}
    for tag, N, reso, C, with_bg in (("c", 1200, 80, 4, False), ("d", 1000, 72, 3, True)):
        sc = make_scene("c3", N=N, reso=reso)
        sc.svec = (sc.svec * 4.0).contiguous()
        cam, c2w = sc.cams[0], sc.c2ws[0]
        rcam = ns["CameraInfo"](cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h, cam.near_plane, cam.far_plane)
        g = torch.Generator().manual_seed(41 + C)
        h = HostSH()
        h.N, h.device = N, "cpu"
        leaves = {k: v.clone().requires_grad_() for k, v in
                  dict(mean=sc.mean, qvec=sc.qvec, svec=sc.svec, sh_coeffs=sc.sh, alpha=sc.alpha).items()}
        for k, v in leaves.items():
            setattr(h, k, v)
        h.skip_frustum_culling, h.frustum_culling_radius, h.tile_culling_radius = False, 6.0, 6.0
        h.tile_size, h.T_thresh, h.depth_detach = 16, 1e-4, True
        h.training, h.split_type = True, "2d_mean_grad"
        h.cnt = torch.zeros(N)
        h.cfg = types.SimpleNamespace(debug=False)
        h.now_C, h.bg = C, with_bg
        h.bg_rgb = torch.tensor([0.2, 0.5, 0.7])
        rgb = h.forward(c2w, rcam)
        w = torch.randn(rgb.shape, generator=g)
        (rgb * w).sum().backward()
        out.update({f"{tag}_rgb": rgb.detach(), f"{tag}_w": w, f"{tag}_c2w": c2w, f"{tag}_bg_rgb": h.bg_rgb,
                    f"{tag}_C": torch.tensor([C, int(with_bg)]),
                    f"{tag}_cam": torch.tensor([cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h, cam.near_plane,
                                                cam.far_plane], dtype=torch.float64),
                    f"{tag}_mask": h.frustum_culling_mask.clone(), f"{tag}_cnt": h.cnt.clone(),
                    f"{tag}_mean2d_grad": h.mean_2d.grad.clone(),
                    f"{tag}_N_with_dub": torch.tensor([int(h.total_dub_gaussians)])})
        out.update({f"{tag}_in_{k}": v.detach() for k, v in leaves.items()})
        out.update({f"{tag}_grad_{k}": v.grad.clone() for k, v in leaves.items()})
        print(tag, "SH C", C, "bg", with_bg, "N", N, "D", int(h.total_dub_gaussians))
    path = os.path.join(ROOT, "tests", "golden", "sh_forward_ref.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print("wrote", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
