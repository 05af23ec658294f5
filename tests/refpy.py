"""Harness that runs the reference's OWN Python for the hot path (oracle/_ref/ref_py.bin: code objects compiled from
/root/reference by oracle/build_ref_py.py) over a pluggable `_backend`:

  * `_gs` (the unmodified reference CUDA extension, oracle/_ref/_gs.so)            -- the reference as it ships
  * `gsgen_b200.backend._backend` (ctypes -> libgsb200.so)                         -- the drop-in (INTEGRATION.md level 1)
  * the CPU oracle behind the `_gs` names (tests/golden/make_render_one_golden.py)  -- CPU check of this harness

Only the modules the reference imports at file top that are absent from this image are stood in for: `console`,
`tic` / `toc` / `print_info` (logging / timers: no-ops) and kornia's `quaternion_to_rotation_matrix` (the restatement
both arms share, oracle.quat_to_rotmat).  TEST INFRASTRUCTURE."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_py.bin")


def load_entries():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import build_ref_py
    finally:
        sys.path.pop(0)
    return build_ref_py.load(REF_BIN)


class TorchNoProfiler:
    """`torch` as _render_sh sees it: its `torch.cuda.profiler.cudart().cudaProfilerStart/Stop()` brackets (profiling
    hooks, not part of the computation) are no-ops -- nothing else is intercepted."""
    _rt = type("RT", (), {"cudaProfilerStart": staticmethod(lambda: 0), "cudaProfilerStop": staticmethod(lambda: 0)})()
    cuda = type("Cuda", (), {"profiler": type("P", (), {"cudart": staticmethod(lambda: TorchNoProfiler._rt)})()})()

    def __getattr__(self, name):
        return getattr(torch, name)


class Recording:
    """`_backend` proxy: forwards every op and keeps the positional arguments of the last call of each (the test feeds
    them to the oracle's margin map to classify pixels that sit on the 1/255 skip threshold)."""

    def __init__(self, inner):
        self._inner = inner
        self.calls = {}

    def __getattr__(self, name):
        fn = getattr(self._inner, name)

        def wrapped(*args):
            self.calls[name] = args
            return fn(*args)

        return wrapped


def namespace(backend, entries=None, torch_mod=torch):
    """the globals the reference's definitions run in, with `_backend` = backend (`torch_mod`: the CPU check passes a
    proxy whose `torch.cuda.profiler.cudart()` -- called by _render_sh -- does not initialise CUDA)"""
    import oracle

    entries = entries if entries is not None else load_entries()
    if entries is None:
        return None
    noop = lambda *a, **k: None
    ns = {"torch": torch_mod, "np": np, "F": torch.nn.functional, "_backend": backend, "tic": noop, "toc": noop,
          "print_info": noop, "console": type("C", (), {"print": staticmethod(noop)})(),
          "QuaternionCoeffOrder": type("Q", (), {"WXYZ": "wxyz"}),
          "quaternion_to_rotation_matrix": lambda q, order: oracle.quat_to_rotmat(q)}
    methods = {}
    for e in entries:
        exec(e["code"], ns)  # (a definition's globals are `ns` itself: `_backend` is looked up there at call time)
        if e["cls"] is not None:
            methods.setdefault(e["cls"], {})[e["name"]] = ns.pop(e["name"])
    ns["render_with_T"], ns["render_scalar"] = ns["_render_with_T"].apply, ns["_render_scalar"].apply
    ns["render_sh"], ns["render_sh_bg"] = ns["_render_sh"].apply, ns["_render_sh_bg"].apply
    ns["Host"] = type("Host", (), dict(methods["GaussianSplattingRenderer"]))
    ns["HostSH"] = type("HostSH", (), dict(methods["SHRenderer"]))
    return ns


def ref_camera(ns, cam):
    return ns["CameraInfo"](cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h, cam.near_plane, cam.far_plane)


def run_render_one(ns, sc, cam, c2w, device, bg, weights):
    """GaussianSplattingRenderer.render_one (gs/gaussian_splatting.py:1198-1421), training mode, densification
    statistics on, per-pixel background `bg` [H,W,3]; then backward of sum_k <out_k, weights_k>.
    Returns (outputs, leaf gradients incl. bg, side effects)."""
    dev = torch.device(device)
    h = ns["Host"]()
    h.N, h.device = sc.N, dev
    leaves = {k: v.detach().clone().to(dev).requires_grad_() for k, v in
              dict(mean=sc.mean, qvec=sc.qvec, svec=sc.svec, color=sc.color, alpha=sc.alpha).items()}
    for k, v in leaves.items():
        setattr(h, k, v)
    h.skip_frustum_culling, h.frustum_culling_radius, h.tile_culling_radius = False, 6.0, 6.0
    h.tile_size, h.T_thresh, h.depth_detach = 16, 1e-4, True
    h.training, h.densify_enabled = True, True
    h.max_radii2d = torch.zeros(sc.N, device=dev)
    h.mean_2ds, h.masks = [], []
    h.cfg = types.SimpleNamespace(debug=False)
    bgl = bg.detach().clone().to(dev).requires_grad_()
    h.bg = lambda rays_d: bgl
    res = h.render_one(c2w.to(dev), ref_camera(ns, cam), use_bg=True, rgb_only=False, return_T=True)
    sum((res[k] * weights[k].to(dev)).sum() for k in weights).backward()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    outs = {k: v.detach() for k, v in res.items()}
    grads = {k: v.grad.detach().clone() for k, v in leaves.items()}
    grads["bg"] = bgl.grad.detach().clone()
    side = {"max_radii2d": h.max_radii2d.clone(), "mask": h.masks[0].clone(),
            "mean2d_grad": h.mean_2ds[0].grad.detach().clone(), "N_with_dub": int(h.total_dub_gaussians)}
    return outs, grads, side


def run_sh_forward(ns, sc, cam, c2w, device, C, with_bg, weight, bg_rgb=(0.2, 0.5, 0.7)):
    """SHRenderer.forward (gs/sh_renderer.py:227-361) over _render_sh / _render_sh_bg, then backward of <rgb, weight>."""
    dev = torch.device(device)
    h = ns["HostSH"]()
    h.N, h.device = sc.N, dev
    leaves = {k: v.detach().clone().to(dev).requires_grad_() for k, v in
              dict(mean=sc.mean, qvec=sc.qvec, svec=sc.svec, sh_coeffs=sc.sh, alpha=sc.alpha).items()}
    for k, v in leaves.items():
        setattr(h, k, v)
    h.skip_frustum_culling, h.frustum_culling_radius, h.tile_culling_radius = False, 6.0, 6.0
    h.tile_size, h.T_thresh, h.depth_detach = 16, 1e-4, True
    h.training, h.split_type = True, "2d_mean_grad"
    h.cnt = torch.zeros(sc.N, device=dev)
    h.cfg = types.SimpleNamespace(debug=False)
    h.now_C, h.bg = C, with_bg
    h.bg_rgb = torch.tensor(bg_rgb, device=dev)
    rgb = h.forward(c2w.to(dev), ref_camera(ns, cam))
    (rgb * weight.to(dev)).sum().backward()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    grads = {k: v.grad.detach().clone() for k, v in leaves.items()}
    side = {"mask": h.frustum_culling_mask.clone(), "cnt": h.cnt.clone(), "mean2d_grad": h.mean_2d.grad.detach().clone(),
            "N_with_dub": int(h.total_dub_gaussians)}
    return rgb.detach(), grads, side
