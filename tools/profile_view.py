"""Profiling driver (not the bench, not the product): a few fwd+bwd steps of one BASELINE workload through the fused
path, for `ncu` captures.  `python tools/profile_view.py [c3|c4|c5] [steps]`"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gsgen_b200.parallel import ViewParallelRenderer  # noqa: E402
from gsgen_b200.rasterizer import render_view  # noqa: E402
from gsgen_b200.scenes import make_scene  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda", 0)
    sc = make_scene(wl)
    cam, c2w = sc.cams[0], sc.c2ws[0]
    vpr = ViewParallelRenderer(dict(mean=sc.mean, qvec=sc.qvec, svec=sc.svec, alpha=sc.alpha, sh=sc.sh), sc.C, dev)
    g = torch.Generator().manual_seed(sc.seed + 100)
    gout = torch.randn(cam.h, cam.w, 3, generator=g).to(dev)
    for _ in range(steps):
        vpr.zero_grad()
        out = render_view(vpr.params["mean"], vpr.params["qvec"], vpr.params["svec"], vpr.params["alpha"], c2w, cam,
                          sh=vpr.params["sh"], C=sc.C, grad_sink=vpr.grad_views)
        out["rgb"].backward(gradient=gout)
    torch.cuda.synchronize()
    print("done", wl, steps, float(vpr.flat_grad.abs().sum()))


if __name__ == "__main__":
    main()
