"""Measurement tool (not bench.py, not the product): kernel-for-kernel timing of the UNMODIFIED reference `_gs`
CUDA extension (oracle/_ref/_gs.so, built by oracle/build_ref.sh for sm_100) against libgsb200.so on the same
B200 and the same tensors, for BASELINE configs C2 (100k / 512^2 / SH deg 2) and C3 (1M / 1024^2 / SH deg 3).

    gpurun -- python tools/compare_reference_ext.py > gpurun_out/vs_reference_ext.json

Each op is timed with CUDA events over `reps` launches after warm-up (median).  The reference launches on the
legacy default stream and cudaMallocs inside its binning op; that is part of what it costs.
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from gsgen_b200.backend import _backend  # noqa: E402
from gsgen_b200.culling import tile_culling_aabb_count  # noqa: E402
from gsgen_b200.rasterizer import render_view  # noqa: E402
from gsgen_b200.renderer import project_gaussians  # noqa: E402
from gsgen_b200.scenes import make_scene  # noqa: E402

DEV = "cuda"


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def run(cfg, ref):
    sc = make_scene(cfg).to(DEV)
    cam, c2w = sc.cams[0], sc.c2ws[0]
    C = sc.C
    H, W = cam.h, cam.w
    th, tw = cam.n_tiles
    normals, pts = cam.get_frustum(c2w)
    out = {"config": f"{cfg}: N={sc.N}, {W}x{H}, SH C={C}"}
    mask_o = torch.zeros(sc.N, dtype=torch.bool, device=DEV)
    mask_r = torch.zeros(sc.N, dtype=torch.bool, device=DEV)
    out["cull_ms"] = {
        "ours": timeit(lambda: _backend.culling_gaussian_bsphere(sc.mean, sc.qvec, sc.svec, normals, pts, mask_o, 6.0)),
        "reference": timeit(lambda: ref.culling_gaussian_bsphere(sc.mean, sc.qvec, sc.svec, normals, pts, mask_r, 6.0))}
    m, q, s = sc.mean[mask_o].contiguous(), sc.qvec[mask_o].contiguous(), sc.svec[mask_o].contiguous()
    al, sh = sc.alpha[mask_o].contiguous(), sc.sh[mask_o].contiguous()
    m2, c2, _, dp = project_gaussians(m, q, s, c2w, True)
    m2, c2, dp = m2.contiguous(), c2.contiguous(), dp.contiguous()
    D, tl, br = tile_culling_aabb_count(m2, c2, 16, cam, 6.0)
    out["N_with_dub"] = D
    mk = lambda: (torch.zeros(D, dtype=torch.int32, device=DEV), -torch.ones(th * tw, dtype=torch.int32, device=DEV),
                  -torch.ones(th * tw, dtype=torch.int32, device=DEV))
    ids, start, end = mk()
    rids, rstart, rend = mk()
    out["bin_sort_ms"] = {
        "ours": timeit(lambda: _backend.tile_culling_aabb_start_end(tl, br, ids, start, end, dp, th, tw)),
        "reference": timeit(lambda: ref.tile_culling_aabb_start_end(tl, br, rids, rstart, rend, dp, th, tw))}
    assert torch.equal(start, rstart) and torch.equal(end, rend)
    topleft = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], device=DEV)
    common = (16, th, tw, 1.0 / cam.fx, 1.0 / cam.fy, H, W)
    o, ro = torch.zeros(H * W * 3, device=DEV), torch.zeros(H * W * 3, device=DEV)
    out["sh_composite_fwd_ms"] = {
        "ours": timeit(lambda: _backend.tile_based_vol_rendering_sh(m2, c2, sh, al, start, end, rids, o, topleft, c2w,
                                                                    *common, C, 1e-4)),
        "reference": timeit(lambda: ref.tile_based_vol_rendering_sh(m2, c2, sh, al, start, end, rids, ro, topleft, c2w,
                                                                    *common, C, 1e-4))}
    out["fwd_max_abs_diff"] = float((o - ro).abs().max())
    g = torch.Generator().manual_seed(1)
    go = torch.randn(H * W * 3, generator=g).to(DEV)
    z = lambda: (torch.zeros_like(m2), torch.zeros_like(c2), torch.zeros_like(sh), torch.zeros_like(al))
    ga, gb = z(), z()
    out["sh_composite_bwd_ms"] = {
        "ours": timeit(lambda: _backend.tile_based_vol_rendering_backward_sh(m2, c2, sh, al, start, end, rids, ro, *ga,
                                                                             go, topleft, c2w, *common, C, 1e-4)),
        "reference": timeit(lambda: ref.tile_based_vol_rendering_backward_sh(m2, c2, sh, al, start, end, rids, ro, *gb,
                                                                             go, topleft, c2w, *common, C, 1e-4),
                            reps=5, warm=1)}
    # RGB + 3 scalar passes (what render_one does) vs the fused RGB walk
    col = sc.color[mask_o].contiguous()
    ro3, rT = torch.zeros(H, W, 3, device=DEV), torch.ones(H, W, 1, device=DEV)
    rs, rsT = torch.zeros(H * W, device=DEV), torch.ones(H, W, 1, device=DEV)
    ones = torch.ones_like(al)

    def ref_rgb_and_scalars():
        ref.tile_based_vol_rendering_start_end_with_T(m2, c2, col, al, start, end, rids, ro3, topleft, *common, 1e-4, rT)
        for payload in (dp, ones, dp):
            ref.tile_based_vol_rendering_scalar(m2, c2, payload, al, start, end, rids, rs, topleft, *common, 1e-4, rsT)

    out["rgb_plus_3_scalar_fwd_ms"] = {"reference_4_walks": timeit(ref_rgb_and_scalars, reps=5, warm=1)}
    with torch.no_grad():
        out["whole_view_fused_rgb_fwd_ms"] = {
            "ours (cull+project+bin+sort+composite with depth/opacity/z2)": timeit(
                lambda: render_view(sc.mean, sc.qvec, sc.svec, sc.alpha, c2w.cpu(), cam, color=sc.color, rgb_only=False))}
        out["whole_view_fused_sh_fwd_ms"] = {
            "ours (cull+project+bin+sort+SH composite)": timeit(
                lambda: render_view(sc.mean, sc.qvec, sc.svec, sc.alpha, c2w.cpu(), cam, sh=sc.sh, C=C))}
    for k, v in out.items():
        if isinstance(v, dict) and "ours" in v and "reference" in v:
            v["speedup"] = v["reference"] / v["ours"]
    return out


def main():
    import _gs as ref  # the reference extension

    res = {"device": torch.cuda.get_device_name(0), "note": "reference = unmodified gsgen `_gs` ext rebuilt for sm_100 "
           "(-DNDEBUG); ours = libgsb200.so through the same-signature ops; ms = median over CUDA-event timed launches"}
    for cfg in ("c2", "c3"):
        res[cfg] = run(cfg, ref)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
