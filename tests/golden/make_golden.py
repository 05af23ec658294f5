"""Generate golden vectors from the UNMODIFIED reference `_gs` CUDA extension (TEST INFRASTRUCTURE).

Run on a B200 box (`gpurun -- python tests/golden/make_golden.py`) after `oracle/build_ref.sh` built
`oracle/_ref/_gs.so` from /root/reference in the dev container.  Writes `gpurun_out/golden/*.npz`; the files
are then copied to `tests/golden/` and committed.  They pin the CPU oracle (`-m "not gpu"` tests) and are a
second, box-independent target for the CUDA path (`-m gpu` tests).

What is pinned: every `_gs` op on the hot path (cull K1, bin/sort K2-K4, RGB composite fwd/bwd K5/K6, scalar
K7/K8, SH K9/K10, SH+bg K11) on seeded inputs.  The per-Gaussian torch stage of the reference (projection,
AABB count) cannot be imported here (kornia/torchtyping/... are absent), so (mean2d, cov2d, depth, aabb) come
from the oracle's torch restatement and are stored as INPUTS.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

import oracle  # noqa: E402
from gsgen_b200.scenes import make_scene, mock_two_gaussians  # noqa: E402
from gsgen_b200.camera import CameraInfo  # noqa: E402


def prep(sc, cam, c2w):
    """CPU: cull -> gather -> project -> aabb (oracle restatement)."""
    ocam = oracle.Cam(cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h, cam.near_plane, cam.far_plane)
    normals, pts = oracle.get_frustum(ocam, c2w)
    cfg = oracle.view_cfg(ocam)
    return ocam, normals, pts, cfg


def run_case(name, sc, cam, c2w, seed, out_dir, sh_variants=((1, False), (4, False), (3, True))):
    import _gs  # the reference extension

    dev = torch.device("cuda")
    ocam, normals, pts, cfg = prep(sc, cam, c2w)
    g = torch.Generator().manual_seed(seed)
    H, W = cam.h, cam.w
    th, tw = cfg["n_tiles_h"], cfg["n_tiles_w"]
    res = {}
    # ---- K1 cull on the full set
    mask = torch.zeros(sc.N, dtype=torch.bool, device=dev)
    _gs.culling_gaussian_bsphere(sc.mean.to(dev), sc.qvec.to(dev), sc.svec.to(dev), normals.to(dev), pts.to(dev),
                                 mask, 6.0)
    torch.cuda.synchronize()
    mask_c = mask.cpu()
    res.update(in_mean=sc.mean, in_qvec=sc.qvec, in_svec=sc.svec, in_normals=normals, in_pts=pts, ref_mask=mask_c)
    # ---- per-Gaussian torch stage (oracle restatement), stored as inputs
    m, q, s = sc.mean[mask_c].contiguous(), sc.qvec[mask_c].contiguous(), sc.svec[mask_c].contiguous()
    alpha = sc.alpha[mask_c].contiguous()
    color = sc.color[mask_c].contiguous()
    mean2d, cov2d, _, depth = oracle.project_gaussians(m, q, s, c2w, True)
    D, tl, br = oracle.tile_culling_aabb_count(mean2d, cov2d, 16, ocam, 6.0)
    topleft = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=torch.float32)
    res.update(in_mean2d=mean2d, in_cov2d=cov2d, in_depth=depth, in_alpha=alpha, in_color=color, in_aabb_tl=tl,
               in_aabb_br=br, in_topleft=topleft, in_c2w=c2w,
               in_cam=torch.tensor([cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h], dtype=torch.float64),
               in_D=torch.tensor([D]))
    # ---- K2-K4
    ids = torch.zeros(D, dtype=torch.int32, device=dev)
    start = -torch.ones(th * tw, dtype=torch.int32, device=dev)
    end = -torch.ones(th * tw, dtype=torch.int32, device=dev)
    _gs.tile_culling_aabb_start_end(tl.to(dev), br.to(dev), ids, start, end, depth.to(dev), th, tw)
    torch.cuda.synchronize()
    res.update(ref_ids=ids.cpu(), ref_start=start.cpu(), ref_end=end.cpu())
    d = lambda t: t.to(dev).contiguous()
    m2, c2, al, col, dp, tlf = d(mean2d), d(cov2d), d(alpha), d(color), d(depth), d(topleft)
    psx, psy = 1.0 / cam.fx, 1.0 / cam.fy
    # ---- K5 / K6 RGB
    out = torch.zeros(H, W, 3, device=dev)
    T = torch.ones(H, W, 1, device=dev)
    _gs.tile_based_vol_rendering_start_end_with_T(m2, c2, col, al, start, end, ids, out, tlf, 16, th, tw, psx, psy, H,
                                                  W, 1e-4, T)
    torch.cuda.synchronize()
    bg = torch.rand(H, W, 3, generator=g)
    final = out + T * bg.to(dev)
    gout = torch.randn(H, W, 3, generator=g)
    gm, gc = torch.zeros_like(m2), torch.zeros_like(c2)
    gcol, ga = torch.zeros_like(col), torch.zeros_like(al)
    _gs.tile_based_vol_rendering_backward_start_end(m2, c2, col, al, start, end, ids, final.contiguous(), gm, gc, gcol,
                                                    ga, d(gout), tlf, 16, th, tw, psx, psy, H, W, 1e-4)
    torch.cuda.synchronize()
    res.update(ref_rgb=out.cpu(), ref_T=T.cpu(), in_bg=bg, in_gout=gout, ref_g_mean2d=gm.cpu(), ref_g_cov2d=gc.cpu(),
               ref_g_color=gcol.cpu(), ref_g_alpha=ga.cpu())
    # ---- K7 / K8 scalar (depth payload, as render_one does)
    so = torch.zeros(H * W, device=dev)
    sT = torch.ones(H, W, 1, device=dev)
    _gs.tile_based_vol_rendering_scalar(m2, c2, dp, al, start, end, ids, so, tlf, 16, th, tw, psx, psy, H, W, 1e-4, sT)
    torch.cuda.synchronize()
    gso = torch.randn(H * W, generator=g)
    gm, gc = torch.zeros_like(m2), torch.zeros_like(c2)
    gs_, ga = torch.zeros_like(dp), torch.zeros_like(al)
    _gs.tile_based_vol_rendering_scalar_backward(m2, c2, dp, al, start, end, ids, so, gm, gc, gs_, ga, d(gso), tlf, 16,
                                                 th, tw, psx, psy, H, W, 1e-4)
    torch.cuda.synchronize()
    res.update(ref_scalar=so.cpu(), ref_scalar_T=sT.cpu(), in_g_scalar_out=gso, ref_s_g_mean2d=gm.cpu(),
               ref_s_g_cov2d=gc.cpu(), ref_s_g_scalar=gs_.cpu(), ref_s_g_alpha=ga.cpu())
    # ---- K9 / K10 / K11 SH
    bg_rgb = torch.tensor([0.2, 0.5, 0.7])
    gout_sh = torch.randn(H * W * 3, generator=g)
    res.update(in_bg_rgb=bg_rgb, in_gout_sh=gout_sh)
    for C, with_bg in sh_variants:
        gsh = torch.Generator().manual_seed(seed + 10 + C)
        sh = (0.5 * torch.randn(m2.shape[0], 3, C * C, generator=gsh)).contiguous()
        res[f"in_sh{C}"] = sh
        shd = d(sh)
        if True:
            o = torch.zeros(H * W * 3, device=dev)
            gm, gc = torch.zeros_like(m2), torch.zeros_like(c2)
            gshc, ga = torch.zeros_like(shd), torch.zeros_like(al)
            if with_bg:
                _gs.tile_based_vol_rendering_sh_with_bg(m2, c2, shd, al, start, end, ids, o, tlf, d(c2w), 16, th, tw,
                                                        psx, psy, H, W, C, 1e-4, d(bg_rgb))
                torch.cuda.synchronize()
                _gs.tile_based_vol_rendering_backward_sh_with_bg(m2, c2, shd, al, start, end, ids, o, gm, gc, gshc, ga,
                                                                 d(gout_sh), tlf, d(c2w), 16, th, tw, psx, psy, H, W,
                                                                 C, 1e-4, d(bg_rgb))
            else:
                _gs.tile_based_vol_rendering_sh(m2, c2, shd, al, start, end, ids, o, tlf, d(c2w), 16, th, tw, psx,
                                                psy, H, W, C, 1e-4)
                torch.cuda.synchronize()
                _gs.tile_based_vol_rendering_backward_sh(m2, c2, shd, al, start, end, ids, o, gm, gc, gshc, ga,
                                                         d(gout_sh), tlf, d(c2w), 16, th, tw, psx, psy, H, W, C, 1e-4)
            torch.cuda.synchronize()
            tag = f"sh{C}{'bg' if with_bg else ''}"
            res[f"ref_{tag}_rgb"] = o.cpu()
            res[f"ref_{tag}_g_mean2d"] = gm.cpu()
            res[f"ref_{tag}_g_cov2d"] = gc.cpu()
            res[f"ref_{tag}_g_sh"] = gshc.cpu()
            res[f"ref_{tag}_g_alpha"] = ga.cpu()
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **{k: v.numpy() for k, v in res.items()})
    print(f"golden {name}: N0={sc.N} N={m2.shape[0]} D={D} {H}x{W}", flush=True)


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    oracle.build()
    # G1: random Gaussians (C1 generator), 128x128 -- 8x8 tiles, lists ~100 long
    sc = make_scene("c1", N=3000, reso=128)
    run_case("g1_c1_3000_128", sc, sc.cams[0], sc.c2ws[0], 101, out_dir)
    # G2: dense small scene with early termination (opaque, larger splats) and a ragged image size (not /16)
    sc = make_scene("c3", N=4000, reso=128)
    sc.svec = (sc.svec * 4.0).contiguous()
    cam = CameraInfo(1.1 * 120, 1.1 * 120, 60.0, 50.0, 120, 100, 0.01, 100.0)
    run_case("g2_dense_4000_120x100", sc, cam, sc.c2ws[0], 202, out_dir, sh_variants=((2, True), (4, False)))
    # G3: the reference's own 2-Gaussian MockRenderer scene (gs/debug.py:52-65), camera scaled by 1/8
    sc = mock_two_gaussians()
    c0 = sc.cams[0]
    cam = CameraInfo(c0.fx / 8, c0.fy / 8, c0.cx / 8, c0.cy / 8, c0.w // 8, c0.h // 8, 0.01, 100.0)
    run_case("g3_mock2_162x105", sc, cam, sc.c2ws[0], 303, out_dir, sh_variants=((2, False), (2, True)))


if __name__ == "__main__":
    main()
