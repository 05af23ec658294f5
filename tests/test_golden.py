"""Golden vectors: outputs of the UNMODIFIED reference `_gs` extension (tests/golden/*.npz, generated on a
B200 by tests/golden/make_golden.py).

CPU tests (no marker): the oracle (oracle/oracle.c + oracle/__init__.py) reproduces them -> the oracle is a
verified stand-in for the reference.  GPU tests (-m gpu): libgsb200.so reproduces them too."""
import pytest
import torch

from tests.util import assert_grad_close, assert_image_close, golden_names, load_golden

NAMES = golden_names()
needs_golden = pytest.mark.skipif(not NAMES, reason="tests/golden/*.npz not generated yet")


def _cfg(z):
    fx, fy, cx, cy, w, h = [float(x) for x in z["in_cam"]]
    H, W = int(h), int(w)
    return dict(H=H, W=W, n_tiles_h=(H + 15) // 16, n_tiles_w=(W + 15) // 16, psx=1.0 / fx, psy=1.0 / fy, thresh=1e-4,
                tile_size=16)


def _sh_tags(z):
    return sorted(k[len("ref_"):-len("_rgb")] for k in z if k.startswith("ref_sh") and k.endswith("_rgb"))


@needs_golden
@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_golden(oracle_mod, name):
    o = oracle_mod
    z = load_golden(name)
    cfg = _cfg(z)
    H, W = cfg["H"], cfg["W"]
    # K1
    mask = o.cull_bsphere(z["in_mean"], z["in_svec"], z["in_normals"], z["in_pts"], 6.0)
    assert torch.equal(mask, z["ref_mask"].bool())
    # K2-K4: bit-exact (ties allowed to differ only between identical keys)
    D = int(z["in_D"][0])
    ids, start, end = o.tile_culling_aabb_start_end(z["in_aabb_tl"], z["in_aabb_br"], z["in_depth"], cfg["n_tiles_h"],
                                                    cfg["n_tiles_w"], D)
    assert torch.equal(start, z["ref_start"]) and torch.equal(end, z["ref_end"])
    neq = ids != z["ref_ids"]
    if bool(neq.any()):
        dp = z["in_depth"].view(-1)
        assert torch.equal(dp[ids[neq].long()], dp[z["ref_ids"][neq].long()])
    ids = z["ref_ids"]
    m2, c2, col, al, tlf = z["in_mean2d"], z["in_cov2d"], z["in_color"], z["in_alpha"], z["in_topleft"]
    # K5
    out, T, stats, margin = o.composite_rgb_fwd(m2, c2, col, al, start, end, ids, tlf, cfg, want_margin=True)
    assert_image_close(out, z["ref_rgb"], margin, what="K5 rgb", atol=1e-5)
    assert_image_close(T, z["ref_T"].view(H, W), margin, what="K5 T", atol=1e-5)
    # K6
    final = z["ref_rgb"] + z["ref_T"] * z["in_bg"]
    gm, gc, gcol, ga = o.composite_rgb_bwd(m2, c2, col, al, start, end, ids, final, z["in_gout"], tlf, cfg)
    assert_grad_close(gm, z["ref_g_mean2d"], 2e-4, "K6 g_mean2d")
    assert_grad_close(gc, z["ref_g_cov2d"], 2e-4, "K6 g_cov2d")
    assert_grad_close(gcol, z["ref_g_color"], 2e-4, "K6 g_color")
    assert_grad_close(ga, z["ref_g_alpha"], 2e-4, "K6 g_alpha")
    # K7 / K8
    so, sT = o.composite_scalar_fwd(m2, c2, z["in_depth"].view(-1), al, start, end, ids, tlf, cfg)
    zs = max(1.0, float(z["ref_scalar"].abs().max()))
    assert_image_close(so / zs, z["ref_scalar"].view(H, W) / zs, margin, what="K7", atol=1e-5)
    gm, gc, gs, ga = o.composite_scalar_bwd(m2, c2, z["in_depth"].view(-1), al, start, end, ids,
                                            z["ref_scalar"].view(H, W), z["in_g_scalar_out"].view(H, W), tlf, cfg)
    assert_grad_close(gm, z["ref_s_g_mean2d"], 2e-4, "K8 g_mean2d")
    assert_grad_close(gc, z["ref_s_g_cov2d"], 2e-4, "K8 g_cov2d")
    assert_grad_close(gs, z["ref_s_g_scalar"].view(-1), 2e-4, "K8 g_scalar")
    assert_grad_close(ga, z["ref_s_g_alpha"], 2e-4, "K8 g_alpha")
    # K9-K11
    for tag in _sh_tags(z):
        C = int(tag[2])
        with_bg = tag.endswith("bg")
        sh = z[f"in_sh{C}"]
        out, T, stats, mg = o.composite_sh_fwd(m2, c2, sh, al, start, end, ids, tlf, z["in_c2w"], C, cfg,
                                               z["in_bg_rgb"] if with_bg else None, want_margin=True)
        assert_image_close(out, z[f"ref_{tag}_rgb"].view(H, W, 3), mg, what=f"{tag} rgb", atol=1e-5)
        gm, gc, gsh, ga = o.composite_sh_bwd(m2, c2, sh, al, start, end, ids, z[f"ref_{tag}_rgb"].view(H, W, 3),
                                             z["in_gout_sh"].view(H, W, 3), tlf, z["in_c2w"], C, cfg)
        assert_grad_close(gm, z[f"ref_{tag}_g_mean2d"], 3e-4, f"{tag} g_mean2d")
        assert_grad_close(gc, z[f"ref_{tag}_g_cov2d"], 3e-4, f"{tag} g_cov2d")
        assert_grad_close(gsh, z[f"ref_{tag}_g_sh"], 3e-4, f"{tag} g_sh")
        assert_grad_close(ga, z[f"ref_{tag}_g_alpha"], 3e-4, f"{tag} g_alpha")


@needs_golden
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_reference_golden(oracle_mod, name):
    from gsgen_b200.backend import _backend

    dev = "cuda"
    z = load_golden(name)
    cfg = _cfg(z)
    H, W = cfg["H"], cfg["W"]
    th, tw = cfg["n_tiles_h"], cfg["n_tiles_w"]
    d = lambda t: t.to(dev).contiguous()
    mask = torch.zeros(z["in_mean"].shape[0], dtype=torch.bool, device=dev)
    _backend.culling_gaussian_bsphere(d(z["in_mean"]), d(z["in_qvec"]), d(z["in_svec"]), d(z["in_normals"]),
                                      d(z["in_pts"]), mask, 6.0)
    assert torch.equal(mask.cpu(), z["ref_mask"].bool())
    D = int(z["in_D"][0])
    ids = torch.zeros(D, dtype=torch.int32, device=dev)
    start = -torch.ones(th * tw, dtype=torch.int32, device=dev)
    end = -torch.ones(th * tw, dtype=torch.int32, device=dev)
    _backend.tile_culling_aabb_start_end(d(z["in_aabb_tl"]), d(z["in_aabb_br"]), ids, start, end, d(z["in_depth"]),
                                         th, tw)
    assert torch.equal(start.cpu(), z["ref_start"]) and torch.equal(end.cpu(), z["ref_end"])
    neq = ids.cpu() != z["ref_ids"]
    if bool(neq.any()):
        dp = z["in_depth"].view(-1)
        assert torch.equal(dp[ids.cpu()[neq].long()], dp[z["ref_ids"][neq].long()])
    ids = d(z["ref_ids"])
    m2, c2, col, al, tlf = d(z["in_mean2d"]), d(z["in_cov2d"]), d(z["in_color"]), d(z["in_alpha"]), d(z["in_topleft"])
    common = (16, th, tw, cfg["psx"], cfg["psy"], H, W, 1e-4)
    _, _, _, margin = oracle_mod.composite_rgb_fwd(z["in_mean2d"], z["in_cov2d"], z["in_color"], z["in_alpha"],
                                                   z["ref_start"], z["ref_end"], z["ref_ids"], z["in_topleft"], cfg,
                                                   want_margin=True)
    out, T = torch.zeros(H, W, 3, device=dev), torch.ones(H, W, 1, device=dev)
    _backend.tile_based_vol_rendering_start_end_with_T(m2, c2, col, al, start, end, ids, out, tlf, *common, T)
    assert_image_close(out, z["ref_rgb"], margin, what="K5 rgb")
    assert_image_close(T, z["ref_T"], margin, what="K5 T")
    final = d(z["ref_rgb"] + z["ref_T"] * z["in_bg"])
    gm, gc = torch.zeros_like(m2), torch.zeros_like(c2)
    gcol, ga = torch.zeros_like(col), torch.zeros_like(al)
    _backend.tile_based_vol_rendering_backward_start_end(m2, c2, col, al, start, end, ids, final, gm, gc, gcol, ga,
                                                         d(z["in_gout"]), tlf, *common)
    assert_grad_close(gm, z["ref_g_mean2d"], what="K6 g_mean2d")
    assert_grad_close(gc, z["ref_g_cov2d"], what="K6 g_cov2d")
    assert_grad_close(gcol, z["ref_g_color"], what="K6 g_color")
    assert_grad_close(ga, z["ref_g_alpha"], what="K6 g_alpha")
    for tag in _sh_tags(z):
        C = int(tag[2])
        with_bg = tag.endswith("bg")
        sh = d(z[f"in_sh{C}"])
        o = torch.zeros(H * W * 3, device=dev)
        gm, gc = torch.zeros_like(m2), torch.zeros_like(c2)
        gsh, ga = torch.zeros_like(sh), torch.zeros_like(al)
        ref_rgb = d(z[f"ref_{tag}_rgb"])
        if with_bg:
            _backend.tile_based_vol_rendering_sh_with_bg(m2, c2, sh, al, start, end, ids, o, tlf, d(z["in_c2w"]),
                                                         *common[:7], C, 1e-4, d(z["in_bg_rgb"]))
            _backend.tile_based_vol_rendering_backward_sh_with_bg(m2, c2, sh, al, start, end, ids, ref_rgb, gm, gc,
                                                                  gsh, ga, d(z["in_gout_sh"]), tlf, d(z["in_c2w"]),
                                                                  *common[:7], C, 1e-4, d(z["in_bg_rgb"]))
        else:
            _backend.tile_based_vol_rendering_sh(m2, c2, sh, al, start, end, ids, o, tlf, d(z["in_c2w"]), *common[:7],
                                                 C, 1e-4)
            _backend.tile_based_vol_rendering_backward_sh(m2, c2, sh, al, start, end, ids, ref_rgb, gm, gc, gsh, ga,
                                                          d(z["in_gout_sh"]), tlf, d(z["in_c2w"]), *common[:7], C,
                                                          1e-4)
        assert_image_close(o.view(H, W, 3), z[f"ref_{tag}_rgb"].view(H, W, 3), margin, what=f"{tag} rgb")
        assert_grad_close(gm, z[f"ref_{tag}_g_mean2d"], what=f"{tag} g_mean2d")
        assert_grad_close(gc, z[f"ref_{tag}_g_cov2d"], what=f"{tag} g_cov2d")
        assert_grad_close(gsh, z[f"ref_{tag}_g_sh"], what=f"{tag} g_sh")
        assert_grad_close(ga, z[f"ref_{tag}_g_alpha"], what=f"{tag} g_alpha")
