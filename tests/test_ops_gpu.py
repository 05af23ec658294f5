"""-m gpu parity tests: every reference-compatible op of libgsb200.so (called through the `_backend`
shim -> C ABI) against the CPU oracle on the same seeded inputs.

Integer / index work (cull mask, AABBs, duplicate count, sorted ids, start/end) must be bit-exact.
Images: 1e-4 abs (threshold-flip pixels accounted for, tests/util.py); gradients: 1e-3 rel."""
import pytest
import torch

from gsgen_b200.scenes import make_scene, mock_two_gaussians
from tests.util import assert_grad_close, assert_image_close, ocam_of

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _prep(oracle, sc, cam, c2w):
    """CPU oracle: cull -> gather -> project -> aabb -> bin."""
    ocam = ocam_of(cam)
    normals, pts = oracle.get_frustum(ocam, c2w)
    mask = oracle.cull_bsphere(sc.mean, sc.svec, normals, pts, 6.0)
    m, q, s = sc.mean[mask].contiguous(), sc.qvec[mask].contiguous(), sc.svec[mask].contiguous()
    mean2d, cov2d, _, depth = oracle.project_gaussians(m, q, s, c2w, True)
    D, tl, br = oracle.tile_culling_aabb_count(mean2d, cov2d, 16, ocam, 6.0)
    cfg = oracle.view_cfg(ocam)
    ids, start, end = oracle.tile_culling_aabb_start_end(tl, br, depth, cfg["n_tiles_h"], cfg["n_tiles_w"], D)
    topleft = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=torch.float32)
    return dict(ocam=ocam, normals=normals, pts=pts, mask=mask, mean2d=mean2d.contiguous(), cov2d=cov2d.contiguous(),
                depth=depth.contiguous(), D=D, tl=tl, br=br, cfg=cfg, ids=ids, start=start, end=end, topleft=topleft,
                alpha=sc.alpha[mask].contiguous(), color=sc.color[mask].contiguous(),
                sh=None if sc.sh is None else sc.sh[mask].contiguous())


SCENES = {
    "c1_small": lambda: make_scene("c1", N=3000, reso=128),
    "c1": lambda: make_scene("c1"),                       # BASELINE config 1: 10k / 256^2
    "dense": lambda: _dense(),
    "mock2": lambda: mock_two_gaussians(),
}


def _dense():
    sc = make_scene("c3", N=6000, reso=160)
    sc.svec = (sc.svec * 4.0).contiguous()
    return sc


@pytest.fixture(scope="module", params=list(SCENES))
def view(request, oracle_mod):
    sc = SCENES[request.param]()
    cam, c2w = sc.cams[0], sc.c2ws[0]
    if request.param == "mock2":  # quarter-size camera keeps the oracle fast
        from gsgen_b200.camera import CameraInfo

        cam = CameraInfo(cam.fx / 4, cam.fy / 4, cam.cx / 4, cam.cy / 4, cam.w // 4, cam.h // 4, 0.01, 100.0)
    p = _prep(oracle_mod, sc, cam, c2w)
    p.update(sc=sc, cam=cam, c2w=c2w, name=request.param)
    return p


def _d(t):
    return t.to(DEV).contiguous()


def test_cull_bit_exact(view):
    from gsgen_b200.backend import _backend

    sc = view["sc"]
    mask = torch.zeros(sc.N, dtype=torch.bool, device=DEV)
    _backend.culling_gaussian_bsphere(_d(sc.mean), _d(sc.qvec), _d(sc.svec), _d(view["normals"]), _d(view["pts"]),
                                      mask, 6.0)
    assert torch.equal(mask.cpu(), view["mask"])


def test_project_gaussians_forward_backward(view, oracle_mod):
    from gsgen_b200.renderer import project_gaussians

    sc, c2w, mask = view["sc"], view["c2w"], view["mask"]
    g = torch.Generator().manual_seed(3)
    m = sc.mean[mask].clone()
    q = (sc.qvec[mask] * (0.5 + torch.rand(int(mask.sum()), 1, generator=g))).clone()
    s = sc.svec[mask].clone()
    N = m.shape[0]
    gm2, gcov, gdp = torch.randn(N, 2, generator=g), torch.randn(N, 2, 2, generator=g), torch.randn(N, 1, generator=g)
    for detach in (True, False):
        mo, qo, so = m.clone().requires_grad_(), q.clone().requires_grad_(), s.clone().requires_grad_()
        m2o, covo, JWo, dpo = oracle_mod.project_gaussians(mo, qo, so, c2w, detach)
        ((m2o * gm2).sum() + (covo * gcov).sum() + (dpo * gdp).sum()).backward()
        mg, qg, sg = _d(m).requires_grad_(), _d(q).requires_grad_(), _d(s).requires_grad_()
        m2, cov, JW, dp = project_gaussians(mg, qg, sg, _d(c2w), detach)
        ((m2 * _d(gm2)).sum() + (cov * _d(gcov)).sum() + (dp * _d(gdp)).sum()).backward()
        assert torch.allclose(m2.cpu(), m2o, rtol=2e-5, atol=1e-6)
        assert torch.allclose(dp.cpu(), dpo, rtol=2e-6, atol=1e-6)
        rel = (cov.cpu() - covo).abs().amax(dim=(1, 2)) / covo.abs().amax(dim=(1, 2))
        assert float(rel.max()) < 5e-5
        assert torch.allclose(JW.cpu(), JWo, rtol=1e-4, atol=1e-5)
        assert_grad_close(mg.grad, mo.grad, 1e-4, "g_mean")
        assert_grad_close(qg.grad, qo.grad, 1e-4, "g_qvec")
        assert_grad_close(sg.grad, so.grad, 1e-4, "g_svec")


def test_aabb_count_bit_exact(view):
    from gsgen_b200.culling import tile_culling_aabb_count

    D, tl, br = tile_culling_aabb_count(_d(view["mean2d"]), _d(view["cov2d"]), 16, view["cam"], 6.0)
    assert D == view["D"]
    assert torch.equal(tl.cpu(), view["tl"]) and torch.equal(br.cpu(), view["br"])


def test_binning_bit_exact(view):
    from gsgen_b200.backend import _backend

    cfg = view["cfg"]
    T = cfg["n_tiles_h"] * cfg["n_tiles_w"]
    ids = torch.zeros(view["D"], dtype=torch.int32, device=DEV)
    start = -torch.ones(T, dtype=torch.int32, device=DEV)
    end = -torch.ones(T, dtype=torch.int32, device=DEV)
    _backend.tile_culling_aabb_start_end(_d(view["tl"]), _d(view["br"]), ids, start, end, _d(view["depth"]),
                                         cfg["n_tiles_h"], cfg["n_tiles_w"])
    assert torch.equal(start.cpu(), view["start"]) and torch.equal(end.cpu(), view["end"])
    assert torch.equal(ids.cpu(), view["ids"])  # ties broken by Gaussian index on both sides
    # a wrong gaussian_ids size is reported like the reference's host assert (aabb_culling.h:228)
    if view["D"] > 1:
        with pytest.raises(RuntimeError):
            _backend.tile_culling_aabb_start_end(_d(view["tl"]), _d(view["br"]), ids[:-1].contiguous(), start, end,
                                                 _d(view["depth"]), cfg["n_tiles_h"], cfg["n_tiles_w"])


def _common_args(view):
    cfg = view["cfg"]
    return (16, cfg["n_tiles_h"], cfg["n_tiles_w"], cfg["psx"], cfg["psy"], cfg["H"], cfg["W"], cfg["thresh"])


def test_rgb_composite_forward_backward(view, oracle_mod):
    from gsgen_b200.renderer import render_with_T

    v = view
    H, W = v["cfg"]["H"], v["cfg"]["W"]
    g = torch.Generator().manual_seed(17)
    bg = torch.rand(H, W, 3, generator=g)
    gout = torch.randn(H, W, 3, generator=g)
    # oracle
    out_o, T_o, stats, margin = oracle_mod.composite_rgb_fwd(v["mean2d"], v["cov2d"], v["color"], v["alpha"],
                                                            v["start"], v["end"], v["ids"], v["topleft"], v["cfg"],
                                                            want_margin=True)
    final_o = out_o + T_o.unsqueeze(-1) * bg
    gm_o, gc_o, gcol_o, ga_o = oracle_mod.composite_rgb_bwd(v["mean2d"], v["cov2d"], v["color"], v["alpha"],
                                                          v["start"], v["end"], v["ids"], final_o, gout, v["topleft"],
                                                          v["cfg"])
    # ours, through the reference-shaped autograd Function
    m2, c2 = _d(v["mean2d"]).requires_grad_(), _d(v["cov2d"]).requires_grad_()
    col, al = _d(v["color"]).requires_grad_(), _d(v["alpha"]).requires_grad_()
    bgd = _d(bg).requires_grad_()
    out = render_with_T(m2, c2, col, al, _d(v["start"]), _d(v["end"]), _d(v["ids"]), _d(v["topleft"]),
                        *_common_args(v), bgd)
    assert_image_close(out, final_o, margin, what=f"{v['name']} rgb")
    (out * _d(gout)).sum().backward()
    assert_grad_close(m2.grad, gm_o, what="g_mean2d")
    assert_grad_close(c2.grad, gc_o, what="g_cov2d")
    assert_grad_close(col.grad, gcol_o, what="g_color")
    assert_grad_close(al.grad, ga_o, what="g_alpha")
    assert_image_close(bgd.grad, torch.nan_to_num(gout * T_o.unsqueeze(-1)), margin, what="g_bg", atol=2e-4)


def test_T_output_and_start_end_variant(view, oracle_mod):
    from gsgen_b200.backend import _backend
    from gsgen_b200.renderer import render_start_end

    v = view
    H, W = v["cfg"]["H"], v["cfg"]["W"]
    out_o, T_o, stats, margin = oracle_mod.composite_rgb_fwd(v["mean2d"], v["cov2d"], v["color"], v["alpha"],
                                                            v["start"], v["end"], v["ids"], v["topleft"], v["cfg"],
                                                            want_margin=True)
    out = torch.zeros(H, W, 3, device=DEV)
    T = torch.ones(H, W, 1, device=DEV)
    _backend.tile_based_vol_rendering_start_end_with_T(
        _d(v["mean2d"]), _d(v["cov2d"]), _d(v["color"]), _d(v["alpha"]), _d(v["start"]), _d(v["end"]), _d(v["ids"]),
        out, _d(v["topleft"]), *_common_args(v), T)
    assert_image_close(out, out_o, margin, what="rgb")
    assert_image_close(T.view(H, W), T_o, margin, what="T")
    out2 = render_start_end(_d(v["mean2d"]), _d(v["cov2d"]), _d(v["color"]), _d(v["alpha"]), _d(v["start"]),
                            _d(v["end"]), _d(v["ids"]), _d(v["topleft"]), *_common_args(v))
    assert torch.equal(out2.view(H, W, 3), out)  # same kernel, deterministic forward


def test_scalar_composite_forward_backward(view, oracle_mod):
    from gsgen_b200.renderer import render_scalar

    v = view
    H, W = v["cfg"]["H"], v["cfg"]["W"]
    g = torch.Generator().manual_seed(23)
    gout = torch.randn(H * W, generator=g)
    scalar = v["depth"].reshape(-1).contiguous()
    out_o, T_o = oracle_mod.composite_scalar_fwd(v["mean2d"], v["cov2d"], scalar, v["alpha"], v["start"], v["end"],
                                                 v["ids"], v["topleft"], v["cfg"])
    gm_o, gc_o, gs_o, ga_o = oracle_mod.composite_scalar_bwd(v["mean2d"], v["cov2d"], scalar, v["alpha"], v["start"],
                                                          v["end"], v["ids"], out_o, gout.view(H, W), v["topleft"],
                                                          v["cfg"])
    _, _, _, margin = oracle_mod.composite_rgb_fwd(v["mean2d"], v["cov2d"], v["color"], v["alpha"], v["start"],
                                                   v["end"], v["ids"], v["topleft"], v["cfg"], want_margin=True)
    m2, c2 = _d(v["mean2d"]).requires_grad_(), _d(v["cov2d"]).requires_grad_()
    sc_, al = _d(v["depth"]).requires_grad_(), _d(v["alpha"]).requires_grad_()
    T = torch.ones(H, W, 1, device=DEV)
    out = render_scalar(m2, c2, sc_, al, _d(v["start"]), _d(v["end"]), _d(v["ids"]), _d(v["topleft"]),
                        *_common_args(v), T)
    scale = max(1.0, float(out_o.abs().max()))
    assert_image_close(out.view(H, W) / scale, out_o / scale, margin, what="scalar")
    assert_image_close(T.view(H, W), T_o, margin, what="scalar T")
    (out * _d(gout)).sum().backward()
    assert_grad_close(m2.grad, gm_o, what="g_mean2d")
    assert_grad_close(c2.grad, gc_o, what="g_cov2d")
    assert_grad_close(sc_.grad.view(-1), gs_o, what="g_scalar")
    assert_grad_close(al.grad, ga_o, what="g_alpha")


@pytest.mark.parametrize("C", [1, 2, 3, 4])
@pytest.mark.parametrize("with_bg", [False, True])
def test_sh_composite_forward_backward(view, oracle_mod, C, with_bg):
    from gsgen_b200.renderer import render_sh, render_sh_bg

    v = view
    if v["name"] == "c1" and (C in (2, 3) or with_bg):
        pytest.skip("full C1 size is exercised for C=1,4 without background; smaller scenes cover the rest")
    H, W = v["cfg"]["H"], v["cfg"]["W"]
    N = v["mean2d"].shape[0]
    g = torch.Generator().manual_seed(100 + C)
    sh = (0.5 * torch.randn(N, 3, C * C, generator=g)).contiguous()
    gout = torch.randn(H * W * 3, generator=g)
    bg_rgb = torch.tensor([0.2, 0.5, 0.7]) if with_bg else None
    c2w = v["c2w"].contiguous()  # [3,4]: the nine-float read is reproduced literally (A.7)
    out_o, T_o, stats, margin = oracle_mod.composite_sh_fwd(v["mean2d"], v["cov2d"], sh, v["alpha"], v["start"],
                                                           v["end"], v["ids"], v["topleft"], c2w, C, v["cfg"], bg_rgb,
                                                           want_margin=True)
    gm_o, gc_o, gsh_o, ga_o = oracle_mod.composite_sh_bwd(v["mean2d"], v["cov2d"], sh, v["alpha"], v["start"],
                                                         v["end"], v["ids"], out_o, gout.view(H, W, 3), v["topleft"],
                                                         c2w, C, v["cfg"])
    m2, c2 = _d(v["mean2d"]).requires_grad_(), _d(v["cov2d"]).requires_grad_()
    shd, al = _d(sh).requires_grad_(), _d(v["alpha"]).requires_grad_()
    args = (m2, c2, shd, al, _d(v["start"]), _d(v["end"]), _d(v["ids"]), _d(v["topleft"]), _d(c2w), 16,
            v["cfg"]["n_tiles_h"], v["cfg"]["n_tiles_w"], v["cfg"]["psx"], v["cfg"]["psy"], H, W, C, v["cfg"]["thresh"])
    out = render_sh_bg(*args, _d(bg_rgb)) if with_bg else render_sh(*args)
    assert_image_close(out.view(H, W, 3), out_o, margin, what=f"sh C={C} bg={with_bg}")
    (out * _d(gout)).sum().backward()
    assert_grad_close(m2.grad, gm_o, what="g_mean2d")
    assert_grad_close(c2.grad, gc_o, what="g_cov2d")
    assert_grad_close(shd.grad, gsh_o, what="g_sh")
    assert_grad_close(al.grad, ga_o, what="g_alpha")


def test_contract_errors():
    from gsgen_b200.backend import _backend

    z = torch.zeros(4, 2, device=DEV)
    with pytest.raises(RuntimeError):  # tile_size != 16 is unsupported
        _backend.tile_based_vol_rendering_start_end_with_T(
            z, torch.zeros(4, 2, 2, device=DEV), torch.zeros(4, 3, device=DEV), torch.zeros(4, device=DEV),
            torch.zeros(4, dtype=torch.int32, device=DEV), torch.zeros(4, dtype=torch.int32, device=DEV),
            torch.zeros(1, dtype=torch.int32, device=DEV), torch.zeros(32, 32, 3, device=DEV),
            torch.zeros(2, device=DEV), 8, 4, 4, 0.01, 0.01, 32, 32, 1e-4, torch.ones(32, 32, 1, device=DEV))
    with pytest.raises(RuntimeError):  # non-contiguous input (CHECK_CONTIGUOUS)
        _backend.culling_gaussian_bsphere(torch.zeros(3, 4, device=DEV).t(), torch.zeros(4, 4, device=DEV),
                                          torch.zeros(4, 3, device=DEV), torch.zeros(6, 3, device=DEV),
                                          torch.zeros(6, 3, device=DEV), torch.zeros(4, dtype=torch.bool, device=DEV),
                                          6.0)


def test_empty_inputs(oracle_mod):
    """Edge cases: empty tile lists everywhere (outputs keep the caller's initial values, A.9-15) and N = 0."""
    from gsgen_b200.backend import _backend

    H = W = 40
    th = tw = 3
    start = -torch.ones(th * tw, dtype=torch.int32, device=DEV)
    end = -torch.ones(th * tw, dtype=torch.int32, device=DEV)
    ids = torch.zeros(0, dtype=torch.int32, device=DEV)
    out = torch.full((H, W, 3), 0.25, device=DEV)
    T = torch.full((H, W, 1), 0.5, device=DEV)
    z = lambda *s: torch.zeros(*s, device=DEV)
    _backend.tile_based_vol_rendering_start_end_with_T(z(0, 2), z(0, 2, 2), z(0, 3), z(0), start, end, ids, out,
                                                       z(2), 16, th, tw, 0.01, 0.01, H, W, 1e-4, T)
    assert float(out.min()) == 0.25 and float(T.max()) == 0.5
    bg_rgb = torch.tensor([0.1, 0.2, 0.3], device=DEV)
    o = torch.zeros(H * W * 3, device=DEV)
    _backend.tile_based_vol_rendering_sh_with_bg(z(0, 2), z(0, 2, 2), z(0, 3, 4), z(0), start, end, ids, o, z(2),
                                                 torch.eye(3, 4, device=DEV), 16, th, tw, 0.01, 0.01, H, W, 2, 1e-4,
                                                 bg_rgb)
    assert torch.allclose(o.view(H, W, 3), bg_rgb.expand(H, W, 3))  # vol_render_bg.h:34-53
    _backend.tile_culling_aabb_start_end(torch.zeros(0, 2, dtype=torch.int32, device=DEV),
                                         torch.zeros(0, 2, dtype=torch.int32, device=DEV), ids, start, end, z(0, 1),
                                         th, tw)
    assert int(start.max()) == -1
