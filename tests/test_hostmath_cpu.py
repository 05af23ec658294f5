"""CPU checks of the kernels' per-Gaussian math (gsgen_b200/csrc/gsb200_math.cuh compiled with g++)
against the oracle: projection forward/backward vs torch autograd of the restated project_gaussians,
integer AABBs bit-exact vs the torch op sequence, frustum cull bit-exact, splat factorisation vs the fp64
per-pixel formula, SH basis."""
import ctypes

import numpy as np
import pytest
import torch

from gsgen_b200.scenes import make_scene
from tests.util import fp, ip, ocam_of


def _scene(name="c1", N=4000, reso=256):
    sc = make_scene(name, N=N, reso=reso)
    return sc, sc.cams[0], sc.c2ws[0]


def test_projection_forward_matches_oracle(hostmath, oracle_mod):
    sc, cam, c2w = _scene()
    m2o, covo, _, dpo = oracle_mod.project_gaussians(sc.mean, sc.qvec * 1.7, sc.svec, c2w, True)
    N = sc.N
    m2, cov, dp = torch.empty(N, 2), torch.empty(N, 4), torch.empty(N)
    q = (sc.qvec * 1.7).contiguous()  # non-unit quaternions exercise the normalisation
    hostmath.hm_project_fwd(N, fp(sc.mean), fp(q), fp(sc.svec), fp(c2w.contiguous()), fp(m2), fp(cov), fp(dp))
    assert torch.allclose(m2, m2o, rtol=2e-5, atol=1e-6)
    assert torch.allclose(dp, dpo.view(-1), rtol=2e-6, atol=1e-6)
    rel = (cov - covo.reshape(N, 4)).abs().max(dim=1).values / covo.reshape(N, 4).abs().max(dim=1).values
    assert float(rel.max()) < 5e-5, float(rel.max())
    assert torch.equal(cov[:, 1], cov[:, 2])  # ours is symmetric by construction


@pytest.mark.parametrize("detach", [True, False])
def test_projection_backward_matches_autograd(hostmath, oracle_mod, detach):
    sc, cam, c2w = _scene(N=3000)
    g = torch.Generator().manual_seed(5)
    mean = sc.mean.clone().requires_grad_()
    q = (sc.qvec * (0.5 + torch.rand(sc.N, 1, generator=g))).clone().requires_grad_()
    s = sc.svec.clone().requires_grad_()
    m2, cov, _, dp = oracle_mod.project_gaussians(mean, q, s, c2w, detach)
    gm2, gcov, gdp = torch.randn(sc.N, 2, generator=g), torch.randn(sc.N, 2, 2, generator=g), torch.randn(sc.N, 1, generator=g)
    (m2 * gm2).sum().add((cov * gcov).sum()).add((dp * gdp).sum()).backward()
    gx, gq, gs = torch.empty(sc.N, 3), torch.empty(sc.N, 4), torch.empty(sc.N, 3)
    hostmath.hm_project_bwd(sc.N, fp(sc.mean), fp(q.detach().contiguous()), fp(sc.svec), fp(c2w.contiguous()),
                            1 if detach else 0, fp(gm2), fp(gcov.reshape(sc.N, 4).contiguous()),
                            fp(gdp.reshape(-1).contiguous()), fp(gx), fp(gq), fp(gs))
    for ours, ref, nm in ((gx, mean.grad, "mean"), (gq, q.grad, "qvec"), (gs, s.grad, "svec")):
        rel = (ours - ref).norm() / ref.norm()
        assert float(rel) < 1e-4, (nm, float(rel))
        # row-wise: no Gaussian is badly off
        rr = (ours - ref).norm(dim=1) / (ref.norm(dim=1) + 1e-3 * ref.norm(dim=1).mean())
        assert float(rr.max()) < 5e-3, (nm, float(rr.max()))


@pytest.mark.parametrize("scene,reso", [("c1", 256), ("c3", 200), ("c2", 144)])
def test_aabb_bit_exact(hostmath, oracle_mod, scene, reso):
    sc, cam, c2w = _scene(scene, 5000, reso)
    m2, cov, _, dp = oracle_mod.project_gaussians(sc.mean, sc.qvec, sc.svec, c2w, True)
    front = dp.view(-1) > 0.05
    m2, cov = m2[front].contiguous(), cov[front].contiguous()
    D, tl, br = oracle_mod.tile_culling_aabb_count(m2, cov, 16, ocam_of(cam), 6.0)
    N = m2.shape[0]
    tlo, bro = torch.empty(N, 2, dtype=torch.int32), torch.empty(N, 2, dtype=torch.int32)
    hostmath.hm_aabb(N, fp(m2), fp(cov.reshape(N, 4).contiguous()), ctypes.c_float(6.0), ctypes.c_float(cam.fx),
                     ctypes.c_float(cam.fy), ctypes.c_float(cam.cx), ctypes.c_float(cam.cy), cam.w, cam.h, 16,
                     ip(tlo), ip(bro))
    assert torch.equal(tlo, tl) and torch.equal(bro, br)
    assert int(((bro - tlo + 1).prod(dim=1)).sum()) == D


def test_cull_bit_exact(hostmath, oracle_mod):
    sc = make_scene("c3", N=20000, reso=128)
    cam, c2w = sc.cams[0], sc.c2ws[0]
    # move the camera inside the ball so that a good share of the Gaussians is outside the frustum
    from gsgen_b200.camera import orbit_c2w

    c2w = orbit_c2w(0.6, 10.0, 70.0)
    normals, pts = oracle_mod.get_frustum(ocam_of(cam), c2w)
    ref = oracle_mod.cull_bsphere(sc.mean, sc.svec, normals, pts, 6.0)
    mask = torch.zeros(sc.N, dtype=torch.uint8)
    n = hostmath.hm_cull(sc.N, fp(sc.mean), fp(sc.svec), fp(normals), fp(pts), ctypes.c_float(6.0),
                         ctypes.cast(mask.data_ptr(), ctypes.POINTER(ctypes.c_ubyte)))
    assert torch.equal(mask.bool(), ref)
    assert 0.2 * sc.N < n < 0.98 * sc.N  # the test actually culls something
    # product-side get_frustum mirrors the oracle's
    n2, p2 = cam.get_frustum(c2w)
    assert torch.allclose(n2, normals) and torch.allclose(p2, pts)


def test_splat_factorisation_matches_fp64_formula(hostmath, oracle_mod):
    """a*G from the Cholesky-factored record == the reference's per-pixel fp64 formula (kernels.h:195-224)."""
    sc, cam, c2w = _scene("c3", 6000, 256)
    m2, cov, _, dp = oracle_mod.project_gaussians(sc.mean, sc.qvec, sc.svec, c2w, True)
    N = m2.shape[0]
    g = torch.Generator().manual_seed(11)
    cov4 = cov.reshape(N, 4).contiguous()
    sig = torch.sqrt(torch.stack([cov4[:, 0], cov4[:, 3]], dim=1))
    query = (m2 + sig * (4.0 * torch.rand(N, 2, generator=g) - 2.0)).contiguous()
    alpha = sc.alpha.contiguous()
    aG, G, vx, vy, hx, hy = (torch.empty(N) for _ in range(6))
    hostmath.hm_splat_eval(N, fp(m2.contiguous()), fp(cov4), fp(alpha), fp(query), fp(aG), fp(G), fp(vx), fp(vy),
                           fp(hx), fp(hy))
    c = cov4.double()
    d = (query - m2).double()
    det = c[:, 0] * c[:, 3] - c[:, 1] * c[:, 2]
    tx = (d[:, 0] * c[:, 3] - d[:, 1] * c[:, 2]) / det
    ty = (-d[:, 0] * c[:, 1] + d[:, 1] * c[:, 0]) / det
    qf = tx * d[:, 0] + ty * d[:, 1]
    Gref = torch.exp(-0.5 * qf)
    assert float((G.double() - Gref).abs().max()) < 2e-6
    assert float(((vx.double() - tx).abs() / (tx.abs() + 1e-3 * tx.abs().mean())).max()) < 5e-4
    assert float(((vy.double() - ty).abs() / (ty.abs() + 1e-3 * ty.abs().mean())).max()) < 5e-4
    # bounding box is conservative: whenever a*G >= 1/255 the query lies inside (hx, hy)
    a = torch.clamp(alpha, max=0.99).double()
    vis = a * Gref >= (1.0 / 255.0) * (1 - 1e-6)
    inside = (d[:, 0].abs() <= hx.double()) & (d[:, 1].abs() <= hy.double())
    assert bool((inside | ~vis).all())
    # ... and tight: extent^2 == 2 ln(255 a) * cov_ii (1e-3)
    qmax = 2 * torch.log(255 * a)
    pos = qmax > 0
    assert torch.allclose(hx.double()[pos] ** 2, (qmax * c[:, 0])[pos], rtol=2e-3, atol=1e-10)
    assert torch.allclose(hy.double()[pos] ** 2, (qmax * c[:, 3])[pos], rtol=2e-3, atol=1e-10)


def test_splat_never_visible_and_non_pd(hostmath):
    m2 = torch.zeros(3, 2)
    cov = torch.tensor([[1e-4, 0, 0, 1e-4], [1e-4, 2e-4, 2e-4, 1e-4], [float("nan"), 0, 0, 1e-4]])
    alpha = torch.tensor([0.003, 0.9, 0.9])  # 0.003 < 1/255: never visible; 2nd: det < 0; 3rd: NaN
    q = torch.zeros(3, 2)
    outs = [torch.empty(3) for _ in range(6)]
    hostmath.hm_splat_eval(3, fp(m2), fp(cov), fp(alpha), fp(q), *[fp(o) for o in outs])
    aG, G, vx, vy, hx, hy = outs
    assert float(hx[0]) < 0 and float(hy[0]) < 0
    assert float(aG[1]) == 0.0 and float(hx[1]) < 0
    assert float(aG[2]) == 0.0 and float(hx[2]) < 0


@pytest.mark.parametrize("C", [1, 2, 3, 4])
def test_sh_basis_matches_oracle(hostmath, oracle_mod, C):
    g = torch.Generator().manual_seed(C)
    c9 = torch.randn(9, generator=g)
    for _ in range(20):
        pos = torch.randn(2, generator=g) * 0.4
        ours = torch.zeros(16)
        hostmath.hm_sh_basis(C, fp(pos), fp(c9), fp(ours))
        d, ref = torch.zeros(3), torch.zeros(16)
        pos3 = torch.tensor([pos[0], pos[1], 1.0])
        oracle_mod.lib().orc_pixel_dir(fp(pos3), fp(c9), fp(d))
        oracle_mod.lib().orc_sh_basis(fp(d), fp(ref), C)
        assert torch.allclose(ours[: C * C], ref[: C * C], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("act", [0, 7, 1, 6])
def test_in_kernel_activations_match_torch_chain_rule(hostmath, oracle_mod, act):
    """raw leaves in, raw-leaf gradients out.  torch applies exp / sigmoid / sigmoid in the
    `svec` / `alpha` / `color` properties (gs/gaussian_splatting.py:113-123); the fused front end evaluates them in
    registers and the projection backward applies their chain rule."""
    sc, cam, c2w = _scene(N=2500)
    g = torch.Generator().manual_seed(11)
    N = sc.N
    svec_raw = (torch.log(sc.svec) if act & 1 else sc.svec).clone().requires_grad_()
    alpha_raw = (torch.logit(sc.alpha.view(-1)) if act & 2 else sc.alpha.view(-1)).clone().requires_grad_()
    color_raw = (torch.logit(sc.color.clamp(0.02, 0.98)) if act & 4 else sc.color).clone().requires_grad_()
    mean = sc.mean.clone().requires_grad_()
    q = sc.qvec.clone().requires_grad_()
    s = torch.exp(svec_raw) if act & 1 else svec_raw
    a = torch.sigmoid(alpha_raw) if act & 2 else alpha_raw
    c = torch.sigmoid(color_raw) if act & 4 else color_raw
    m2, cov, _, _ = oracle_mod.project_gaussians(mean, q, s, c2w, True)
    gm2, gcov = torch.randn(N, 2, generator=g), torch.randn(N, 2, 2, generator=g)
    ga, gc = torch.randn(N, generator=g), torch.randn(N, 3, generator=g)
    ((m2 * gm2).sum() + (cov * gcov).sum() + (a * ga).sum() + (c * gc).sum()).backward()
    o_m2, o_cov, o_a, o_c = torch.empty(N, 2), torch.empty(N, 4), torch.empty(N), torch.empty(N, 3)
    gx, gq, gs, gar, gcr = torch.empty(N, 3), torch.empty(N, 4), torch.empty(N, 3), torch.empty(N), torch.empty(N, 3)
    hostmath.hm_front_end_raw(N, act, fp(sc.mean), fp(sc.qvec), fp(svec_raw.detach()), fp(alpha_raw.detach()),
                              fp(color_raw.detach().contiguous()), fp(c2w.contiguous()), fp(o_m2), fp(o_cov), fp(o_a),
                              fp(o_c), fp(gm2), fp(gcov.reshape(N, 4).contiguous()), fp(ga), fp(gc), fp(gx), fp(gq),
                              fp(gs), fp(gar), fp(gcr))
    assert torch.allclose(o_a, a.detach(), rtol=1e-6, atol=1e-7)
    assert torch.allclose(o_c, c.detach(), rtol=1e-6, atol=1e-7)
    assert torch.allclose(o_m2, m2.detach(), rtol=2e-5, atol=1e-6)
    for ours, ref, nm in ((gx, mean.grad, "mean"), (gq, q.grad, "qvec"), (gs, svec_raw.grad, "svec_raw"),
                          (gar, alpha_raw.grad, "alpha_raw"), (gcr, color_raw.grad, "color_raw")):
        rel = (ours - ref).norm() / ref.norm()
        assert float(rel) < 1e-4, (nm, float(rel))
