"""Self-consistency of the oracle (independent of the golden pin): its analytic backward agrees with finite
differences of its own forward, the whole-view wrapper matches the op-by-op calls, D_eff statistics are sane."""
import pytest
import torch

from gsgen_b200.scenes import make_scene
from tests.util import ocam_of


def _view(oracle, N=400, reso=32, scale=2.0):
    sc = make_scene("c1", N=N, reso=reso)
    sc.svec = (sc.svec * scale).contiguous()
    cam, c2w = sc.cams[0], sc.c2ws[0]
    return sc, cam, c2w


def test_rgb_backward_matches_finite_differences(oracle_mod):
    o = oracle_mod
    sc, cam, c2w = _view(o)
    ocam = ocam_of(cam)
    g = torch.Generator().manual_seed(0)
    w = torch.randn(cam.h, cam.w, 3, generator=g)

    def loss(color, alpha):
        out = o.render_view(sc.mean, sc.qvec, sc.svec, alpha, c2w, ocam, color=color)
        return (out["rgb"].double() * w.double()).sum()

    color = sc.color.clone().requires_grad_()
    alpha = sc.alpha.clone().requires_grad_()
    loss(color, alpha).backward()
    # colour is a linear input: central differences are exact up to fp32 rounding
    idx = torch.argsort(color.grad.abs().sum(dim=1), descending=True)[:5]
    for i in idx.tolist():
        for c in range(3):
            e = torch.zeros_like(sc.color); e[i, c] = 1e-2
            fd = (loss(sc.color + e, sc.alpha) - loss(sc.color - e, sc.alpha)) / 2e-2
            assert abs(float(fd) - float(color.grad[i, c])) <= 2e-3 * max(1.0, abs(float(fd))), (i, c)
    # opacity: smooth away from the 1/255 and T<thresh discontinuities; check the strongest entries
    idx = torch.argsort(alpha.grad.abs(), descending=True)[:5]
    for i in idx.tolist():
        e = torch.zeros_like(sc.alpha); e[i] = 2e-3
        fd = (loss(sc.color, sc.alpha + e) - loss(sc.color, sc.alpha - e)) / 4e-3
        assert abs(float(fd) - float(alpha.grad[i])) <= 3e-2 * max(1.0, abs(float(fd))), i


def test_whole_view_equals_op_chain_and_stats(oracle_mod):
    o = oracle_mod
    sc, cam, c2w = _view(o, N=2000, reso=64, scale=3.0)
    ocam = ocam_of(cam)
    out = o.render_view(sc.mean, sc.qvec, sc.svec, sc.alpha, c2w, ocam, color=sc.color, rgb_only=False)
    aux = out["aux"]
    rgb, T, stats = o.composite_rgb_fwd(aux["mean2d"].contiguous(), aux["cov2d"].contiguous(),
                                        sc.color[aux["mask"]].contiguous(), sc.alpha[aux["mask"]].contiguous(),
                                        aux["start"], aux["end"], aux["ids"], aux["topleft"], aux["cfg"])
    assert torch.equal(rgb, out["rgb"])
    # opacity = sum w = 1 - T  (blend identity)
    assert float((out["opacity"].squeeze(-1) + T - 1).abs().max()) < 1e-5
    D = aux["D"]
    assert 0 < stats["d_eff"] <= D
    assert stats["pairs_blended"] <= stats["pairs_evaluated"] <= 256 * D
    # sorted: within each tile depths ascend
    dp = aux["depth"].view(-1)
    for t in range(aux["start"].numel()):
        s, e = int(aux["start"][t]), int(aux["end"][t])
        if s >= 0 and e - s > 1:
            d = dp[aux["ids"][s:e].long()]
            assert bool((d[1:] >= d[:-1]).all())


def test_sh_bg_semantics(oracle_mod):
    o = oracle_mod
    sc, cam, c2w = _view(o, N=300, reso=48, scale=1.0)
    ocam = ocam_of(cam)
    bg = torch.tensor([0.2, 0.4, 0.6])
    a = o.render_view(sc.mean, sc.qvec, sc.svec, sc.alpha, c2w, ocam, sh=sc.sh, C=1)
    b = o.render_view(sc.mean, sc.qvec, sc.svec, sc.alpha, c2w, ocam, sh=sc.sh, C=1, bg_rgb=bg)
    aux = a["aux"]
    _, T, _ = o.composite_sh_fwd(aux["mean2d"].contiguous(), aux["cov2d"].contiguous(),
                                 sc.sh[aux["mask"]].contiguous(), sc.alpha[aux["mask"]].contiguous(), aux["start"],
                                 aux["end"], aux["ids"], aux["topleft"], c2w, 1, aux["cfg"])
    assert torch.allclose(b["rgb"], a["rgb"] + T.unsqueeze(-1) * bg, atol=1e-6)
