"""-m gpu integration of the §8(f) rows around the rasterizer: raw leaves in a GaussianStore (capacity arena) ->
render_view(raw_params=True, grad_sink=...) -> FlatAdam.step -> densify / prune -> render again.  The row
operations themselves are pinned on CPU against the reference's own methods (tests/test_store_cpu.py); here the same
operations run on the device and must agree with a CPU twin, and the optimizer / renderer must follow the arena
through a change of N and a re-allocation."""
import pytest
import torch

from gsgen_b200.scenes import make_scene
from gsgen_b200.store import GaussianStore

pytestmark = pytest.mark.gpu
DEV = "cuda"
LR = {"mean": [0.005, 3.0e-05, 15000, "exp"], "svec": [0.003, 0.001, 15000, "exp"], "qvec": 0.003, "color": 0.01,
      "alpha": 0.003}


def _raw_params(sc):
    return dict(mean=sc.mean, qvec=sc.qvec, svec=torch.log(sc.svec), alpha=torch.logit(sc.alpha.reshape(-1)),
                color=torch.logit(sc.color.clamp(0.02, 0.98)))


def _twin(st: GaussianStore) -> GaussianStore:
    """CPU copy of a store (same rows, moments and statistics)"""
    tw = GaussianStore({k: v.detach().cpu() for k, v in st.params.items()}, st.C, "cpu", capacity=st.cap)
    for name in st.params:
        for src, dst in ((st.exp_avg, tw.exp_avg), (st.exp_avg_sq, tw.exp_avg_sq)):
            tw._rows(dst, name, tw.N).copy_(st._rows(src, name, st.N).cpu())
    for s in ("max_radii2d", "mean_2d_grad_accum", "cnt"):
        setattr(tw, s, getattr(st, s).cpu().clone())
    return tw


def _assert_same(st, tw, tol):
    assert st.N == tw.N
    for name in st.params:
        for a, b, what in ((st.flat_param, tw.flat_param, "param"), (st.exp_avg, tw.exp_avg, "exp_avg"),
                           (st.exp_avg_sq, tw.exp_avg_sq, "exp_avg_sq")):
            x, y = st._rows(a, name, st.N).cpu(), tw._rows(b, name, tw.N)
            assert torch.allclose(x, y, rtol=tol, atol=tol), (name, what, float((x - y).abs().max()))


def test_store_render_adam_densify_prune_round():
    from gsgen_b200.rasterizer import render_view

    sc = make_scene("c1", N=3000, reso=128)
    cam, c2w = sc.cams[0], sc.c2ws[0]
    st = GaussianStore(_raw_params(sc), None, DEV)  # capacity == N: the densify below has to re-allocate
    opt = st.make_optimizer(LR)
    g = torch.Generator().manual_seed(3)
    w = torch.randn(cam.h, cam.w, 3, generator=g).to(DEV)

    def train_step(step):
        st.zero_grad()
        p = st.params
        out = render_view(p["mean"], p["qvec"], p["svec"], p["alpha"], c2w, cam, color=p["color"], rgb_only=True,
                          raw_params=True, grad_sink=st.grad_views)
        out["rgb"].backward(gradient=w)
        aux = out["aux"]
        st.update_densify_info(aux["mask"], aux["mean2d_grad"], aux["radii2d"])
        before = st.flat_param.clone()
        opt.step(step)
        assert torch.isfinite(st.flat_param).all()
        return out, float((st.flat_param - before).abs().max())

    for step in range(3):
        out, moved = train_step(step)
        assert 0.0 < moved <= 4 * 0.01  # Adam moves a parameter by at most lr*(1-b1)/sqrt(1-b2) ~ 3.2 lr per step
    assert float(st.cnt.max()) == 3.0 and float(st.max_radii2d.max()) > 0.0
    assert float(st.exp_avg.abs().max()) > 0.0

    # ---- densify ("official") + prune on the device vs the same operations on a CPU twin
    tw = _twin(st)
    n0 = st.N
    noise = torch.randn(4 * n0, 3, generator=g)  # more rows than any selection needs
    grads = tw.mean_2d_grad_accum / tw.cnt
    grads[grads.isnan()] = 0.0
    thresh = float(grads.median())
    sel_split = int(((grads >= thresh) & (tw.svec_act.max(dim=1).values > 0.02)).sum())
    res_dev = st.densify_official(thresh, 0.02, 2, 0.8, noise=noise[: 2 * sel_split].to(DEV))
    res_cpu = tw.densify_official(thresh, 0.02, 2, 0.8, noise=noise[: 2 * sel_split])
    assert res_dev == res_cpu and st.N == tw.N and st.N != n0 and st.cap > n0
    _assert_same(st, tw, 1e-5)
    for store in (st, tw):
        store.max_radii2d = torch.linspace(0.0, 1.2, store.N, device=store.device)
    assert st.prune(1.0, 0.05, 0.6) == tw.prune(1.0, 0.05, 0.6)
    _assert_same(st, tw, 1e-5)
    assert opt.flat_param.data_ptr() == st.flat_param.data_ptr() and opt.exp_avg.data_ptr() == st.exp_avg.data_ptr()

    # ---- the renderer and the optimizer follow the new N / the new buffers
    out, moved = train_step(3)
    assert out["aux"]["mask"].shape[0] == st.N and moved > 0.0
    # dead capacity rows never move
    for name in st.params:
        assert float(st._rows(st.flat_param, name, st.cap - st.N, st.N).abs().max()) == 0.0
        assert float(st._rows(st.exp_avg, name, st.cap - st.N, st.N).abs().max()) == 0.0
