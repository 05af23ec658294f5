"""The oracle's restatement of one whole view (`oracle.render_view`, which every -m gpu parity test of the fused path
compares against) versus the reference's OWN `GaussianSplattingRenderer.render_one` + autograd Functions, executed
unmodified over the same CPU kernels (tests/golden/make_render_one_golden.py): images, all parameter gradients, the
background gradient, the frustum mask, the duplicate count, the densification gradient of mean2d and max_radii2d."""
import os

import numpy as np
import pytest
import torch

from tests.util import fp

GOLD = os.path.join(os.path.dirname(__file__), "golden", "render_one_ref.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("tag", ["a", "b"])
def test_render_view_restates_render_one(gold, oracle_mod, hostmath, tag):
    fx, fy, cx, cy, w, h, near, far = gold[f"{tag}_cam"].tolist()
    ocam = oracle_mod.Cam(fx, fy, cx, cy, int(w), int(h), near, far)
    leaves = {k: gold[f"{tag}_in_{k}"].clone().requires_grad_() for k in ("mean", "qvec", "svec", "color", "alpha")}
    bg = gold[f"{tag}_in_bg"].clone().requires_grad_()
    out = oracle_mod.render_view(leaves["mean"], leaves["qvec"], leaves["svec"], leaves["alpha"], gold[f"{tag}_c2w"],
                                 ocam, color=leaves["color"], bg=bg, rgb_only=False)
    for k in ("rgb", "depth", "opacity", "z_var"):
        ref = gold[f"{tag}_{k}"]
        assert torch.allclose(out[k].reshape(ref.shape), ref, rtol=1e-6, atol=1e-7), k
    sum((out[k].reshape(gold[f"{tag}_w_{k}"].shape) * gold[f"{tag}_w_{k}"]).sum()
        for k in ("rgb", "depth", "opacity", "z_var")).backward()
    for k, v in leaves.items():
        ref = gold[f"{tag}_grad_{k}"]
        assert float((v.grad - ref).norm() / ref.norm()) < 1e-6, k
    assert torch.allclose(bg.grad, gold[f"{tag}_grad_bg"], rtol=1e-6, atol=1e-8)
    aux = out["aux"]
    assert torch.equal(aux["mask"], gold[f"{tag}_mask"])
    assert aux["D"] == int(gold[f"{tag}_N_with_dub"])
    ref_g = gold[f"{tag}_mean2d_grad"]
    assert float((aux["mean2d"].grad - ref_g).norm() / ref_g.norm()) < 1e-6
    # max_radii2d side effect (zero-initialised, so it equals this view's radii under the mask): the kernels' radius2d
    cov = aux["cov2d"].detach().reshape(-1, 4).contiguous()
    radii = torch.empty(cov.shape[0])
    hostmath.hm_radius2d(cov.shape[0], fp(cov), fp(radii))
    ref_r = gold[f"{tag}_max_radii2d"]
    assert float(ref_r[~gold[f"{tag}_mask"]].abs().max() if (~gold[f"{tag}_mask"]).any() else 0.0) == 0.0
    assert torch.allclose(radii, ref_r[gold[f"{tag}_mask"]], rtol=1e-5, atol=1e-9)
