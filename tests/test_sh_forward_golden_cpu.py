"""The oracle's restatement of the SH view (`oracle.render_view(sh=...)`) versus the reference's OWN
`SHRenderer.forward` + `_render_sh` / `_render_sh_bg` autograd Functions executed unmodified over the same CPU
kernels (tests/golden/make_render_one_golden.py): image, parameter gradients, mask, duplicate count, mean2d gradient,
including the nine-float read of the [3,4] c2w by the SH kernels (the Functions pass `c2w` through untouched)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sh_forward_ref.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("tag", ["c", "d"])
def test_render_view_restates_sh_forward(gold, oracle_mod, tag):
    fx, fy, cx, cy, w, h, near, far = gold[f"{tag}_cam"].tolist()
    ocam = oracle_mod.Cam(fx, fy, cx, cy, int(w), int(h), near, far)
    C, with_bg = gold[f"{tag}_C"].tolist()
    leaves = {k: gold[f"{tag}_in_{k}"].clone().requires_grad_() for k in ("mean", "qvec", "svec", "sh_coeffs", "alpha")}
    out = oracle_mod.render_view(leaves["mean"], leaves["qvec"], leaves["svec"], leaves["alpha"], gold[f"{tag}_c2w"],
                                 ocam, sh=leaves["sh_coeffs"], C=C, bg_rgb=gold[f"{tag}_bg_rgb"] if with_bg else None)
    ref = gold[f"{tag}_rgb"]
    assert torch.allclose(out["rgb"].reshape(ref.shape), ref, rtol=1e-6, atol=1e-7)
    (out["rgb"].reshape(ref.shape) * gold[f"{tag}_w"]).sum().backward()
    for k, v in leaves.items():
        r = gold[f"{tag}_grad_{k}"]
        assert float((v.grad - r).norm() / r.norm()) < 1e-6, k
    aux = out["aux"]
    assert torch.equal(aux["mask"], gold[f"{tag}_mask"]) and aux["D"] == int(gold[f"{tag}_N_with_dub"])
    assert torch.equal(gold[f"{tag}_cnt"], gold[f"{tag}_mask"].float())
    rg = gold[f"{tag}_mean2d_grad"]
    assert float((aux["mean2d"].grad - rg).norm() / rg.norm()) < 1e-6
