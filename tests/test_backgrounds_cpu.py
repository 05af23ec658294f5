"""Background models (gsgen_b200/backgrounds.py) against the reference's OWN classes (gs/backgrounds.py compiled out of
the file with `ast` in the dev container; skipped where /root/reference is absent) draw for draw, and their wiring into
the trainer-facing renderer (`cfg.background` -> setup_bg, `.bg`, train / eval, checkpoint `bg` state)."""
import ast
import os
import random

import pytest
import torch
import torch.nn as nn

from gsgen_b200 import backgrounds as B

REF = "/root/reference/gs/backgrounds.py"


class Cfg(dict):
    __getattr__ = dict.__getitem__


def _ref_classes():
    if not os.path.exists(REF):
        pytest.skip("reference sources not present")
    from einops import repeat

    ns = {"torch": torch, "nn": nn, "random": random, "repeat": repeat, "tcnn_capable": False}
    tree = ast.parse(open(REF).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return ns


@pytest.mark.parametrize("kind,extra", [("random", dict(range=[0.2, 0.9])), ("fixed", dict(color=[0.1, 0.5, 0.9])),
                                        ("random", dict(range=[0.0, 1.0], random_aug=True, random_aug_prob=0.5))])
def test_backgrounds_match_reference_draw_for_draw(kind, extra):
    ns = _ref_classes()
    cfg = Cfg(type=kind, device="cpu", random_aug=False, random_aug_prob=0.0)
    cfg.update(extra)
    ref = {"random": ns["RandomBackground"], "fixed": ns["FixedBackground"]}[kind](cfg)
    ours = B.make_background(cfg)
    dirs = torch.randn(5, 7, 3)
    for mode in (True, False, True):
        ref.train(mode); ours.train(mode)
        for _ in range(6):
            random.seed(11); torch.manual_seed(11)
            a = ref(dirs)
            random.seed(11); torch.manual_seed(11)
            b = ours(dirs)
            assert a.shape == b.shape == (5, 7, 3) and torch.equal(a, b)
    if kind == "fixed":  # the colour is a parameter: gradient of sum(bg * w) is the sum of w over the pixels
        w = torch.randn(5, 7, 3)
        (ours(dirs) * w).sum().backward()
        assert torch.allclose(ours.bg_color.grad, w.sum(dim=(0, 1)))
        assert set(ours.state_dict()) == set(ref.state_dict()) == {"bg_color"}


def test_learned_const_and_unknown_types():
    bg = B.make_background({"type": "learned_const", "initial_color": [0.3, 0.4, 0.5]})
    out = bg(torch.zeros(4, 6, 3))
    assert out.shape == (4, 6, 3) and torch.equal(out[2, 3], torch.tensor([0.3, 0.4, 0.5])) and bg.bg_color.requires_grad
    with pytest.raises(NotImplementedError):
        B.make_background({"type": "mlp"})
    with pytest.raises(NotImplementedError):
        B.make_background({"type": "sky"})


def test_renderer_builds_its_background_from_cfg(oracle_mod):
    from gsgen_b200.splatting import GaussianSplattingRenderer
    from gsgen_b200.store import quat_to_rotmat

    g = torch.Generator().manual_seed(0)
    N = 40
    init = {"mean": torch.randn(N, 3, generator=g), "qvec": torch.randn(N, 4, generator=g),
            "svec": torch.rand(N, 3, generator=g) * 0.1 + 0.01, "color": torch.rand(N, 3, generator=g) * 0.8 + 0.1,
            "alpha": torch.rand(N, generator=g) * 0.8 + 0.1}
    cfg = {"background": {"type": "fixed", "color": [0.2, 0.3, 0.4], "device": "cpu", "random_aug": False,
                          "random_aug_prob": 0.0}}
    r = GaussianSplattingRenderer(cfg, init, device="cpu", render_fn=lambda *a, **k: None)
    assert isinstance(r.bg, B.FixedBackground) and r.bg is r.background
    r.eval()
    assert not r.bg.training
    r.train()
    assert r.bg.training
    assert "bg_color" in r.get_params_for_save()["bg"]
    # the attributes guidance / export code reads (gs/gaussian_splatting.py:146-156)
    assert torch.allclose(r.cov, r.svec.unsqueeze(-2) * quat_to_rotmat(r.qvec))
    assert torch.equal(r.principal_axis, r.rotmat) and r.cov.shape == (N, 3, 3)
    # no cfg.background and no argument: black (None), as before
    r2 = GaussianSplattingRenderer({}, init, device="cpu", render_fn=lambda *a, **k: None)
    assert r2.bg is None
