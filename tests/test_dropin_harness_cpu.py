"""CPU check of the drop-in harness (tests/refpy.py): the reference's own compiled Python (oracle/_ref/ref_py.bin) run
over the CPU oracle behind the `_gs` names must reproduce the committed fixtures that the golden generator produced from
the reference's SOURCE with the same adapter (tests/golden/render_one_ref.npz, sh_forward_ref.npz) -- so the harness
the `-m gpu` drop-in test uses (same code objects, `_backend` = `_gs.so` / libgsb200) assembles the reference correctly.
Skipped where oracle/_ref/ref_py.bin does not exist (it is built from /root/reference by __graft_entry__.build())."""
import os
import sys

import numpy as np
import pytest
import torch

from gsgen_b200.scenes import make_scene
from tests import refpy

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="module")
def cpu_ns(oracle_mod):
    entries = refpy.load_entries()
    if entries is None:
        pytest.skip("oracle/_ref/ref_py.bin not built (needs /root/reference: python oracle/build_ref_py.py)")
    sys.path.insert(0, GOLD)
    try:
        import make_render_one_golden as mrg
    finally:
        sys.path.pop(0)
    return refpy.namespace(mrg.OracleBackend(oracle_mod), entries, torch_mod=mrg.TorchProxy())


def test_render_one_over_the_oracle_reproduces_the_fixture(cpu_ns):
    g = _load("render_one_ref.npz")
    sc = make_scene("c1", N=1000, reso=96)
    cam, c2w = sc.cams[0], sc.c2ws[0]
    assert torch.equal(sc.mean, g["a_in_mean"]) and torch.equal(c2w, g["a_c2w"])
    weights = {k: g[f"a_w_{k}"] for k in ("rgb", "depth", "opacity", "z_var")}
    outs, grads, side = refpy.run_render_one(cpu_ns, sc, cam, c2w, "cpu", g["a_in_bg"], weights)
    for k in ("rgb", "depth", "opacity", "z_var", "T"):
        assert torch.equal(outs[k], g[f"a_{k}"]), k
    for k in ("mean", "qvec", "svec", "color", "alpha", "bg"):
        assert torch.equal(grads[k], g[f"a_grad_{k}"]), k
    assert torch.equal(side["max_radii2d"], g["a_max_radii2d"]) and torch.equal(side["mask"], g["a_mask"])
    assert torch.equal(side["mean2d_grad"], g["a_mean2d_grad"]) and side["N_with_dub"] == int(g["a_N_with_dub"])


def test_sh_forward_over_the_oracle_reproduces_the_fixture(cpu_ns):
    g = _load("sh_forward_ref.npz")
    sc = make_scene("c3", N=1200, reso=80)
    sc.svec = (sc.svec * 4.0).contiguous()
    cam, c2w = sc.cams[0], sc.c2ws[0]
    assert torch.equal(sc.sh, g["c_in_sh_coeffs"])
    rgb, grads, side = refpy.run_sh_forward(cpu_ns, sc, cam, c2w, "cpu", 4, False, g["c_w"])
    assert torch.equal(rgb, g["c_rgb"])
    for k in ("mean", "qvec", "svec", "sh_coeffs", "alpha"):
        assert torch.equal(grads[k], g[f"c_grad_{k}"]), k
    assert torch.equal(side["mask"], g["c_mask"]) and torch.equal(side["cnt"], g["c_cnt"])
    assert side["N_with_dub"] == int(g["c_N_with_dub"])
