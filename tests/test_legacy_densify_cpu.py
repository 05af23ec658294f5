"""`densify(step)` with `use_legacy: True` -- the rule the reference's top-level experiment configs select
(conf/base.yaml, corgi.yaml, shrink_then_densify.yaml) -- against a fixture produced by executing the reference's OWN
`densify` / `densify_legacy` / compactness methods (tests/golden/make_legacy_densify_golden.py): the trace of N over the
step gates, every parameter row in the reference's order (unsplit rows, clones, children), and the optimizer reset
(`set_optimizer(opt_cfg, step)` inside densify_legacy drops every Adam moment)."""
import os

import numpy as np
import pytest
import torch

from gsgen_b200.store import GaussianStore

GOLD = os.path.join(os.path.dirname(__file__), "golden", "densify_legacy.npz")
FIELDS = ("mean", "qvec", "svec", "alpha", "color")


def _load():
    z = np.load(GOLD)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _store(g, oracle_mod, capacity=None):
    st = GaussianStore({f: g[f"s0_{f}"] for f in FIELDS}, C=None, device="cpu", knn_fn=oracle_mod.knn_points,
                       capacity=capacity)
    for f in FIELDS:
        n = st.N
        st._rows(st.exp_avg, f, n).copy_(g[f"s0_{f}_exp_avg"].reshape(n, -1))
        st._rows(st.exp_avg_sq, f, n).copy_(g[f"s0_{f}_exp_avg_sq"].reshape(n, -1))
    assert float(st.exp_avg.abs().max()) > 0  # the moments the reset has to drop
    return st


@pytest.mark.parametrize("capacity", [None, 4096])
@pytest.mark.parametrize("kind", ["official", "shrink_then_compatness", "compatness"])
def test_legacy_dispatcher_matches_reference(oracle_mod, kind, capacity):
    g = _load()
    st = _store(g, oracle_mod, capacity)
    N0 = st.N
    cfg = dict(enabled=True, type=kind, warm_up=100, end=1000, period=100, use_legacy=True, K=2, surface_shrink=1.25,
               mean2d_thresh=0.02, split_thresh=0.03, n_splits=2, split_shrink=0.8, noise=g[f"noise_{kind}"])

    class Opt:  # what FlatAdam exposes to the store: an update counter and rebind()
        n_steps = 7

        def rebind(self, *a):
            pass

    st.optimizer = Opt()
    trace = []
    for step in (0, 100, 150):
        if st.N == N0:
            st.mean_2d_grad_accum, st.cnt = g["accum"].clone(), g["cnt"].clone()
        n_before = st.N
        res = st.densify_step(step, cfg)
        trace.append([step, n_before, st.N])
        if step == 100:
            n_noise = g[f"noise_{kind}"].shape[0]
            assert res is not None and res[0] == n_noise // 2 and len(res) == (2 if kind == "official" else 3)
            assert st.optimizer.n_steps == 0  # the re-created optimizer counts from zero
            assert float(st.cnt.abs().max()) == 0.0 and st.cnt.shape[0] == st.N
        else:
            assert res is None
    assert trace == g[f"trace_{kind}"].tolist(), (trace, g[f"trace_{kind}"].tolist())
    n = st.N
    n_split = g[f"noise_{kind}"].shape[0] // 2
    for f in FIELDS:
        ours, ref = st._rows(st.flat_param, f, n), g[f"s1_{kind}_{f}"].reshape(n, -1)
        if kind == "official":
            keep = N0 - n_split  # unsplit rows + clones are copies: exact
            n_clone = n - keep - 2 * n_split
            assert torch.equal(ours[: keep + n_clone], ref[: keep + n_clone]), f
        if f == "svec":  # log(...) of quantities that may be tiny: compare the scale itself
            ours, ref = torch.exp(ours), torch.exp(ref)
        assert torch.allclose(ours, ref, rtol=2e-5, atol=2e-6), (kind, f, float((ours - ref).abs().max()))
        # no Adam state survives (has_state == 0 for every group in the reference)
        assert int(g[f"s1_{kind}_has_state"].sum()) == 0
        assert float(st._rows(st.exp_avg, f, n).abs().max()) == 0.0
        assert float(st._rows(st.exp_avg_sq, f, n).abs().max()) == 0.0


def test_legacy_rule_is_reached_only_with_use_legacy(oracle_mod):
    g = _load()
    st = _store(g, oracle_mod)
    st.mean_2d_grad_accum, st.cnt = g["accum"].clone(), g["cnt"].clone()
    cfg = dict(enabled=True, type="official", warm_up=100, end=1000, period=100, use_legacy=False, mean2d_thresh=0.02,
               split_thresh=0.03, n_splits=2, split_shrink=0.8)
    res = st.densify_step(100, cfg)
    assert res is not None and float(st.exp_avg.abs().max()) > 0  # the non-legacy rules keep the moments of old rows
