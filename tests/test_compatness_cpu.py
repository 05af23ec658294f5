"""Compactness-based densification and the neighbour penalties (gs/gaussian_splatting.py:634-743,
:1032-1094; utils/ops.py:103-158) on the CPU:

  * the grid search the CUDA kernel runs (gsgen_b200/csrc/knn_grid.cuh compiled with g++, tests/hostmath) against the
    oracle's brute-force restatement of pytorch3d.ops.knn_points -- indices AND distances bit for bit, on clouds chosen
    to break a grid: clusters, a plane, a line, duplicates, coincident points, an outlier, queries outside the box;
  * `gsgen_b200.knn` / `GaussianStore` / `GaussianSplattingRenderer` against tests/golden/compatness.npz, produced by
    executing the reference's OWN methods (tests/golden/make_compatness_golden.py) over the same brute-force search.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from gsgen_b200.knn import K_nearest_neighbors, distance_to_gaussian_surface, nearest_neighbor
from gsgen_b200.store import GaussianStore, quat_to_rotmat

GOLD = os.path.join(os.path.dirname(__file__), "golden", "compatness.npz")
FIELDS = ("mean", "qvec", "svec", "alpha", "color")


def _load():
    z = np.load(GOLD)
    return {k: torch.from_numpy(z[k]) for k in z.files}


# ---- the grid search itself ------------------------------------------------------------------------------------------
def _grid_knn(hostmath, pts, K, queries=None, max_cells=None):
    pts = np.ascontiguousarray(pts, np.float32)
    n = pts.shape[0]
    q = None if queries is None else np.ascontiguousarray(queries, np.float32)
    nq = n if q is None else q.shape[0]
    idx = np.empty((nq, K), np.int64)
    d2 = np.empty((nq, K), np.float32)
    stats = np.zeros(2, np.int32)
    vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    rc = hostmath.hm_knn(n, vp(pts), nq, vp(q), K, ctypes.c_uint(max_cells or 4 * n + 64), vp(idx), vp(d2), vp(stats))
    assert rc == 0
    return torch.from_numpy(d2), torch.from_numpy(idx), stats


def _clouds():
    rng = np.random.default_rng(0)
    out = {"ball": rng.normal(size=(3000, 3)) * 0.5}
    c = rng.normal(size=(2000, 3)) * 0.01
    c[:1000] += 5.0
    out["two_clusters"] = c
    p = rng.uniform(-1, 1, size=(1500, 3))
    p[:, 2] = 0.25
    out["plane"] = p
    p = rng.uniform(-1, 1, size=(800, 3))
    p[400:] = p[:400]
    out["duplicates"] = p
    out["line"] = np.stack([np.linspace(0, 1, 500), np.zeros(500), np.zeros(500)], 1)
    out["coincident"] = np.ones((40, 3))
    out["three"] = rng.normal(size=(3, 3))
    out["one"] = rng.normal(size=(1, 3))
    o = rng.normal(size=(2000, 3)) * 0.1
    o[0] = [100, 100, 100]
    out["outlier"] = o
    out["lattice"] = np.stack(np.meshgrid(*[np.arange(10.0)] * 3, indexing="ij"), -1).reshape(-1, 3)  # ties everywhere
    return {k: v.astype(np.float32) for k, v in out.items()}


@pytest.mark.parametrize("name", sorted(_clouds()))
def test_grid_search_equals_brute_force(hostmath, oracle_mod, name):
    pts = _clouds()[name]
    for K in (1, 2, 4, 5, 9, 17):
        d2, idx, stats = _grid_knn(hostmath, pts, K)
        d2_ref, idx_ref = oracle_mod.knn_points(None, torch.from_numpy(pts), K)
        assert torch.equal(idx, idx_ref), (name, K)
        assert torch.equal(d2, d2_ref), (name, K)
    if pts.shape[0] >= 4:  # self query: column 0 is a point at distance 0 (the point itself unless it has a duplicate)
        assert float(d2[:, 0].max()) == 0.0


def test_grid_search_with_external_queries_and_tiny_cell_budget(hostmath, oracle_mod):
    rng = np.random.default_rng(3)
    pts = (rng.normal(size=(2500, 3)) * 0.5).astype(np.float32)
    q = (rng.normal(size=(400, 3)) * 2.0).astype(np.float32)  # most of them outside the grid box
    for K in (1, 3, 8):
        d2, idx, _ = _grid_knn(hostmath, pts, K, q)
        d2_ref, idx_ref = oracle_mod.knn_points(torch.from_numpy(q), torch.from_numpy(pts), K)
        assert torch.equal(idx, idx_ref) and torch.equal(d2, d2_ref)
    d2, idx, stats = _grid_knn(hostmath, pts, 4, max_cells=10)  # degenerate grid: still exact
    d2_ref, idx_ref = oracle_mod.knn_points(None, torch.from_numpy(pts), 4)
    assert stats[0] <= 10 and torch.equal(idx, idx_ref) and torch.equal(d2, d2_ref)
    # fewer points than K: the tail is (-1, +inf)
    d2, idx, _ = _grid_knn(hostmath, pts[:3], 5)
    assert idx[:, 3:].eq(-1).all() and torch.isinf(d2[:, 3:]).all() and idx[:, :3].ge(0).all()


def test_grid_search_is_local_for_a_uniform_cloud(hostmath):
    """the point of the grid: a query of a uniform cloud finishes within a few shells (27-125 cells of ~2 points)"""
    rng = np.random.default_rng(4)
    pts = rng.uniform(-1, 1, size=(20000, 3)).astype(np.float32)
    _, _, stats = _grid_knn(hostmath, pts, 4)
    assert stats[0] > 5000 and stats[1] <= 6, stats


# ---- host mirrors vs the reference's own functions ------------------------------------------------------------------
def test_knn_wrappers_and_surface_distance_match_reference(oracle_mod):
    g = _load()
    mean, qvec, svec = g["in_mean"], g["in_qvec"], torch.exp(g["in_svec"])
    # s0 = the inputs after two Adam steps (what the reference methods saw)
    mean, qvec, svec = g["s0_mean"], g["s0_qvec"], torch.exp(g["s0_svec"])
    nn_pos, idx = K_nearest_neighbors(mean, K=4, knn=oracle_mod.knn_points)
    assert idx.shape == (mean.shape[0], 3) and torch.equal(idx, g["knn_idx_K4"]) and torch.equal(nn_pos, g["knn_nn_K4"])
    _, idx2, d2 = K_nearest_neighbors(mean, K=4, return_dist=True, knn=oracle_mod.knn_points)
    assert torch.equal(idx2, idx) and torch.allclose(d2, (nn_pos - mean[:, None]).pow(2).sum(-1), rtol=1e-5, atol=1e-7)
    p1, i1 = nearest_neighbor(mean, knn=oracle_mod.knn_points)
    assert torch.equal(i1, g["nn_idx"]) and torch.equal(p1, g["nn_pos"]) and torch.equal(i1, idx[:, 0])
    ours = distance_to_gaussian_surface(mean, svec, quat_to_rotmat(qvec), p1)
    assert torch.allclose(ours, g["surface_self_to_nn"], rtol=1e-6, atol=1e-8), float((ours - g["surface_self_to_nn"]).abs().max())


def _store_from(gold, tag, knn, **kw):
    st = GaussianStore({f: gold[f"{tag}_{f}"] for f in FIELDS}, C=None, device="cpu", knn_fn=knn, **kw)
    for f in FIELDS:
        n = st.N
        st._rows(st.exp_avg, f, n).copy_(gold[f"{tag}_{f}_exp_avg"].reshape(n, -1))
        st._rows(st.exp_avg_sq, f, n).copy_(gold[f"{tag}_{f}_exp_avg_sq"].reshape(n, -1))
    return st


def _check(st, gold, tag, n_old):
    """rows [0, n_old) are untouched (bit-exact, moments included); the appended rows match the reference's new
    Gaussians (same torch ops on both sides; tolerance only for op-fusion differences) and carry zero moments"""
    assert st.N == int(gold[f"{tag}_N"]), (st.N, int(gold[f"{tag}_N"]))
    for f in FIELDS:
        n = st.N
        for buf, suffix in ((st.flat_param, ""), (st.exp_avg, "_exp_avg"), (st.exp_avg_sq, "_exp_avg_sq")):
            ours, ref = st._rows(buf, f, n), gold[f"{tag}_{f}{suffix}"].reshape(n, -1)
            assert torch.equal(ours[:n_old], ref[:n_old]), (tag, f, suffix)
            if suffix:
                assert float(ours[n_old:].abs().max()) == 0.0 and float(ref[n_old:].abs().max()) == 0.0
            else:
                assert torch.allclose(ours[n_old:], ref[n_old:], rtol=1e-5, atol=1e-6), (
                    tag, f, float((ours[n_old:] - ref[n_old:]).abs().max()))
        assert st.params[f].shape[0] == n and st.params[f].requires_grad


def test_with_idx_matches_reference(oracle_mod):
    g = _load()
    st = _store_from(g, "s0", oracle_mod.knn_points)
    new = st.densify_by_compatness_with_idx(g["knn_idx_K4"][:, 0])
    k = g["with_idx0_mean"].shape[0]
    assert 0 < k < st.N  # the gap test rejects some pairs in this fixture
    for f in FIELDS:
        ours, ref = new[f].reshape(k, -1), g[f"with_idx0_{f}"].reshape(k, -1)
        assert torch.allclose(ours, ref, rtol=1e-5, atol=1e-6), (f, float((ours - ref).abs().max()))
    for f in ("qvec", "alpha", "color"):  # copied rows are exact
        assert torch.equal(new[f].reshape(k, -1), g[f"with_idx0_{f}"].reshape(k, -1))


@pytest.mark.parametrize("capacity", [None, 4096])
def test_densify_by_compatness_matches_reference(oracle_mod, capacity):
    g = _load()
    st = _store_from(g, "s0", oracle_mod.knn_points, capacity=capacity)
    n0 = st.N
    n_new = st.densify_by_compatness(K=3)
    assert n_new == int(g["s1_num"]) and st.N == n0 + n_new
    _check(st, g, "s1", n0)
    st2 = _store_from(g, "s0", oracle_mod.knn_points, capacity=capacity)
    n_new2 = st2.densify_by_shrink_then_compatness(1.5, K=2)
    assert n_new2 == int(g["s2_num"])
    # the shrink rewrites every old svec row: log(exp(raw) / 1.5), same two torch ops as the reference's setter
    assert torch.equal(st2._rows(st2.flat_param, "svec", n0), g["s2_svec"][:n0])
    _check(st2, g, "s2", n0)


@pytest.mark.parametrize("kind", ["compatness", "shrink_then_compatness"])
def test_dispatcher_matches_reference_trace(oracle_mod, kind):
    g = _load()
    st = _store_from(g, "s0", oracle_mod.knn_points)
    cfg = dict(enabled=True, type=kind, warm_up=100, end=1000, period=100, use_legacy=False, K=2, surface_shrink=1.25)
    trace = []
    for step in (0, 99, 100, 150, 200, 1100):
        st.mean_2d_grad_accum, st.cnt = torch.ones(st.N), torch.ones(st.N)
        n_before = st.N
        res = st.densify_step(step, cfg)
        trace.append([step, n_before, st.N])
        if res is not None:  # the accumulators were reset (:816-817)
            assert float(st.cnt.abs().max()) == 0.0 and st.cnt.shape[0] == st.N
    rows = g["dispatch_trace"].tolist()
    ref = rows[:6] if kind == "compatness" else rows[6:]
    assert trace == ref, (trace, ref)
    n = st.N
    for f in FIELDS:
        ours, refp = st._rows(st.flat_param, f, n), g[f"s3_{kind}_{f}"].reshape(n, -1)
        assert torch.allclose(ours, refp, rtol=2e-5, atol=2e-6), (f, float((ours - refp).abs().max()))


def test_unknown_densify_type_raises(oracle_mod):
    g = _load()
    st = _store_from(g, "s0", oracle_mod.knn_points)
    with pytest.raises(NotImplementedError):
        st.densify_step(100, dict(enabled=True, type="nope", warm_up=0, end=1000, period=100))


def test_knn_has_no_cpu_fallback():
    from gsgen_b200.knn import knn_points

    with pytest.raises(RuntimeError, match="CUDA"):
        knn_points(None, torch.zeros(8, 3), 2)
    st = GaussianStore({"mean": torch.randn(8, 3), "qvec": torch.randn(8, 4), "svec": torch.zeros(8, 3),
                        "alpha": torch.zeros(8), "color": torch.zeros(8, 3)}, C=None, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        st.densify_by_compatness(K=1)


# ---- penalties on the trainer surface -------------------------------------------------------------------------------
def _renderer(g, oracle_mod, penalty):
    from gsgen_b200.splatting import GaussianSplattingRenderer

    init = {"mean": g["s0_mean"], "qvec": g["s0_qvec"], "svec": g["s0_svec"], "color": g["s0_color"],
            "alpha": g["s0_alpha"], "raw": True}
    return GaussianSplattingRenderer({"penalty": penalty}, init, device="cpu", render_fn=lambda *a, **k: None,
                                     knn_fn=oracle_mod.knn_points)


class _Writer:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, step):
        self.scalars[tag] = float(value)


@pytest.mark.parametrize("kind", ["l1", "l2"])
def test_compat_penalty_matches_reference(oracle_mod, kind):
    g = _load()
    r = _renderer(g, oracle_mod, {"compat": {"value": 0.7, "type": kind}})
    r.store.zero_grad()
    w = _Writer()
    loss = r.compat_penalty_loss(10, w)
    loss.backward()
    assert torch.allclose(loss.detach().reshape(1), g[f"compat_{kind}_loss"], rtol=1e-5, atol=1e-8)
    assert abs(w.scalars["auxiliary/effective_rate"] - float(g[f"compat_{kind}_effective_rate"])) < 1e-6
    for f, key in (("mean", "g_mean"), ("svec", "g_svec"), ("qvec", "g_qvec")):
        ours, ref = r.store.grad_views[f], g[f"compat_{kind}_{key}"]
        assert torch.allclose(ours, ref, rtol=1e-4, atol=1e-7), (f, float((ours - ref).abs().max()))
    # total through auxiliary_loss
    r.store.zero_grad()
    total = r.auxiliary_loss(10, w)
    assert torch.allclose(total.detach(), loss.detach()) and "auxiliary/total" in w.scalars


def test_trainer_level_compatness_calls(oracle_mod):
    """trainer.py:799-801 calls renderer.densify_by_compatness(3) and renderer.reset_densify_info() directly"""
    g = _load()
    r = _renderer(g, oracle_mod, {})
    n0 = r.N
    r.store.cnt += 1.0
    n_new = r.densify_by_compatness(3)
    assert n_new == int(g["s1_num"]) and r.N == n0 + n_new and r.mean.shape[0] == r.N
    r.reset_densify_info()
    assert r.store.cnt.shape[0] == r.N and float(r.store.cnt.abs().max()) == 0.0


def test_nn_penalty_matches_reference(oracle_mod):
    g = _load()
    r = _renderer(g, oracle_mod, {"NN": {"value": 0.3}})
    r.store.zero_grad()
    loss = r.NN_penalty_loss(10, _Writer())
    loss.backward()
    assert torch.allclose(loss.detach().reshape(1), g["NN_loss"], rtol=1e-5, atol=1e-8)
    assert torch.allclose(r.store.grad_views["mean"], g["NN_g_mean"], rtol=1e-4, atol=1e-8)
    zero = _renderer(g, oracle_mod, {"NN": {"value": 0.0}}).NN_penalty_loss(10, None)
    assert float(zero) == 0.0 and not zero.requires_grad
