"""Property test of the grid search (gsgen_b200/csrc/knn_grid.cuh, host build): for ANY cloud, query set, K and cell
budget the result equals the brute-force search bit for bit.  `hypothesis` draws small clouds built from the shapes that
stress a uniform grid -- points snapped to a coarse lattice (exact ties, duplicates), clusters at very different scales,
flat and collinear sets, far outliers (the 3-sigma clip puts them in border cells), huge and tiny coordinates -- and
queries inside, on and far outside the box."""
import ctypes

import numpy as np
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st


def _grid_knn(hostmath, pts, K, queries, max_cells):
    pts = np.ascontiguousarray(pts, np.float32)
    n = pts.shape[0]
    q = None if queries is None else np.ascontiguousarray(queries, np.float32)
    nq = n if q is None else q.shape[0]
    idx = np.empty((nq, K), np.int64)
    d2 = np.empty((nq, K), np.float32)
    vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    rc = hostmath.hm_knn(n, vp(pts), nq, vp(q), K, ctypes.c_uint(max_cells), vp(idx), vp(d2), None)
    assert rc == 0
    return torch.from_numpy(d2), torch.from_numpy(idx)


@st.composite
def clouds(draw):
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    n = draw(st.integers(1, 160))
    kind = draw(st.sampled_from(["ball", "lattice", "clusters", "plane", "line", "outliers", "scaled", "mixed"]))
    scale = 10.0 ** draw(st.integers(-4, 4))
    offset = 10.0 ** draw(st.integers(-2, 3)) * draw(st.sampled_from([0.0, 1.0, -1.0]))
    if kind == "ball":
        p = rng.normal(size=(n, 3))
    elif kind == "lattice":
        p = rng.integers(0, draw(st.integers(1, 4)) + 1, size=(n, 3)).astype(np.float64)
    elif kind == "clusters":
        c = rng.normal(size=(3, 3)) * 10
        p = c[rng.integers(0, 3, size=n)] + rng.normal(size=(n, 3)) * 10.0 ** rng.integers(-4, 0, size=(n, 1))
    elif kind == "plane":
        p = rng.normal(size=(n, 3))
        p[:, draw(st.integers(0, 2))] = 0.5
    elif kind == "line":
        p = np.outer(rng.uniform(-1, 1, size=n), rng.normal(size=3))
    elif kind == "outliers":
        p = rng.normal(size=(n, 3)) * 0.01
        p[: max(1, n // 20)] *= 1e4
    elif kind == "scaled":
        p = rng.normal(size=(n, 3)) * np.array([1.0, 1e-3, 1e3])
    else:
        p = rng.normal(size=(n, 3))
        p[n // 2:] = p[: n - n // 2]  # exact duplicates
    p = (p * scale + offset).astype(np.float32)
    K = draw(st.integers(1, 32))
    qkind = draw(st.sampled_from(["self", "inside", "outside", "points"]))
    if qkind == "self":
        q = None
    elif qkind == "inside":
        q = (rng.normal(size=(20, 3)) * scale + offset).astype(np.float32)
    elif qkind == "outside":
        q = (rng.normal(size=(20, 3)) * scale * 50 + offset).astype(np.float32)
    else:
        q = p[rng.integers(0, n, size=15)].copy()
    max_cells = draw(st.sampled_from([1, 7, 64, 4 * n + 64, 100000]))
    return p, q, K, max_cells


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(case=clouds())
def test_grid_search_equals_brute_force_on_any_cloud(hostmath, oracle_mod, case):
    p, q, K, max_cells = case
    d2, idx = _grid_knn(hostmath, p, K, q, max_cells)
    d2_ref, idx_ref = oracle_mod.knn_points(None if q is None else torch.from_numpy(q), torch.from_numpy(p), K)
    assert torch.equal(idx, idx_ref), (p.shape, K, max_cells)
    assert torch.equal(d2, d2_ref)
