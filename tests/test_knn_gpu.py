"""-m gpu: gsb200_knn (csrc/knn.cu, uniform-grid shell search) through the C ABI against brute force -- indices AND
squared distances bit for bit (the kernel's distance is the sum of three separately rounded squares, which is what
elementwise torch ops produce; ties are ordered by index on both sides) -- and compactness-based densification /
the neighbour penalties on a CUDA store against the CPU twin the reference-generated fixture pins
(tests/test_compatness_cpu.py)."""
import numpy as np
import pytest
import torch

from gsgen_b200.knn import K_nearest_neighbors, knn_points, nearest_neighbor
from gsgen_b200.store import GaussianStore
from tests.util import note

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _brute(points, K, queries=None, chunk=None):
    """exact K nearest by ((dx*dx + dy*dy) + dz*dz) in fp32 with separate elementwise ops, ties by smaller index
    (stable sort), on the GPU.  Returns (dist2 [Q,K], idx [Q,K])."""
    p = points.to(DEV, torch.float32)
    q = p if queries is None else queries.to(DEV, torch.float32)
    n = p.shape[0]
    chunk = chunk or max(32, min(1024, (1 << 24) // max(n, 1)))
    out_d = torch.full((q.shape[0], K), float("inf"), device=DEV)
    out_i = torch.full((q.shape[0], K), -1, dtype=torch.int64, device=DEV)
    px, py, pz = p[:, 0].contiguous(), p[:, 1].contiguous(), p[:, 2].contiguous()
    for s in range(0, q.shape[0], chunk):
        qq = q[s:s + chunk]
        dx, dy, dz = px[None] - qq[:, 0:1], py[None] - qq[:, 1:2], pz[None] - qq[:, 2:3]
        d = (dx * dx + dy * dy) + dz * dz
        ds, di = torch.sort(d, dim=1, stable=True)
        k = min(K, n)
        out_d[s:s + chunk, :k], out_i[s:s + chunk, :k] = ds[:, :k], di[:, :k]
    return out_d, out_i


def _clouds():
    rng = np.random.default_rng(0)
    out = {"ball": rng.normal(size=(20000, 3)) * 0.5, "uniform": rng.uniform(-1, 1, size=(30011, 3))}
    c = rng.normal(size=(6000, 3)) * 0.01
    c[:3000] += 5.0
    out["two_clusters"] = c
    p = rng.uniform(-1, 1, size=(5000, 3))
    p[:, 2] = 0.25
    out["plane"] = p
    p = rng.uniform(-1, 1, size=(4000, 3))
    p[2000:] = p[:2000]
    out["duplicates"] = p
    out["coincident"] = np.ones((300, 3))
    out["three"] = rng.normal(size=(3, 3))
    out["one"] = rng.normal(size=(1, 3))
    o = rng.normal(size=(8000, 3)) * 0.1
    o[0] = [100, 100, 100]
    out["outlier"] = o
    out["lattice"] = np.stack(np.meshgrid(*[np.arange(16.0)] * 3, indexing="ij"), -1).reshape(-1, 3)
    return {k: torch.from_numpy(v.astype(np.float32)) for k, v in out.items()}


@pytest.mark.parametrize("name", sorted(_clouds()))
def test_knn_equals_brute_force(name):
    pts = _clouds()[name].to(DEV)
    for K in (1, 2, 4, 7, 16, 32):
        d2, idx = knn_points(None, pts, K)
        d2_ref, idx_ref = _brute(pts, K)
        assert torch.equal(idx, idx_ref), (name, K, int((idx != idx_ref).sum()))
        assert torch.equal(d2, d2_ref), (name, K)


def test_knn_external_queries_and_contract():
    g = torch.Generator().manual_seed(1)
    pts = (torch.randn(15000, 3, generator=g) * 0.5).to(DEV)
    q = (torch.randn(3000, 3, generator=g) * 2.0).to(DEV)  # most of them outside the grid box
    for K in (1, 3, 8):
        d2, idx = knn_points(q, pts, K)
        d2_ref, idx_ref = _brute(pts, K, q)
        assert torch.equal(idx, idx_ref) and torch.equal(d2, d2_ref)
    # no distances requested
    none, idx2 = knn_points(q, pts, 8, return_dist=False)
    assert none is None and torch.equal(idx2, idx)
    # fewer points than K
    d2, idx = knn_points(None, pts[:3], 5)
    assert idx[:, 3:].eq(-1).all() and torch.isinf(d2[:, 3:]).all() and idx[:, :3].ge(0).all()
    with pytest.raises(RuntimeError):
        knn_points(None, pts, 33)
    with pytest.raises(RuntimeError):
        knn_points(None, pts.cpu(), 2)
    # the reference-facing wrappers (utils/ops.py:103-134): column 0 (the point itself) dropped
    nn_pos, nidx = K_nearest_neighbors(pts, K=4)
    d2_ref, idx_ref = _brute(pts, 4)
    assert torch.equal(nidx, idx_ref[:, 1:]) and torch.equal(nn_pos, pts[idx_ref[:, 1:]])
    p1, i1 = nearest_neighbor(pts)
    assert torch.equal(i1, idx_ref[:, 1]) and torch.equal(p1, pts[idx_ref[:, 1]])


def test_knn_one_million_points_sampled_check_and_time():
    """C3-sized cloud: every query's row is ascending and starts with itself; 512 sampled rows against brute force;
    the time is printed (pytorch3d's knn_points is O(N^2): 1e12 distance evaluations at this size)."""
    g = torch.Generator().manual_seed(2)
    n = 1_000_000
    v = torch.randn(n, 3, generator=g)
    pts = (v / v.norm(dim=1, keepdim=True) * torch.rand(n, 1, generator=g) ** (1 / 3)).to(DEV)  # uniform in the unit ball
    knn_points(None, pts, 4)  # warm-up (scratch allocation)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d2, idx = knn_points(None, pts, 4)
    e1.record()
    torch.cuda.synchronize()
    note(f"gsb200_knn: {n} points (uniform ball), K=4, self query: {e0.elapsed_time(e1):.3f} ms")
    assert torch.equal(idx[:, 0], torch.arange(n, device=DEV)) and float(d2[:, 0].max()) == 0.0
    assert bool((d2[:, 1:] >= d2[:, :-1]).all()) and int(idx.min()) >= 0 and int(idx.max()) < n
    sel = torch.randint(0, n, (512,), generator=g).to(DEV)
    d2_ref, idx_ref = _brute(pts, 4, pts[sel], chunk=32)
    assert torch.equal(idx[sel], idx_ref) and torch.equal(d2[sel], d2_ref)


def _twin(N, seed=0):
    import oracle

    g = torch.Generator().manual_seed(seed)
    lattice = torch.stack(torch.meshgrid(*[torch.arange(16.0)] * 3, indexing="ij"), -1).reshape(-1, 3)[:N] * 0.25
    p = dict(mean=lattice + 0.05 * torch.randn(N, 3, generator=g),
             qvec=torch.nn.functional.normalize(torch.randn(N, 4, generator=g), dim=-1),
             svec=torch.log(0.02 + 0.22 * torch.rand(N, 3, generator=g)), alpha=torch.randn(N, generator=g),
             color=torch.randn(N, 3, generator=g))
    return GaussianStore(p, None, "cpu", knn_fn=oracle.knn_points), GaussianStore(p, None, DEV)


def test_compatness_densify_on_the_device_matches_the_cpu_twin():
    """same selection (the neighbour indices are identical) and the same new Gaussians up to the rounding of the
    elementwise torch ops on the two devices; old rows and Adam moments bit for bit"""
    a, b = _twin(3000)
    for st in (a, b):
        for buf in (st.exp_avg, st.exp_avg_sq):
            for name in st._field:
                st._rows(buf, name, st.N).copy_(3.0 * st._rows(st.flat_param, name, st.N) + 0.5)
    n0 = a.N
    na, nb = a.densify_by_compatness(K=3), b.densify_by_compatness(K=3)
    assert 0 < na < 3 * n0 and a.N == n0 + na and b.N == n0 + nb
    # (the gap test `surface + nn_surface < dist` is evaluated with each device's sqrt / normalize: a pair sitting on
    # the threshold to the last ulp may legitimately fall on either side -- a handful in 9 000 at most)
    assert abs(na - nb) <= 3, (na, nb)
    for x, y, what in zip(a._buffers(), b._buffers(), ("param", "grad", "exp_avg", "exp_avg_sq")):
        for name in a._field:
            ra, rb = a._rows(x, name, a.N), b._rows(y, name, b.N).cpu()
            assert torch.equal(ra[:n0], rb[:n0]), (what, name)
            if what != "param":
                assert float(rb[n0:].abs().max()) == 0.0, (what, name)
            elif na == nb:
                if name == "svec":  # log(gap / 6): compare the scale itself (a gap near zero amplifies in the log)
                    ra, rb = torch.exp(ra), torch.exp(rb)
                assert torch.allclose(ra[n0:], rb[n0:], rtol=1e-4, atol=1e-6), (name, float((ra[n0:] - rb[n0:]).abs().max()))
    # the dispatcher (shrink, then K neighbours) on the grown device store.  (Not compared with the CPU twin: mutual
    # neighbours put their new Gaussians at the SAME point of the gap, so round two is full of near-coincident pairs
    # whose neighbour order is decided by the last ulp of either device.)
    cfg = dict(enabled=True, type="shrink_then_compatness", warm_up=0, end=1000, period=100, K=2, surface_shrink=1.25)
    n1 = b.N
    res = b.densify_step(100, cfg)
    assert res is not None and 0 < res[0] <= 2 * n1 and b.N == n1 + res[0]
    assert float(b.cnt.abs().max()) == 0.0 and b.cnt.shape[0] == b.N
    for y in b._buffers()[1:]:
        for name in b._field:
            assert float(b._rows(y, name, b.N - n1, n1).abs().max()) == 0.0  # gradient / moment rows of the children
    assert bool(torch.isfinite(b._rows(b.flat_param, "mean", b.N)).all())


def test_neighbour_penalties_on_the_device_match_the_cpu_twin():
    import oracle
    from gsgen_b200.splatting import GaussianSplattingRenderer

    g = torch.Generator().manual_seed(5)
    N = 2500
    init = {"mean": torch.randn(N, 3, generator=g) * 0.5, "qvec": torch.randn(N, 4, generator=g),
            "svec": torch.log(0.01 + 0.05 * torch.rand(N, 3, generator=g)), "color": torch.randn(N, 3, generator=g),
            "alpha": torch.randn(N, generator=g), "raw": True}
    pen = {"compat": {"value": 0.7, "type": "l2"}, "NN": {"value": 0.3}}
    ra = GaussianSplattingRenderer({"penalty": pen}, init, device="cpu", render_fn=lambda *a, **k: None,
                                   knn_fn=oracle.knn_points)
    rb = GaussianSplattingRenderer({"penalty": pen}, init, device=DEV)
    out = []
    for r in (ra, rb):
        r.store.zero_grad()
        loss = r.auxiliary_loss(10, None)
        loss.backward()
        out.append((loss.detach().cpu(), {f: r.store.grad_views[f].detach().cpu().clone() for f in ("mean", "svec", "qvec")}))
    assert torch.allclose(out[0][0], out[1][0], rtol=1e-5, atol=1e-8), (out[0][0], out[1][0])
    for f in ("mean", "svec", "qvec"):
        assert torch.allclose(out[0][1][f], out[1][1][f], rtol=1e-3, atol=1e-7), (f, float((out[0][1][f] - out[1][1][f]).abs().max()))
