"""Replica consistency of the view-parallel mode on CPU: two gloo ranks, each rendering its
shard of every step's view batch through `GaussianSplattingRenderer` (CPU oracle plugged in as the renderer), then
densify(step) + prune(step).  Without the extra state the ranks select different Gaussians at the first densify and
diverge silently (N may even stay equal): the per-view side effects `mean_2d_grad_accum += |g_mean2d|`, `cnt += 1`
(gs/gaussian_splatting.py:464-469), `max_radii2d = max(..)` (:1240-1245) and the split noise `torch.randn` (:576-579)
are all rank-local.  Required: bit-identical arenas on both ranks, and the same Gaussians as ONE process rendering
all the views."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.test_parallel_cpu import _free_port
from tests.test_splatting_cpu import GOLD, _oracle_render_fn

CFG = {"densify": dict(enabled=True, type="official", warm_up=2, end=100, period=2, mean2d_thresh=2e-5,
                       split_thresh=0.02, n_splits=2, split_shrink=0.8, use_legacy=False),
       "prune": dict(enabled=True, warm_up=0, end=100, period=2, radii2d_thresh=0.0, alpha_thresh=0.05,
                     radii3d_thresh=0.0)}
# the rule the top-level experiment configs select: legacy split / clone (its own noise draw + the optimizer reset), then
# a compactness pass over the neighbour search -- every rank must end with the same arena here too
CFG_LEGACY = {"densify": dict(enabled=True, type="shrink_then_compatness", warm_up=2, end=100, period=2,
                              mean2d_thresh=2e-5, split_thresh=0.02, n_splits=2, split_shrink=0.8, use_legacy=True, K=2,
                              surface_shrink=1.5),
              "prune": CFG["prune"]}
CFGS = {"official": CFG, "legacy_then_compatness": CFG_LEGACY}
N_VIEWS, N_STEPS = 4, 3


def _setup(group, kind="official"):
    import oracle
    from gsgen_b200.camera import CameraInfo, orbit_c2w
    from gsgen_b200.splatting import GaussianSplattingRenderer

    z = np.load(GOLD)
    gold = {k: torch.from_numpy(z[k]) for k in z.files}
    init = {k: gold[f"a_in_{k}"] for k in ("mean", "qvec", "svec", "color", "alpha")}
    r = GaussianSplattingRenderer(CFGS[kind], init, device="cpu", background=None, render_fn=_oracle_render_fn(oracle),
                                  group=group, capacity=None, knn_fn=oracle.knn_points)
    fx, fy, cx, cy, w, h, near, far = gold["a_cam"].tolist()
    cam = CameraInfo(fx, fy, cx, cy, int(w), int(h), near, far)
    base = gold["a_c2w"]
    views = []
    for v in range(N_VIEWS):  # the fixture's pose, rotated about the scene's up axis
        a = 0.35 * v
        R = torch.tensor([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32)
        views.append(torch.cat([R @ base[:3, :3], (R @ base[:3, 3]).unsqueeze(-1)], dim=-1).contiguous())
    return r, cam, views


def _run(r, cam, views, mine, sizes):
    for step in range(N_STEPS):
        r.update(step)
        out = r({"c2w": torch.stack([views[v] for v in mine]), "camera_info": [cam] * len(mine)}, use_bg=False)
        g = torch.Generator().manual_seed(1000 * step + 7)
        w = torch.randn(N_VIEWS, cam.h, cam.w, 3, generator=g)
        (out["rgb"] * w[mine]).sum().backward()
        r.store.all_reduce()  # gradients: one SUM per step (not used further here: the optimizer step needs the GPU)
        r.post_backward()
        n0 = r.N
        r.densify(step)
        r.prune(step)
        sizes.append((n0, r.N))
        r.store.zero_grad()


def _worker(rank, world, port, ret, kind):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        torch.manual_seed(1234 + rank)  # rank-local RNG streams DIFFER: the split noise must not come from them
        r, cam, views = _setup(None, kind)
        sizes = []
        _run(r, cam, views, [v for v in range(N_VIEWS) if v % world == rank], sizes)
        st = r.store
        flat = st.flat_param[: st.layout[-1][2] + st.layout[-1][3]].clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], g) for g in gathered), "replicas diverged"
        stats = torch.stack((st.mean_2d_grad_accum, st.cnt, st.max_radii2d))
        gs = [torch.empty_like(stats) for _ in range(world)]
        dist.all_gather(gs, stats)
        assert all(torch.equal(gs[0], g) for g in gs), "densification statistics differ between ranks"
        if rank == 0:
            ret["sizes"] = sizes
            ret["params"] = {k: v.detach().clone() for k, v in st.params.items()}
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("kind", ["official", "legacy_then_compatness"])
def test_two_ranks_stay_identical_through_densify_and_prune(oracle_mod, kind):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret, kind), nprocs=world, join=True)
    sizes = ret["sizes"]
    assert sizes[2][1] != sizes[2][0], "the densify / prune step did not change N: the test exercises nothing"
    # one process, all views: same Gaussians selected (the split children differ only through the noise draw)
    # (official: the selection does not depend on the noise; legacy + compactness: the compactness pass runs on the
    # noisy split children, so the single process must draw what rank 0 drew)
    torch.manual_seed(99 if kind == "official" else 1234)
    r, cam, views = _setup(None, kind)
    single = []
    _run(r, cam, views, list(range(N_VIEWS)), single)
    assert single == sizes, (single, sizes)
    for k in ("qvec", "alpha", "color", "svec"):  # fields the split noise does not touch
        assert torch.allclose(ret["params"][k], r.store.params[k].detach(), rtol=1e-5, atol=1e-6), k
