"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes, each rendering its shard of the
views with the CPU oracle standing in for the CUDA kernels; the single flat-buffer all-reduce must reproduce the
single-process gradient of the whole view batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gsgen_b200.parallel import ViewParallelRenderer, field_layout, shard_views


def test_shard_views_round_robin():
    assert shard_views(8, 0, 8) == [0] and shard_views(8, 7, 8) == [7]
    assert shard_views(8, 1, 4) == [1, 5]
    assert sorted(sum((shard_views(8, r, 3) for r in range(3)), [])) == list(range(8))
    assert shard_views(2, 3, 4) == []


def test_field_layout_sizes():
    from gsgen_b200.parallel import layout_total

    lay = field_layout(10, 4)
    assert [l[0] for l in lay] == ["mean", "qvec", "svec", "alpha", "sh"]
    assert sum(l[3] for l in lay) == 10 * (3 + 4 + 3 + 1 + 48)  # 59 floats = 236 B / Gaussian at SH degree 3
    assert field_layout(10, None)[-1][0] == "color"
    assert sum(l[3] for l in field_layout(10, None)) == 10 * 14  # 56 B / Gaussian RGB
    # every field starts on a 16-byte boundary whatever N is (float4 accesses to qvec / g_qvec; round-1 advice)
    for N in (1, 2, 3, 5, 10, 4097, 6145):
        for C in (None, 1, 3, 4):
            lay = field_layout(N, C)
            assert all(off % 4 == 0 for _, _, off, _ in lay), (N, C)
            assert all(lay[i][2] + lay[i][3] <= lay[i + 1][2] for i in range(len(lay) - 1))
            assert layout_total(lay) % 4 == 0 and layout_total(lay) - (lay[-1][2] + lay[-1][3]) < 4
    # with N % 4 == 0 there is no padding: the buffer is exactly the 236 B / Gaussian the all-reduce is quoted at
    assert layout_total(field_layout(1000, 4)) == 1000 * 59


def _scene_and_views():
    from gsgen_b200.camera import orbit_c2w
    from gsgen_b200.scenes import make_scene

    sc = make_scene("c1", N=600, reso=48)
    views = [orbit_c2w(2.5, 15.0, 30.0 + 90.0 * v) for v in range(4)]
    return sc, views


def _render_and_backward_factory(sc, views):
    import oracle
    from tests.util import ocam_of

    cam = sc.cams[0]

    def f(params, v):
        g = torch.Generator().manual_seed(100 + v)
        w = torch.randn(cam.h, cam.w, 3, generator=g)
        out = oracle.render_view(params["mean"], params["qvec"], params["svec"], params["alpha"], views[v],
                                 ocam_of(cam), sh=params["sh"], C=1)
        out["rgb"].backward(gradient=w)

    return f


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        sc, views = _scene_and_views()
        vpr = ViewParallelRenderer(dict(mean=sc.mean, qvec=sc.qvec, svec=sc.svec, alpha=sc.alpha, sh=sc.sh), 1, "cpu")
        mine = vpr.step(len(views), _render_and_backward_factory(sc, views))
        assert mine == shard_views(len(views), rank, world)
        # replicas identical after the collective
        gathered = [torch.empty_like(vpr.flat_grad) for _ in range(world)]
        dist.all_gather(gathered, vpr.flat_grad)
        assert all(torch.equal(gathered[0], g) for g in gathered)
        if rank == 0:
            ret["flat"] = vpr.flat_grad.clone()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_view_sharded_allreduce_matches_single_process(oracle_mod):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    sc, views = _scene_and_views()
    single = ViewParallelRenderer(dict(mean=sc.mean, qvec=sc.qvec, svec=sc.svec, alpha=sc.alpha, sh=sc.sh), 1, "cpu")
    single.step(len(views), _render_and_backward_factory(sc, views))
    a, b = ret["flat"], single.flat_grad
    assert float(b.abs().max()) > 0
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * float(b.abs().max()))
    # .grad of every parameter is a view of the flat buffer (no pack pass before the collective)
    for name, p in single.params.items():
        assert p.grad.data_ptr() == single.grad_views[name].data_ptr()


def _store_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gsgen_b200.store import GaussianStore

        g = torch.Generator().manual_seed(5)
        N = 40
        params = dict(mean=torch.randn(N, 3, generator=g), qvec=torch.randn(N, 4, generator=g),
                      svec=torch.randn(N, 3, generator=g), alpha=torch.randn(N, generator=g),
                      color=torch.randn(N, 3, generator=g))
        for cap in (None, 64):  # full arena: one flat collective; spare capacity: live rows only
            st = GaussianStore(params, None, "cpu", capacity=cap)
            st.zero_grad()
            loss = sum(((rank + 1.0) * (i + 1) * p).sum() for i, p in enumerate(st.params.values()))
            loss.backward()
            st.all_reduce()
            for i, (name, gv) in enumerate(st.grad_views.items()):
                assert torch.equal(gv, torch.full_like(gv, 3.0 * (i + 1))), (cap, name)  # (1 + 2) * (i + 1)
            if cap is not None:  # dead rows were neither sent nor touched
                for name in st.grad_views:
                    assert float(st._rows(st.flat_grad, name, st.cap - st.N, st.N).abs().max()) == 0.0
        if rank == 0:
            ret["ok"] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_store_allreduce_sums_live_rows_only():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_store_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret.get("ok")
