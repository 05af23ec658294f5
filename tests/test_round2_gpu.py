"""-m gpu tests of the round-2 additions to the fused path:
  * one library context per in-flight view + the generation stamp (round-1 advice: a batch rendered before ONE
    loss.backward() used to composite every view's backward against the LAST view's binning);
  * parameter counts that are not multiples of 4 (16-byte field alignment of the flat layouts);
  * asynchronous-count mode (no host wait; capacity-sized tile sort; overflow -> loud error -> re-render);
  * both SH backward kernels (direct vector-reduction flush / round-1 shared accumulator) against each other;
  * gsb200_view_stats;
  * 2-GPU replica consistency through densify + prune (skipped on a 1-GPU box)."""
import os
import socket

import pytest
import torch

from gsgen_b200.scenes import make_scene
from tests.util import assert_grad_close, ocam_of

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _small_scene(N=5000, reso=160, scale=3.0, cfg="c3"):
    sc = make_scene(cfg, N=N, reso=reso)
    sc.svec = (sc.svec * scale).contiguous()
    return sc


def _raw_init(sc):
    return {"mean": sc.mean, "qvec": sc.qvec, "svec": sc.svec, "color": sc.color.clamp(1e-3, 1 - 1e-3),
            "alpha": sc.alpha.clamp(1e-3, 1 - 1e-3)}


def test_batch_of_views_uses_one_context_per_view():
    """`renderer(batch)` then ONE backward (the reference trainer's flow, trainer.py:575-599) == the sum of per-view
    render + backward."""
    from gsgen_b200.camera import orbit_c2w
    from gsgen_b200.splatting import GaussianSplattingRenderer

    sc = _small_scene(N=4001)  # 4001: not a multiple of 4 either
    cam = sc.cams[0]
    views = [orbit_c2w(2.5, 15.0, 30.0 + 70.0 * v) for v in range(3)]
    g = torch.Generator().manual_seed(3)
    ws = [torch.randn(cam.h, cam.w, 3, generator=g).to(DEV) for _ in views]

    def fresh():
        r = GaussianSplattingRenderer({}, _raw_init(sc), device=DEV)
        r.store.zero_grad()
        return r

    r = fresh()
    out = r({"c2w": torch.stack(views), "camera_info": [cam] * len(views)}, use_bg=False, rgb_only=False)
    sum((out["rgb"][i] * ws[i]).sum() + out["depth"][i].sum() for i in range(len(views))).backward()
    batch_grad = r.store.flat_grad.clone()
    assert len({a["slot"] for a in r._pending}) == len(views)
    r.post_backward()

    r2 = fresh()
    for i, v in enumerate(views):
        o = r2.render_one(v, cam, use_bg=False, rgb_only=False)
        ((o["rgb"] * ws[i]).sum() + o["depth"].sum()).backward()
        r2.post_backward()
    assert_grad_close(batch_grad, r2.store.flat_grad, rtol=1e-5, what="batch vs per-view gradients")
    assert float(r2.store.flat_grad.abs().max()) > 0


def test_backward_fails_loudly_when_its_context_was_overwritten():
    from gsgen_b200.rasterizer import render_view

    sc = _small_scene(N=3000).to(DEV)
    cam, c2w = sc.cams[0], sc.c2ws[0].cpu()
    leaf = lambda t: t.clone().requires_grad_()
    a = render_view(leaf(sc.mean), leaf(sc.qvec), leaf(sc.svec), leaf(sc.alpha), c2w, cam, sh=leaf(sc.sh), C=4, slot=5)
    with torch.no_grad():  # a later forward on the same slot (e.g. an evaluation render) overwrites the view's state
        render_view(sc.mean, sc.qvec, sc.svec, sc.alpha, c2w, cam, sh=sc.sh, C=4, slot=5)
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        a["rgb"].sum().backward()


@pytest.mark.parametrize("N", [4097, 6145, 3])
def test_parameter_counts_that_are_not_multiples_of_four(oracle_mod, N):
    """ViewParallelRenderer / GaussianStore layouts with N % 4 != 0 (6145 = int(4096*1.5)+1, the capacity the round-1
    store grew to): the kernels read qvec / write g_qvec as float4."""
    from gsgen_b200.parallel import ViewParallelRenderer
    from gsgen_b200.rasterizer import render_view

    sc = _small_scene(N=N, reso=96)
    cam, c2w = sc.cams[0], sc.c2ws[0]
    vpr = ViewParallelRenderer(dict(mean=sc.mean, qvec=sc.qvec, svec=sc.svec, alpha=sc.alpha, sh=sc.sh), 4, DEV)
    assert all(p.data_ptr() % 16 == 0 for p in vpr.params.values())
    vpr.zero_grad()
    g = torch.Generator().manual_seed(1)
    w = torch.randn(cam.h, cam.w, 3, generator=g)
    out = render_view(vpr.params["mean"], vpr.params["qvec"], vpr.params["svec"], vpr.params["alpha"], c2w, cam,
                      sh=vpr.params["sh"], C=4, grad_sink=vpr.grad_views)
    out["rgb"].backward(gradient=w.to(DEV))
    torch.cuda.synchronize()
    leaf = lambda t: t.clone().requires_grad_()
    mo, qo, so, ao, sho = leaf(sc.mean), leaf(sc.qvec), leaf(sc.svec), leaf(sc.alpha), leaf(sc.sh)
    ref = oracle_mod.render_view(mo, qo, so, ao, c2w, ocam_of(cam), sh=sho, C=4)
    ref["rgb"].backward(gradient=w)
    if float(sho.grad.abs().max()) > 0:
        for name, r_ in (("mean", mo), ("qvec", qo), ("svec", so), ("alpha", ao), ("sh", sho)):
            assert_grad_close(vpr.grad_views[name], r_.grad, rtol=3e-3, what=f"N={N} g_{name}")


@pytest.mark.parametrize("cfg", ["rgb", "sh"])
def test_async_count_mode_matches_the_synchronous_path(cfg):
    from gsgen_b200 import _lib
    from gsgen_b200.rasterizer import render_view, view_stats

    sc = _small_scene(N=20000, reso=256, scale=2.0).to(DEV)
    cam, c2w = sc.cams[0], sc.c2ws[0].cpu()
    g = torch.Generator().manual_seed(2)
    w = torch.randn(cam.h, cam.w, 3, generator=g).to(DEV)
    kw = dict(sh=sc.sh, C=4) if cfg == "sh" else dict(color=sc.color)

    def run(slot, async_count, scale=1.0):
        leaf = lambda t: t.clone().requires_grad_()
        m, q, s, a = leaf(sc.mean), leaf(sc.qvec), leaf((sc.svec * scale).contiguous()), leaf(sc.alpha)
        k2 = {k: (leaf(v) if torch.is_tensor(v) else v) for k, v in kw.items()}
        out = render_view(m, q, s, a, c2w, cam, slot=slot, async_count=async_count, **k2)
        out["rgb"].backward(gradient=w)
        pay = k2["sh"] if cfg == "sh" else k2["color"]
        return out, (m.grad, q.grad, s.grad, a.grad, pay.grad)

    s_sync, s_async, s_big = (20, 21, 22) if cfg == "rgb" else (23, 24, 25)  # fresh contexts per case
    ref_out, ref_g = run(s_sync, False)
    first, _ = run(s_async, True)  # first view of a context: synchronous (no capacity known yet)
    assert first["aux"]["N_with_dub"] == ref_out["aux"]["N_with_dub"]
    out, gr = run(s_async, True)   # second view: no host wait, capacity-sized sort with padding keys
    assert out["aux"]["N_with_dub"] is None
    assert torch.equal(out["rgb"], ref_out["rgb"])  # same lists, same order -> same image bit for bit
    for a_, b_ in zip(gr, ref_g):
        assert_grad_close(a_, b_, rtol=1e-5, what="async vs sync gradient")
    assert view_stats(DEV, s_async)[0] == ref_out["aux"]["N_with_dub"]
    # a view that needs more tile-list entries than the capacity learnt so far: rejected loudly, then fits
    big_ref, _ = run(s_big, False, scale=2.5)
    assert big_ref["aux"]["N_with_dub"] > 1.3 * ref_out["aux"]["N_with_dub"]
    with pytest.raises(_lib.TileListOverflow):
        run(s_async, True, scale=2.5)   # its own backward reports it if the count has arrived by then ...
        view_stats(DEV, s_async)        # ... else the blocking check does
    big, _ = run(s_async, True, scale=2.5)  # the capacity was raised by the rejected view
    assert torch.equal(big["rgb"], big_ref["rgb"])


@pytest.mark.parametrize("C", [2, 3, 4])
def test_sh_backward_kernels_agree(C):
    """direct vector-reduction flush (composite_bwd_sh.cu, default) vs the round-1 shared-accumulator kernel"""
    from gsgen_b200 import _lib
    from gsgen_b200.rasterizer import render_view

    sc = _small_scene(N=12000, reso=208, scale=2.5).to(DEV)
    cam, c2w = sc.cams[0], sc.c2ws[0].cpu()
    g = torch.Generator().manual_seed(C)
    sh = (0.5 * torch.randn(sc.N, 3, C * C, generator=g)).to(DEV)
    w = torch.randn(cam.h, cam.w, 3, generator=g).to(DEV)
    grads = []
    for variant, slot in ((0, 30), (1, 31)):
        _lib.set_option(torch.device(DEV), slot, _lib.OPT_BWD_SH_VARIANT, variant)
        leaf = lambda t: t.clone().requires_grad_()
        m, q, s, a, p = leaf(sc.mean), leaf(sc.qvec), leaf(sc.svec), leaf(sc.alpha), leaf(sh)
        out = render_view(m, q, s, a, c2w, cam, sh=p, C=C, slot=slot)
        out["rgb"].backward(gradient=w)
        grads.append((m.grad, q.grad, s.grad, a.grad, p.grad))
    for a_, b_, n_ in zip(grads[0], grads[1], ("mean", "qvec", "svec", "alpha", "sh")):
        assert float(b_.abs().max()) > 0
        assert_grad_close(a_, b_, rtol=1e-5, what=f"C={C} g_{n_} direct vs shared-accumulator")


def test_view_stats(oracle_mod):
    from gsgen_b200.rasterizer import render_view, view_stats

    sc = _small_scene(N=8000, reso=192)
    cam, c2w = sc.cams[0], sc.c2ws[0]
    d = sc.to(DEV)
    with torch.no_grad():
        out = render_view(d.mean, d.qvec, d.svec, d.alpha, c2w, cam, sh=d.sh, C=4, slot=40)
    n_dup, n_vis, longest = view_stats(DEV, 40)
    ref = oracle_mod.render_view(sc.mean, sc.qvec, sc.svec, sc.alpha, c2w, ocam_of(cam), sh=sc.sh, C=4)["aux"]
    assert n_dup == ref["D"] == out["aux"]["N_with_dub"]
    assert n_vis == int(ref["mask"].sum())
    lens = torch.where(ref["start"] >= 0, ref["end"] - ref["start"], torch.zeros_like(ref["start"]))
    assert longest == int(lens.max())


# ---- 2 GPUs: replicas stay bit-identical through render -> all-reduce -> Adam -> densify -> prune ------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _replica_worker(rank, world, port, ret):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    import datetime

    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=60))
    try:
        from gsgen_b200.camera import orbit_c2w
        from gsgen_b200.splatting import GaussianSplattingRenderer

        dev = torch.device("cuda", rank)
        torch.manual_seed(100 + rank)  # rank-local RNG streams differ on purpose
        sc = _small_scene(N=3001, reso=128)
        cam = sc.cams[0]
        cfg = {"densify": dict(enabled=True, type="official", warm_up=2, end=100, period=2, mean2d_thresh=1e-7,
                               split_thresh=0.02, n_splits=2, split_shrink=0.8, use_legacy=False),
               "prune": dict(enabled=True, warm_up=0, end=100, period=2, radii2d_thresh=0.0, alpha_thresh=0.06,
                             radii3d_thresh=0.0)}
        r = GaussianSplattingRenderer(cfg, _raw_init(sc), device=dev)
        r.setup_lr({"mean": 1e-3, "svec": 1e-3, "qvec": 1e-3, "color": 1e-2, "alpha": 1e-2})
        opt = r.set_optimizer({"type": "Adam", "opt_args": {"eps": 1e-15}})
        views = [orbit_c2w(2.5, 15.0, 30.0 + 50.0 * v) for v in range(4)]
        mine = [v for v in range(4) if v % world == rank]
        sizes = []
        for step in range(3):
            r.update(step)
            out = r({"c2w": torch.stack([views[v] for v in mine]), "camera_info": [cam] * len(mine)}, use_bg=False)
            g = torch.Generator().manual_seed(10 * step)
            w = torch.randn(4, cam.h, cam.w, 3, generator=g)[mine].to(dev)
            (out["rgb"] * w).sum().backward()
            r.store.all_reduce()
            opt.step()
            r.post_backward()
            n0 = r.N
            r.densify(step)
            r.prune(step)
            sizes.append((n0, r.N))
            opt.zero_grad()
        st = r.store
        live = torch.cat([st._rows(b, name, st.N).reshape(-1) for b in st._buffers() for name in st._field])
        gathered = [torch.empty_like(live) for _ in range(world)]
        dist.all_gather(gathered, live)
        same = all(torch.equal(gathered[0], x) for x in gathered)
        if rank == 0:
            ret["same"], ret["sizes"] = bool(same), sizes
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_replicas_stay_bit_identical():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_replica_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["same"], "arenas (parameters, gradients, Adam moments) differ between the two ranks"
    assert ret["sizes"][2][0] != ret["sizes"][2][1], "densify / prune did not run: the test exercises nothing"


def test_viewer_loop_renders_uint8_frames_in_eval_mode():
    """the viewer side (utils/viewer/viser_viewer.py:129-171): eval-mode render_one under no_grad -> uint8 frame, equal to
    the reference's host-side expression applied to the same image; nothing touches the gradient arena or the statistics"""
    import numpy as np

    from gsgen_b200.splatting import GaussianSplattingRenderer
    from gsgen_b200.viewer import ViewerLoop

    sc = _small_scene(N=6000, reso=128)
    r = GaussianSplattingRenderer({}, _raw_init(sc), device=DEV).eval()
    loop = ViewerLoop(r, resolution=200)
    q = np.array([0.0, 1.0, 0.0, 0.0])  # (w,x,y,z): some rotation; the camera looks at the scene from a distance
    from gsgen_b200.camera import orbit_c2w

    c2w = orbit_c2w(2.5, 15.0, 30.0).numpy()
    # a rotation matrix -> quaternion for the viewer interface (w,x,y,z)
    R = c2w[:3, :3]
    w_ = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    q = np.array([w_, (R[2, 1] - R[1, 2]) / (4 * w_), (R[0, 2] - R[2, 0]) / (4 * w_), (R[1, 0] - R[0, 1]) / (4 * w_)])
    frame = loop.render_frame(fov=0.8, aspect=1.25, wxyz=q, position=c2w[:3, 3])
    assert frame.dtype == np.uint8 and frame.shape == (160, 200, 3) and frame.max() > 0
    with torch.no_grad():
        from gsgen_b200.camera import CameraInfo
        from gsgen_b200.viewer import get_c2w

        cam = CameraInfo.from_fov_camera(0.8, 1.25, 200, 0.01, 100.0)
        img = r.render_one(torch.from_numpy(get_c2w(q, c2w[:3, 3])), cam)["rgb"]
    want = (img.detach().cpu().clamp(min=0.0, max=1.0).numpy() * 255.0).astype(np.uint8)
    assert np.array_equal(frame, want)
    assert float(r.store.flat_grad.abs().max()) == 0 and not r._pending and loop.fps > 0


def _sparse_worker(rank, world, port, ret):
    import datetime

    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=60))
    try:
        from gsgen_b200.camera import orbit_c2w
        from gsgen_b200.parallel import ViewParallelRenderer
        from gsgen_b200.rasterizer import render_view

        dev = torch.device("cuda", rank)
        sc = _small_scene(N=60001, reso=256, scale=4.0)  # N % 4 != 0 on purpose; opaque: most rows stay untouched
        cam = sc.cams[0]
        p = dict(mean=sc.mean, qvec=sc.qvec, svec=sc.svec, alpha=sc.alpha, sh=sc.sh)
        out = {}
        for sparse in (True, False):
            vpr = ViewParallelRenderer(p, 4, dev, sparse_allreduce=sparse, sparse_max_fraction=0.9)
            vpr.zero_grad()
            for v in range(2):  # two views per rank
                c2w = orbit_c2w(1.6, 15.0, 30.0 + 45.0 * (2 * rank + v))  # close: part of the ball is outside the frustum
                g = torch.Generator().manual_seed(10 * rank + v)
                w = torch.randn(cam.h, cam.w, 3, generator=g).to(dev)
                o = render_view(vpr.params["mean"], vpr.params["qvec"], vpr.params["svec"], vpr.params["alpha"], c2w,
                                cam, sh=vpr.params["sh"], C=4, slot=v, grad_sink=vpr.grad_views)
                o["rgb"].backward(gradient=w)
            local = vpr.flat_grad.clone()
            vpr.all_reduce()
            out[sparse] = (vpr.flat_grad.clone(), dict(vpr.last_allreduce), local)
        sp, la, local = out[True]
        dn = out[False][0]
        ok = la["mode"] == "sparse" and 0 < la["rows"] < sc.N
        ok = ok and torch.allclose(sp, dn, rtol=1e-5, atol=1e-6 * float(dn.abs().max()))
        ok = ok and bool((sp[dn == 0] == 0).all())   # rows nobody touched stay exactly zero
        ok = ok and float(dn.abs().max()) > 0
        # every row with a local gradient is inside the union that was reduced (nothing dropped)
        ok = ok and bool(((local != 0) <= (sp != 0) | (dn == 0)).all())
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret["ok"], ret["rows"], ret["N"] = bool(flag.item()), la["rows"], sc.N
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_sparse_allreduce_equals_dense():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sparse_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["ok"], dict(ret)
    assert ret["rows"] < 0.9 * ret["N"]
