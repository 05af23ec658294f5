"""Dry run of bench.py's end-to-end measurement logic on CPU with stand-ins for the CUDA stream / event API and for
the renderer: checks the control flow, the bookkeeping of bytes and runs, the image check of the side-stream schedule
and its fall-back -- not any timing.  (The real thing runs on the GPU box; this keeps a Python-level slip from costing
a bench line.)"""
import contextlib
import importlib
import os
import types

import pytest
import torch


class _Event:
    def __init__(self, enable_timing=False):
        self.recorded = 0

    def record(self, stream=None):
        self.recorded += 1

    def elapsed_time(self, other):
        return 50.0

    def wait(self, stream=None):
        pass


class _Stream:
    def wait_event(self, ev):
        assert isinstance(ev, _Event)

    def wait_stream(self, s):
        assert isinstance(s, _Stream)


@pytest.fixture()
def bench_mod(monkeypatch):
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS")}
    mod = importlib.import_module("bench")
    for k, v in saved.items():  # importing bench.py sets thread-count defaults for ITS process; undo for the suite
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: _Stream())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None)
    return mod


class _Vpr:
    def __init__(self, n):
        mk = lambda *shape: torch.randn(*shape).requires_grad_()
        self.params = dict(mean=mk(n, 3), qvec=mk(n, 4), svec=mk(n, 3), alpha=mk(n), sh=mk(n, 3, 4))
        self.grad_views = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self.n_zero = self.n_reduce = 0

    def zero_grad(self):
        self.n_zero += 1

    def all_reduce(self):
        self.n_reduce += 1


def _setup(n_views=2, H=8, W=6):
    cams = [types.SimpleNamespace(h=H, w=W) for _ in range(n_views)]
    c2ws = [torch.eye(4)[:3] * (v + 1) for v in range(n_views)]
    gouts = [torch.randn(H, W, 3) for _ in range(n_views)]
    calls = []

    def render_view(mean, qvec, svec, alpha, c2w, cam, sh=None, C=1, slot=0, grad_sink=None):
        calls.append((slot, grad_sink is not None, torch.is_grad_enabled()))
        img = (sh.sum() * 0 + c2w[0, 0]) * torch.ones(cam.h, cam.w, 3) + mean.sum() * 0
        return {"rgb": img}

    return cams, c2ws, gouts, render_view, calls


@pytest.mark.parametrize("world", [1, 2])
def test_measure_e2e_bookkeeping(bench_mod, monkeypatch, world):
    import torch.distributed as dist

    monkeypatch.setattr(dist, "all_reduce", lambda t, op=None, group=None: None)
    cams, c2ws, gouts, render_view, calls = _setup()
    vpr = _Vpr(5)
    args = types.SimpleNamespace(steps=4)
    mine = [0, 1]
    barriers = []
    e2e = bench_mod.measure_e2e(args, vpr, mine, render_view, c2ws, cams, gouts, 2, {0: 0, 1: 1}, 5, 2, world, "cpu",
                                lambda: barriers.append(1))
    H, W = 8, 6
    assert e2e["h2d_bytes_per_step"] == 2 * (H * W * 3 * 4 + 240) and e2e["d2h_bytes_per_step"] == 2 * H * W * 3 * 4
    assert e2e["ms_per_step"] == 50.0 / 4 and len(e2e["ms_per_step_all_runs"]) == 3
    assert e2e["value"] == pytest.approx(2 * 5 * H * W / (12.5e-3))
    loops = 1 if world > 1 else 2  # the side-stream schedule runs at world_size 1 only
    assert vpr.n_zero == vpr.n_reduce == loops * (3 + 3 * 4)
    assert len(barriers) == loops * 6
    if world == 1:
        assert e2e["copy_schedule"].startswith("copies on side streams") and e2e["side_stream_error"] is None
        assert calls[-1][2] is False  # the image check re-renders under no_grad
    else:
        assert e2e["copy_schedule"] == "copies on the launching stream" and "world_size" in e2e["side_stream_error"]
    assert all(c[1] for c in calls[:-1])  # every timed call accumulates into the flat gradient buffer


def test_measure_e2e_falls_back_when_the_host_image_is_wrong(bench_mod):
    cams, c2ws, gouts, render_view, calls = _setup(n_views=1)
    n = [0]

    def drifting(*a, **kw):  # every call returns a different image: the D2H check of the side-stream schedule fails
        n[0] += 1
        out = render_view(*a, **kw)
        return {"rgb": out["rgb"] + n[0]}

    e2e = bench_mod.measure_e2e(types.SimpleNamespace(steps=2), _Vpr(3), [0], drifting, c2ws, cams, gouts, 2, {0: 0}, 3,
                                1, 1, "cpu", lambda: None)
    assert e2e["copy_schedule"] == "copies on the launching stream"
    assert "differs" in e2e["side_stream_error"] and e2e["value"] > 0
