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


def _setup(n_views=2, H=8, W=6, world=1, n=5):
    cams = [types.SimpleNamespace(h=H, w=W) for _ in range(n_views)]
    c2ws = [torch.eye(4)[:3] * (v + 1) for v in range(n_views)]
    gouts = {v: torch.randn(H, W, 3) for v in range(n_views)}
    calls = []

    def render_view(mean, qvec, svec, alpha, c2w, cam, sh=None, C=1, slot=0, grad_sink=None, async_count=False):
        calls.append((slot, grad_sink is not None, torch.is_grad_enabled()))
        img = (sh.sum() * 0 + c2w[0, 0]) * torch.ones(cam.h, cam.w, 3) + mean.sum() * 0
        return {"rgb": img}

    w = types.SimpleNamespace(vpr=_Vpr(n), mine=list(range(n_views)), cams=cams, c2ws_cpu=c2ws, C=2,
                              slot_of={v: v for v in range(n_views)}, dev="cpu", world=world, render_view=render_view,
                              gouts=gouts, async_count=False, n_views=n_views, scene=types.SimpleNamespace(N=n))
    return w, calls


@pytest.mark.parametrize("world", [1, 2])
def test_measure_e2e_bookkeeping(bench_mod, monkeypatch, world):
    import torch.distributed as dist

    monkeypatch.setattr(dist, "all_reduce", lambda t, op=None, group=None: None)
    w, calls = _setup(world=world)
    args = types.SimpleNamespace(steps=4)
    barriers = []
    e2e = bench_mod.measure_e2e(args, w, lambda: barriers.append(1))
    H, W = 8, 6
    assert e2e["h2d_bytes_per_step"] == 2 * (H * W * 3 * 4 + 240) and e2e["d2h_bytes_per_step"] == 2 * H * W * 3 * 4
    assert e2e["ms_per_step"] == 50.0 / 4 and len(e2e["ms_per_step_all_runs"]) == bench_mod.N_LOOPS
    assert e2e["value"] == pytest.approx(2 * 5 * H * W / (12.5e-3))
    loops = 2  # single-stream schedule, then the side-stream schedule (all world sizes since round 2)
    assert w.vpr.n_zero == w.vpr.n_reduce == loops * (3 + bench_mod.N_LOOPS * 4)
    assert len(barriers) == loops * 2 * bench_mod.N_LOOPS
    assert e2e["copy_schedule"].startswith("copies on side streams") and e2e["side_stream_error"] is None
    assert calls[-1][2] is False  # the image check re-renders under no_grad
    assert all(c[1] for c in calls[:-1])  # every timed call accumulates into the flat gradient buffer


def test_measure_e2e_falls_back_when_the_host_image_is_wrong(bench_mod):
    w, calls = _setup(n_views=1, n=3)
    n = [0]
    inner = w.render_view

    def drifting(*a, **kw):  # every call returns a different image: the D2H check of the side-stream schedule fails
        n[0] += 1
        out = inner(*a, **kw)
        return {"rgb": out["rgb"] + n[0]}

    w.render_view = drifting
    e2e = bench_mod.measure_e2e(types.SimpleNamespace(steps=2), w, lambda: None)
    assert e2e["copy_schedule"] == "copies on the launching stream"
    assert "differs" in e2e["side_stream_error"] and e2e["value"] > 0


def test_reference_arm_sets_its_thread_count(monkeypatch):
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm must override it (round-1 SCALE runs at N >= 2 timed
    the reference on ONE core)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, os; sys.argv=['bench.py','--impl','reference']; os.environ['OMP_NUM_THREADS']='1'; "
            "import importlib.util as u; sp=u.spec_from_file_location('b', %r); m=u.module_from_spec(sp); "
            "sp.loader.exec_module(m); print(os.environ['OMP_NUM_THREADS'], m.usable_cpus())" % os.path.join(root, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    got, want = out.stdout.split()[-2:]
    assert got == want, out.stdout + out.stderr


def test_roofline_extras_compulsory_bytes_and_dram_fraction():
    """the compulsory lower bound and the DRAM-side fraction next to the contract-definition number"""
    import bench

    r = bench.roofline_extras(4, 1_000_000, 686_820.0, 4096, 1024, 1024, 65_536_000.0, 0.379, 6564.2)
    assert r["compulsory_bytes"] == 1_000_000 * 4 * 55 + 4 * 686_820 + 8 * 4096 + 16 * 1024 * 1024
    assert abs(r["dram_gbs"] - 65.536 / 0.379) < 1e-6 and abs(r["dram_frac"] - r["dram_gbs"] / 6564.2) < 1e-12
    assert 0 < r["compulsory_frac"] < 1
    r = bench.roofline_extras(4, 10, 5.0, 4, 32, 32, None, 0.0, 6564.2)  # no capture / no time: no division
    assert "dram_frac" not in r and r["compulsory_frac"] is None


def test_dominant_kernel_roofline_from_a_committed_bench_line():
    """fed with the stages of the last committed bench line and profiles/traffic.json"""
    import json
    import os

    import bench

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.loads(open(os.path.join(root, "profiles", "r2_bench_c3_default_final2_async.json")).read())
    ncu = bench.ncu_numbers("c3")
    ab = {k: v["alg_gbs"] * 1e9 * v["ms"] / 1e3 for k, v in line["stages"].items()}
    r = bench.dominant_kernel_roofline(line["stages"], ab, ncu, 6564.2, 1965.0)
    assert r["stage"] == "composite_bwd" and 0.5 < r["share_of_stage_sum"] < 0.65
    assert abs(r["frac"] - line["stages"]["composite_bwd"]["frac_of_hbm_peak"]) < 1e-9
    assert 0.5 < r["issue_roofline"]["frac"] < 0.6 and 0.015 < r["dram_frac"] < 0.02
    assert r["smem_data_pipe_pct_in_capture"] == 73.5
    # no capture: still a record
    r2 = bench.dominant_kernel_roofline(line["stages"], ab, {}, 6564.2, 1965.0)
    assert r2["traffic"] is None and "issue_roofline" not in r2
