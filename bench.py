#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200 (see BASELINE.json / SURVEY.md §8(d)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c2|c4|c5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward + backward pass of the hot path over one batch of views of a synthetic scene:
cull -> project -> tile bin / radix sort -> SH composite forward -> composite backward -> projection backward
(-> one NCCL all-reduce of the flat gradient buffer when N > 1).  Default workload = BASELINE config C3, the
configuration the metric is quoted on: 1M Gaussians, 1024x1024, SH degree 3.  With N GPUs each rank renders
one view of its own (weak scaling: per-GPU work fixed) and the gradients of all N views are all-reduced.
Metric: Gaussians x pixels / s = (sum over rendered views of N_gauss*H*W) / time.

`--impl reference` times the reference algorithm's CPU restatement (oracle/, kind "port": the reference has
no CPU path of its own and its CUDA extension is not a CPU implementation) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def usable_cpus() -> int:
    """CPUs this container may actually use: min(affinity, cgroup quota).  The GPU boxes show 128 cores but carry a
    CFS quota (cpu.max 1600000/100000 = 16 CPUs, measured round 1); threads beyond the quota only get the whole
    process throttled in 100 ms slices -- which showed up as random 100-250 ms stalls in the timed loop."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(int(q) / int(per))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return n


_IS_REFERENCE_ARM = any(a == "reference" or a == "--impl=reference" for a in sys.argv[1:])
if _IS_REFERENCE_ARM:
    os.environ.setdefault("OMP_NUM_THREADS", str(usable_cpus()))
else:
    # the GPU arm's host work is a few tiny CPU tensor ops per view: an OpenMP pool of 64 spinning workers would
    # only burn the container's CPU quota
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")

import torch  # noqa: E402

METRIC = "fwd+bwd Gaussians*pixels/s"
UNIT = "Gaussian*pixel/s"
SH_C = {"c1": 1, "c2": 3, "c3": 4, "c4": 4, "c5": 4}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(SH_C))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--svec-scale", type=float, default=1.0, help="C5 tile-occupancy sweep")
    ap.add_argument("--clock-period", type=float, default=0.1, help="NVML sampling period in s (0 = off)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------
def make_views(workload, scene, n_views):
    """(cams, c2ws) for the step.  c4 has its own 8 orbit views; the single-view configs get one extra
    orbit pose per additional rank (azimuth +45 deg each) so that every GPU renders a different view."""
    from gsgen_b200.camera import orbit_c2w

    if workload == "c4":
        return scene.cams, scene.c2ws
    cam = scene.cams[0]
    c2ws = [orbit_c2w(2.5, 15.0 if workload != "c2" else 20.0, (30.0 if workload != "c2" else 45.0) + 45.0 * v)
            for v in range(n_views)]
    return [cam] * n_views, c2ws


def algorithmic_bytes(C, D, D_eff, N, N0, T, H, W):
    """SURVEY.md §8(d) per-stage algorithmic bytes of one view."""
    K = 3 * C * C
    b_inst = 4 + 4 * (7 + K)
    return {
        "preprocess": N0 * 25 + 84 * N,                 # cull + project(+aabb)
        "scan": N0 * 8,
        "bin": N * 20 + D * 12 + D * 24 + D * 8 + T * 8,  # key emit + sort + ranges
        "composite_fwd": D_eff * b_inst + 8 * T + 16 * H * W,
        "composite_bwd": D_eff * b_inst + D_eff * 4 * (7 + K) + 8 * T + 24 * H * W,
        "project_bwd": N * 40 + N * 28 + N * 40,
        "b_inst": b_inst,
    }


class ClockSampler:
    """SM clock / power / throttle reasons sampled from NVML on a background thread (every 10 ms), the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.*` prints.  A thread
    inside this process avoids spawning nvidia-smi next to the timed region (its NVML start-up stalled CUDA calls for
    hundreds of ms in round 1).  mark() returns the sample index; stop(begin, end) summarises that window."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, period_s=0.01):
        self.gpu_index, self.period = gpu_index, period_s
        self.samples = []  # (sm_mhz, power_w, reasons_bitmask)
        self.ok = False
        self._stop = False
        self.t = None

    def start(self):
        try:
            import threading

            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def mark(self):
        return len(self.samples)

    def stop(self, begin=0, end=None):
        self._stop = True
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: " + getattr(self, "err", "?")]}
        if self.t is not None:
            self.t.join(timeout=1.0)
        win = self.samples[max(0, begin - 1):(None if end is None else end + 1)]
        if not win:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"]}
        mask = 0
        for _, _, r in win:
            mask |= r
        reasons = sorted(n for b, n in self.REASONS.items() if mask & b)
        return {"sm_mhz": statistics.median(x[0] for x in win), "sm_max_mhz": self.sm_max,
                "power_w_max": max(x[1] for x in win), "samples": len(win), "reasons": reasons,
                "source": "NVML (pynvml), %g s period, window = timed region" % self.period}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """dram bytes per launch of the forward composite from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload, {}).get("composite_fwd_dram_bytes")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------------
# CPU restatement (oracle) -- cpu_baseline leg and the --impl reference arm.  The only place bench.py
# executes oracle/.
# ------------------------------------------------------------------------------------------------------
def cpu_step_factory(workload, scene, cam, c2w, window_frac=1.0):
    import oracle

    C = SH_C[workload]
    ocam = oracle.Cam(cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h, cam.near_plane, cam.far_plane)
    g = torch.Generator().manual_seed(scene.seed + 100)
    gout = torch.randn(cam.h, cam.w, 3, generator=g)
    cfg = oracle.view_cfg(ocam)
    th, tw = cfg["n_tiles_h"], cfg["n_tiles_w"]
    keep = torch.ones(th, tw, dtype=torch.bool)
    if window_frac < 1.0:  # centred window of tiles
        import math

        s = math.sqrt(window_frac)
        hh, ww = max(1, int(round(th * s))), max(1, int(round(tw * s)))
        y0, x0 = (th - hh) // 2, (tw - ww) // 2
        keep[:] = False
        keep[y0:y0 + hh, x0:x0 + ww] = True
    px = 0
    for ty in range(th):
        for tx in range(tw):
            if keep[ty, tx]:
                px += (min(16, cam.h - 16 * ty)) * (min(16, cam.w - 16 * tx))
    keep_flat = keep.view(-1)

    def step():
        """-> (seconds in the per-Gaussian + binning stages incl. their backward, seconds in the window composite fwd+bwd)"""
        t0 = time.perf_counter()
        mean = scene.mean.clone().requires_grad_()
        qvec = scene.qvec.clone().requires_grad_()
        svec = scene.svec.clone().requires_grad_()
        alpha = scene.alpha.clone().requires_grad_()
        sh = scene.sh.clone().requires_grad_()
        normals, pts = oracle.get_frustum(ocam, c2w)
        mask = oracle.cull_bsphere(mean.detach(), svec.detach(), normals, pts, 6.0)
        m, q, s, a = mean[mask].contiguous(), qvec[mask].contiguous(), svec[mask].contiguous(), alpha[mask].contiguous()
        m2, c2, _, dp = oracle.project_gaussians(m, q, s, c2w, True)
        D, tl, br = oracle.tile_culling_aabb_count(m2.detach(), c2.detach(), 16, ocam, 6.0)
        ids, start, end = oracle.tile_culling_aabb_start_end(tl, br, dp.detach(), th, tw, D)
        if window_frac < 1.0:
            start = torch.where(keep_flat, start, torch.full_like(start, -1))
            end = torch.where(keep_flat, end, torch.full_like(end, -1))
        topleft = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=torch.float32)
        shm = sh[mask].contiguous()
        # composite forward + backward on detached leaves (timed separately), then the chain rule through the
        # per-Gaussian stages with the gradients it produced
        m2d, c2d = m2.detach().requires_grad_(), c2.detach().requires_grad_()
        shd, ad = shm.detach().requires_grad_(), a.detach().requires_grad_()
        t1 = time.perf_counter()
        rgb = oracle.render_sh(m2d, c2d, shd, ad, start, end, ids, topleft, c2w, C, cfg, None)
        rgb.backward(gradient=gout)
        t2 = time.perf_counter()
        torch.autograd.backward([m2, c2, shm, a], [m2d.grad, c2d.grad, shd.grad, ad.grad])
        t3 = time.perf_counter()
        return (t1 - t0) + (t3 - t2), (t2 - t1)

    return step, px


def run_cpu_baseline(args):
    """cpu_baseline leg: the oracle (CPU restatement of the reference path) timed on the host cores, in a child
    process so that it gets its own OpenMP pool sized to the usable CPUs (the GPU arm runs single-threaded)."""
    n = usable_cpus()
    env = dict(os.environ, OMP_NUM_THREADS=str(n), MKL_NUM_THREADS=str(n), CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "1",
           "--warmup", "0", "--svec-scale", str(args.svec_scale)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        raise RuntimeError("reference arm printed no JSON: " + out.stderr[-300:])
    d = json.loads(line[-1])
    cb = d["cpu_baseline"]
    cb["seconds_per_step"] = d["ms_per_step"] / 1e3
    return cb


def run_reference_arm(args):
    """--impl reference: the reference algorithm's CPU restatement, same metric/config, K timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from gsgen_b200.scenes import make_scene

    oracle.build()
    threads = int(os.environ.get("OMP_NUM_THREADS", usable_cpus()))
    torch.set_num_threads(threads)
    wl = args.workload
    scene = make_scene(wl, svec_scale=args.svec_scale)
    cams, c2ws = make_views(wl, scene, 1)
    cam, c2w = cams[0], c2ws[0]
    # bounded sample.  The per-Gaussian stages (project, 5M-key sort, their backward) do not shrink with a tile window,
    # so a windowed step would understate the reference if its time were divided by the window's pixels only.  Rule:
    # composite the FULL image whenever K+W such steps fit in ~4 minutes (16 usable cores: ~2.1 s/step); otherwise
    # composite a centred window of fraction f and report value = N*H*W / (t_per_gaussian + t_composite / f), i.e. the
    # composite time extrapolated to the whole image.
    total = args.steps + args.warmup
    step, px = cpu_step_factory(wl, scene, cam, c2w, 1.0)
    tp, tc = step()
    frac = 1.0
    if (tp + tc) * total > 240.0:
        for f in (0.25, 1.0 / 16):
            frac = f
            if (tp + tc * f) * total <= 240.0:
                break
        step, px = cpu_step_factory(wl, scene, cam, c2w, frac)
    for _ in range(args.warmup):
        step()
    acc_p = acc_c = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a_, b_ = step()
        acc_p += a_
        acc_c += b_
    dt = (time.perf_counter() - t0) / args.steps
    H, W = cam.h, cam.w
    t_full = acc_p / args.steps + (acc_c / args.steps) / frac
    val = scene.N * H * W / t_full
    sample = (f"each step = all per-Gaussian stages + binning on {scene.N} Gaussians (fwd+bwd) + SH composite fwd+bwd "
              f"on a centred window of {frac:.4g} of the tiles ({px} px); measured {dt:.3f} s/step of which "
              f"{acc_c / args.steps:.3f} s composite; value = N*H*W / (t_per_gaussian + t_composite/{frac:.4g})")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(wl, scene, cam), "views_per_step": 1, "parallelism": "cpu-openmp"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(wl, scene, cam):
    return (f"{wl}: {scene.N} Gaussians, {cam.w}x{cam.h}, SH deg {SH_C[wl] - 1} (C={SH_C[wl]}), tile 16, "
            f"T_thresh 1e-4, radii 6.0")


# ------------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------------
def last_rgb_of(vpr, mine, render_view, c2ws_cpu, cams, C, slot_of, h_c2w, v0):
    """re-render the last view of the e2e step (no grad): the reference image for the D2H check"""
    v = mine[-1]
    with torch.no_grad():
        out = render_view(vpr.params["mean"], vpr.params["qvec"], vpr.params["svec"], vpr.params["alpha"],
                          h_c2w if v == v0 else c2ws_cpu[v], cams[v], sh=vpr.params["sh"], C=C, slot=slot_of[v])
    return out["rgb"]


def measure_e2e(args, vpr, mine, render_view, c2ws_cpu, cams, gouts, C, slot_of, N, n_views, world, dev, barrier):
    """The metric through the public API with HOST buffers: every step copies the upstream gradient image from pinned
    host memory (H2D), renders + back-propagates through render_view(), and copies the rendered image to pinned host
    memory (D2H).  Measured twice: everything on the launching stream, and with the two copies on side streams."""
    import torch.distributed as dist

    v0 = mine[0]
    H, W = cams[v0].h, cams[v0].w
    h_c2w = c2ws_cpu[v0].clone().pin_memory()
    h_gout = gouts[v0].cpu().pin_memory()
    h_rgb = torch.empty(H, W, 3).pin_memory()
    d_gout = torch.empty_like(gouts[v0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def view_args(v):
        return (vpr.params["mean"], vpr.params["qvec"], vpr.params["svec"], vpr.params["alpha"],
                h_c2w if v == v0 else c2ws_cpu[v], cams[v])

    def e2e_step_serial(i):
        vpr.zero_grad()
        for v in mine:
            d_gout.copy_(h_gout, non_blocking=True)
            out = render_view(*view_args(v), sh=vpr.params["sh"], C=C, slot=slot_of[v], grad_sink=vpr.grad_views)
            out["rgb"].backward(gradient=d_gout)
            h_rgb.copy_(out["rgb"].detach(), non_blocking=True)
        vpr.all_reduce()

    # Same bytes, same calls, but the two copies ride their own streams: the upstream-gradient H2D of a view
    # overlaps that view's forward, the image D2H overlaps its backward (PCIe is full duplex; the copy engines
    # are independent of the SMs).  Every step still moves both buffers and the loop ends with all streams joined.
    cs_in, cs_out = torch.cuda.Stream(), torch.cuda.Stream()
    d_gout2 = [torch.empty_like(d_gout), torch.empty_like(d_gout)]
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]
    ev_free = [torch.cuda.Event(), torch.cuda.Event()]
    ev_img = torch.cuda.Event()
    n_copy = [0]

    def e2e_step_pipelined(i):
        main = torch.cuda.current_stream()
        vpr.zero_grad()
        for v in mine:
            k = n_copy[0] & 1
            n_copy[0] += 1
            with torch.cuda.stream(cs_in):
                cs_in.wait_event(ev_free[k])  # the backward that last read this buffer is done
                d_gout2[k].copy_(h_gout, non_blocking=True)
                ev_in[k].record(cs_in)
            out = render_view(*view_args(v), sh=vpr.params["sh"], C=C, slot=slot_of[v], grad_sink=vpr.grad_views)
            rgb = out["rgb"]
            ev_img.record(main)
            with torch.cuda.stream(cs_out):
                cs_out.wait_event(ev_img)
                h_rgb.copy_(rgb.detach(), non_blocking=True)
            rgb.record_stream(cs_out)
            main.wait_event(ev_in[k])
            rgb.backward(gradient=d_gout2[k])
            ev_free[k].record(main)
        vpr.all_reduce()

    def e2e_measure(step_fn):
        for i in range(3):
            step_fn(i)
        runs_ = []
        for _ in range(3):  # best of three K-step loops, like the device-resident number (host CFS throttling)
            barrier(); torch.cuda.synchronize()
            e0.record()
            for i in range(args.steps):
                step_fn(i)
            main = torch.cuda.current_stream()
            main.wait_stream(cs_in); main.wait_stream(cs_out)  # the step's copies belong to the timed region
            e1.record()
            torch.cuda.synchronize(); barrier()
            t2 = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            runs_.append(float(t2.item()))
        return runs_

    serial_runs = e2e_measure(e2e_step_serial)
    mode, e2e_runs, pipe_err = "copies on the launching stream", serial_runs, None
    if world == 1:  # the side-stream schedule was validated on one GPU only in round 1
        try:
            pipe_runs = e2e_measure(e2e_step_pipelined)
            torch.cuda.synchronize()
            # the image that reached the host must be the image the device holds
            ref_img = last_rgb_of(vpr, mine, render_view, c2ws_cpu, cams, C, slot_of, h_c2w, v0)
            img_diff = float((h_rgb.to(dev) - ref_img).abs().max())
            if img_diff != 0.0:
                raise RuntimeError(f"side-stream D2H image differs from the device image by {img_diff}")
            mode, e2e_runs = "copies on side streams (H2D under the forward, D2H under the backward)", pipe_runs
        except Exception as ex:  # report the single-stream number rather than no number
            pipe_err = repr(ex)
    else:
        pipe_err = "not enabled for world_size > 1"
    ms_e2e = min(e2e_runs)
    bi = len(mine) * (h_gout.numel() * 4 + 240)  # gradient image + the by-value camera struct
    bo = len(mine) * h_rgb.numel() * 4
    return {"value": n_views * N * H * W / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
            "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo, "ms_per_step_all_runs": e2e_runs,
            "copy_schedule": mode, "ms_per_step_single_stream": min(serial_runs),
            "ms_per_step_single_stream_all_runs": serial_runs, "side_stream_error": pipe_err,
            "what": "render_view()+backward through the public API; per step: host camera pose (by-value kernel "
                    "argument) + pinned upstream gradient image H2D, rendered image D2H to pinned memory; Gaussian "
                    "parameters stay resident (they are the model state, like weights)"}


def run_ours(args):
    import ctypes

    import torch.distributed as dist

    from gsgen_b200 import _lib
    from gsgen_b200.parallel import ViewParallelRenderer, shard_views
    from gsgen_b200.rasterizer import render_view
    from gsgen_b200.scenes import make_scene

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        # a mismatched collective must fail fast, not burn GPU time until the default 10-minute watchdog
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    wl = args.workload
    C = SH_C[wl]
    scene = make_scene(wl, svec_scale=args.svec_scale)
    n_views = 8 if wl == "c4" else world
    cams, c2ws_cpu = make_views(wl, scene, n_views)
    H, W = cams[0].h, cams[0].w
    vpr = ViewParallelRenderer(dict(mean=scene.mean, qvec=scene.qvec, svec=scene.svec, alpha=scene.alpha,
                                    sh=scene.sh), C, dev)
    mine = shard_views(n_views, rank, world)
    gouts = {}
    for v in mine:
        g = torch.Generator().manual_seed(scene.seed + 100 + v)
        gouts[v] = torch.randn(cams[v].h, cams[v].w, 3, generator=g).to(dev)
    slot_of = {v: i for i, v in enumerate(mine)}
    last = {}

    def render_and_backward(params, v):
        # the pose is host data (it comes from the data loader): passing the CPU tensor avoids a D2H sync
        out = render_view(params["mean"], params["qvec"], params["svec"], params["alpha"], c2ws_cpu[v], cams[v],
                          sh=params["sh"], C=C, slot=slot_of[v], grad_sink=vpr.grad_views)
        out["rgb"].backward(gradient=gouts[v])
        last["rgb"], last["aux"] = out["rgb"], out["aux"]

    def step():
        vpr.step(n_views, render_and_backward)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- warm-up (clock sampler already running)
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]), args.clock_period)
    if rank == 0 and args.clock_period > 0:
        sampler.start()
        time.sleep(0.5)
    for _ in range(max(3, args.warmup)):
        step()
    # ... plus ~1.5 s of extra untimed steps: the first backward spawns autograd / CUDA helper threads and the
    # container's CPU quota needs a few periods to settle (measured: sporadic 100-250 ms host stalls otherwise).
    # The COUNT is agreed across ranks (every step contains a collective): max over ranks of a 5-step estimate.
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    est = torch.tensor([(time.perf_counter() - t_w) / 5], device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    n_extra = int(min(1000, max(10, 1.5 / max(float(est.item()), 1e-4))))
    for _ in range(n_extra):
        step()
    torch.cuda.synchronize()
    ctxs = [_lib.ctx(dev, s) for s in slot_of.values()]
    hm = (ctypes.c_float * 6)()
    hc = (ctypes.c_int64 * 5)()

    def timed_loop():
        """exactly K steps, barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); torch.cuda.synchronize()
        m0 = sampler.mark() if rank == 0 else 0
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize(); barrier()
        m1 = sampler.mark() if rank == 0 else 0
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), m0, m1

    # ---- timed region (the reported value): no instrumentation inside.  The K-step loop is run three times and the
    # fastest is reported (all three are listed): the GPU boxes throttle the container's CPU in 100 ms CFS slices
    # (cpu.max = 16 CPUs), and a throttled host stalls the launch stream for up to one slice -- a measurement
    # artefact of the host, like a thermal event, not a property of the path.
    runs = [timed_loop() for _ in range(3)]
    ms_max, mark0, mark1 = min(runs, key=lambda r: r[0])
    all_runs_ms = [r[0] for r in runs]
    # ---- the same K steps again with every stage bracketed by CUDA events on the launching stream
    # (gsb200_ctx_set_profiling).  Kept out of the headline loop because the bracketing perturbs it (reported).
    for c in ctxs:
        _lib.check(_lib.lib().gsb200_ctx_set_profiling(c, 1))
    step()
    torch.cuda.synchronize()
    for c in ctxs:
        _lib.check(_lib.lib().gsb200_ctx_get_profile(c, hm, hc, 1))
    ms_profiled, _, _ = timed_loop()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- stage profile of the timed region (events recorded on the launching stream)
    stage_ms = [0.0] * 6
    counts = [0] * 5
    for c in ctxs:
        _lib.check(_lib.lib().gsb200_ctx_get_profile(c, hm, hc, 1))
        for i in range(6):
            stage_ms[i] += float(hm[i])
        for i in range(5):
            counts[i] += int(hc[i])
        _lib.check(_lib.lib().gsb200_ctx_set_profiling(c, 0))
    n_fwd = max(1, counts[0])
    D = counts[2] / n_fwd
    D_eff = counts[3] / n_fwd
    staged = counts[4] / n_fwd
    aux = last["aux"]
    N_vis = int(aux["mask"].sum().item())
    th, tw = cams[0].n_tiles
    ab = algorithmic_bytes(C, D, D_eff, N_vis, scene.N, th * tw, H, W)
    names = ["preprocess", "scan", "bin", "composite_fwd", "composite_bwd", "project_bwd"]
    stages = {}
    for i, nm in enumerate(names):
        per = stage_ms[i] / n_fwd
        stages[nm] = {"ms": per, "alg_gbs": (ab[nm] / 1e9) / (per / 1e3) if per > 0 else None}
    peak, peak_src = measured_peaks()
    fwd_ms = stages["composite_fwd"]["ms"]
    achieved = (ab["composite_fwd"] / 1e9) / (fwd_ms / 1e3) if fwd_ms > 0 else 0.0
    roofline = {"kernel": "k_composite_fwd<SH,C=%d>" % C, "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(wl), "peak_source": peak_src,
                "alg_bytes_per_launch": ab["composite_fwd"], "avg_launch_ms": fwd_ms,
                "alg_bytes_formula": "D_eff*(4+4*(7+3C^2)) + 8*tiles + 16*H*W (SURVEY.md §8(d))"}

    # ---- end to end through the public API with host buffers (pinned): per step H2D of the step's inputs
    # (camera pose + upstream gradient image) and D2H of the rendered image
    e2e = None
    if not args.no_e2e:
        try:
            e2e = measure_e2e(args, vpr, mine, render_view, c2ws_cpu, cams, gouts, C, slot_of, scene.N, n_views,
                              world, dev, barrier)
        except Exception as ex:  # the bench line must still print; every rank takes the same branch
            e2e = {"value": None, "unit": UNIT, "error": repr(ex)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    time.sleep(0.1)
    clocks = sampler.stop(mark0, mark1)
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu_base = run_cpu_baseline(args)
        except Exception as e:  # the bench line must still print
            cpu_base = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    value = n_views * scene.N * H * W / (ms_max / 1e3)
    # preprocess, depth_keys, gather_counts, fill x2, emit_tiles, tile_ranges, composite_fwd, composite_bwd, project_bwd
    my_kernels_per_view = 10
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max, "higher_is_better": True, "scaling": "strong" if wl == "c4" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(wl, scene, cams[0]), "views_per_step": n_views,
                   "views_per_gpu": len(mine), "parallelism": f"view-dp{world}",
                   "grad_allreduce_bytes": vpr.grad_bytes() if world > 1 else 0,
                   "l2": "no explicit flush: inputs larger than L2 -- per step %.0f MB of Gaussian parameters + as many "
                         "gradients + %.0f MB of sort keys/ids stream through the 126 MB L2"
                         % (vpr.grad_bytes() / 1e6, D * 12 / 1e6)},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": my_kernels_per_view * len(mine) * args.steps,
        "library_launches_note": "plus cub::DeviceScan (2 kernels), 2x cub::DeviceRadixSort onesweep (6 + 4 kernels) and 2 memsets per view",
        "roofline": roofline,
        "stages": stages,
        "ms_per_step_with_stage_events": ms_profiled,
        "ms_per_step_all_runs": all_runs_ms,
        "view_stats": {"N_visible": N_vis, "N_with_dub": D, "D_eff": D_eff, "entries_staged_fwd": staged,
                       "tiles": th * tw},
        # SURVEY.md §8(d) number (2): tile-list pairs P = sum_tiles (end - start) * 256 = N_with_dub * 256 per view,
        # identical for the reference and for us when the binning matches (it is bit-exact)
        "tile_list_pairs_per_s": n_views * D * 256.0 / (ms_max / 1e3),
        "cpu_baseline": cpu_base,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
