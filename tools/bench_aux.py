"""Measurement tool for the §8(f) rows (not bench.py): times the flat Adam step and the raw-leaf front end at the
BASELINE C3 size on one B200 with CUDA events and reports algorithmic GB/s against MEASURED_PEAKS.json.

    gpurun -- python tools/bench_aux.py > gpurun_out/bench_aux.json
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gsgen_b200.optim import FlatAdam  # noqa: E402
from gsgen_b200.parallel import field_layout  # noqa: E402
from gsgen_b200.rasterizer import render_view  # noqa: E402
from gsgen_b200.scenes import make_scene  # noqa: E402

DEV = "cuda"


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)


def main():
    peak = 6564.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    res = {"device": torch.cuda.get_device_name(0), "hbm_gbs_peak": peak}
    # ---- flat Adam at C3: N = 1M, SH deg 3 -> 59 floats per Gaussian
    N, C = 1_000_000, 4
    layout = field_layout(N, C)
    total = layout[-1][2] + layout[-1][3]
    p = torch.randn(total, device=DEV)
    g = torch.randn(total, device=DEV) * 1e-3
    opt = FlatAdam(p, g, layout, {"mean": [0.005, 3e-5, 15000, "exp"], "svec": [0.003, 0.001, 15000, "exp"],
                                  "qvec": 0.003, "sh": 0.01, "alpha": 0.003})
    med, best = timeit(lambda: opt.step())
    alg = total * 28  # 16 B read (param, grad, exp_avg, exp_avg_sq) + 12 B written per parameter
    res["adam_flat_c3"] = {"elements": total, "alg_bytes": alg, "ms_median": med, "ms_min": best,
                           "alg_gbs": alg / 1e9 / (med / 1e3), "frac_of_hbm_peak": alg / 1e9 / (med / 1e3) / peak}
    # torch.optim.Adam (foreach) on the same fields as separate tensors, the reference's configuration
    params = [torch.randn(n, device=DEV).view(shape).requires_grad_() for _, shape, _, n in layout]
    for q in params:
        q.grad = torch.randn_like(q) * 1e-3
    topt = torch.optim.Adam([{"params": [q], "lr": 0.003} for q in params], lr=0.0, eps=1e-15)
    med_t, best_t = timeit(lambda: topt.step())
    res["adam_torch_foreach_c3"] = {"ms_median": med_t, "ms_min": best_t, "speedup_of_flat": med_t / med}
    del params, topt, opt, p, g
    # ---- whole view forward+backward at C3 with raw leaves vs torch activations in front
    sc = make_scene("c3").to(DEV)
    cam, c2w = sc.cams[0], sc.c2ws[0].cpu()
    w = torch.randn(cam.h, cam.w, 3, device=DEV)
    raw_s, raw_a = torch.log(sc.svec), torch.logit(sc.alpha)

    def leaves():
        return [t.clone().requires_grad_() for t in (sc.mean, sc.qvec, raw_s, raw_a, sc.sh)]

    def fused():
        m, q, s, a, sh = lv
        for t in lv:
            t.grad = None
        render_view(m, q, s, a, c2w, cam, sh=sh, C=4, raw_params=True)["rgb"].backward(gradient=w)

    def torch_act():
        m, q, s, a, sh = lv
        for t in lv:
            t.grad = None
        render_view(m, q, torch.exp(s), torch.sigmoid(a), c2w, cam, sh=sh, C=4)["rgb"].backward(gradient=w)

    lv = leaves()
    mf, bf = timeit(fused, reps=8, warm=3)
    mt, bt = timeit(torch_act, reps=8, warm=3)
    res["view_c3_raw_leaves"] = {"ms_median_in_kernel_activations": mf, "ms_median_torch_activations": mt,
                                 "ms_min_in_kernel": bf, "ms_min_torch": bt}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
