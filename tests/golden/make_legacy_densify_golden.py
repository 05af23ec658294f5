"""Golden fixture for `densify(step)` with `use_legacy: True` -- what the reference's TOP-LEVEL experiment configs use
(conf/base.yaml, conf/corgi.yaml, conf/shrink_then_densify.yaml: `renderer.densify.use_legacy: True`) -- produced by
executing the UNMODIFIED reference methods (TEST INFRASTRUCTURE; dev container only, needs /root/reference):

  gs/gaussian_splatting.py  densify (:751-817), densify_legacy (:820-946), densify_by_shrink_then_compatness,
                            densify_by_compatness, densify_by_compatnes_with_idx, densify_with_new_params,
                            densify_on_optimizer, update_params_with_dict, reset_densify_info, the `svec` setter
  utils/ops.py              distance_to_gaussian_surface, K_nearest_neighbors

Stubs: those of make_densify_golden.py / make_compatness_golden.py (CPU torch proxy with recorded noise, kornia
quaternion matrix for unit quaternions, brute-force `knn_points`), and `set_optimizer(cfg, step)` -- densify_legacy
re-creates the optimizer through it (:938): the stand-in builds the fresh `torch.optim.Adam` the reference's own
set_optimizer would (no state) and records the call.
Result: tests/golden/densify_legacy.npz
"""
import os
import sys
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_compatness_golden as mcg  # noqa: E402
import make_densify_golden as mdg  # noqa: E402

LR = dict(mean=0.005, qvec=0.003, svec=0.003, color=0.01, alpha=0.003)


def snapshot(h, tag, out):
    """parameters + Adam state as it is: a freshly re-created optimizer has NO state (recorded as zeros + has_state 0)"""
    has = []
    for f, raw in (("mean", "mean"), ("qvec", "qvec"), ("svec", "svec_before_activation"),
                   ("color", "color_before_activation"), ("alpha", "alpha_before_activation")):
        p = getattr(h, raw)
        out[f"{tag}_{f}"] = p.detach().clone()
        grp = next(g for g in h.optimizer.param_groups if g["name"] == f)
        assert grp["params"][0] is p, "the optimizer must hold the live parameter"
        st = h.optimizer.state.get(grp["params"][0], {})
        has.append(int("exp_avg" in st))
        out[f"{tag}_{f}_exp_avg"] = st["exp_avg"].clone() if "exp_avg" in st else torch.zeros_like(p)
        out[f"{tag}_{f}_exp_avg_sq"] = st["exp_avg_sq"].clone() if "exp_avg_sq" in st else torch.zeros_like(p)
    out[f"{tag}_has_state"] = torch.tensor(has)
    out[f"{tag}_N"] = torch.tensor([h.N])
    for s_ in ("max_radii2d", "mean_2d_grad_accum", "cnt"):
        out[f"{tag}_{s_}"] = getattr(h, s_).clone().float()


def main():
    import oracle

    def knn_points(p1, p2, K, return_nn=True):
        d2, idx = oracle.knn_points(p1[0], p2[0], K)
        return d2[None], idx[None], p2[0][idx][None]

    g = torch.Generator().manual_seed(321)
    noise_log = []
    ns = {"torch": mdg.TorchProxy(g, noise_log), "nn": nn, "F": F, "Optional": Optional,
          "qvec2rotmat_batched": oracle.quat_to_rotmat, "C": lambda v, step, _=None: v, "step_check": mdg.step_check,
          "console": type("Con", (), {"print": staticmethod(lambda *a, **k: None)})(), "knn_points": knn_points,
          "pytorch3d_capable": True,
          "field2raw": dict(mean="mean", qvec="qvec", svec="svec_before_activation", color="color_before_activation",
                            alpha="alpha_before_activation")}
    mcg.METHODS = list(mcg.METHODS) + ["densify_legacy"]
    methods, svec_setter = mcg.load(ns)

    calls = []

    class Host(mdg.Host):
        rotmat = property(lambda self: oracle.quat_to_rotmat(self.qvec))
        svec = property(lambda self: torch.exp(self.svec_before_activation), svec_setter)

        def set_optimizer(self, cfg, step=0):
            calls.append(int(step))
            self.optimizer = torch.optim.Adam(
                [{"params": [getattr(self, ns["field2raw"][f])], "lr": LR[f], "name": f} for f in Host.fields],
                lr=0.0, eps=1e-15)

    for name, fn in methods.items():
        setattr(Host, name, fn)
    N = 400
    lattice = torch.stack(torch.meshgrid(*[torch.arange(8.0)] * 3, indexing="ij"), -1).reshape(-1, 3)[:N] * 0.25
    state = {"mean": lattice + 0.05 * torch.randn(N, 3, generator=g),
             "qvec": F.normalize(torch.randn(N, 4, generator=g), dim=-1),
             "svec": torch.log(0.005 + 0.05 * torch.rand(N, 3, generator=g)),
             "color": torch.randn(N, 3, generator=g), "alpha": 2.0 * torch.randn(N, generator=g)}
    accum = torch.rand(N, generator=g) * 0.1
    cnt = torch.randint(0, 4, (N,), generator=g).float()
    out = {f"in_{k}": v.clone() for k, v in state.items()}
    for kind in ("official", "shrink_then_compatness", "compatness"):
        h = mcg.make_host(Host, ns, state, LR)
        if kind == "official":
            mdg.snapshot(h, "s0", out)
        for f in Host.fields:  # densify_legacy asserts that the backward has run (:822-824)
            getattr(h, ns["field2raw"][f]).grad = torch.zeros_like(getattr(h, ns["field2raw"][f]))
        h.opt_cfg = mdg.Cfg(type="Adam", opt_args=mdg.Cfg(eps=1e-15))
        h.densify_cfg = mdg.Cfg(enabled=True, type=kind, warm_up=100, end=1000, period=100, use_legacy=True, K=2,
                                surface_shrink=1.25, mean2d_thresh=0.02, split_thresh=0.03, n_splits=2,
                                split_shrink=0.8)
        h.cfg = mdg.Cfg(densify=h.densify_cfg)
        noise_log.clear()
        calls.clear()
        trace = []
        for step in (0, 100, 150):
            h.mean_2d_grad_accum, h.cnt = accum[: h.N].clone() if h.N == N else torch.zeros(h.N), \
                cnt[: h.N].clone() if h.N == N else torch.zeros(h.N)
            n_before = h.N
            h.densify(step, verbose=False)
            trace.append([step, n_before, h.N])
        snapshot(h, f"s1_{kind}", out)
        out[f"trace_{kind}"] = torch.tensor(trace)
        out[f"noise_{kind}"] = torch.cat(noise_log) if noise_log else torch.zeros(0, 3)
        out[f"set_optimizer_calls_{kind}"] = torch.tensor(calls)
        print(kind, "trace", trace, "set_optimizer calls at steps", calls, "noise rows", out[f"noise_{kind}"].shape[0])
    out["accum"], out["cnt"] = accum, cnt
    path = os.path.join(ROOT, "tests", "golden", "densify_legacy.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print("wrote", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
