"""Generate the compactness-densification / neighbour-penalty golden fixture from the UNMODIFIED reference code
(TEST INFRASTRUCTURE; runs in the dev container only, needs /root/reference).

Executed as they are (parsed with `ast`, compiled into a scratch namespace, nothing of the source is written anywhere):
  gs/gaussian_splatting.py  densify_by_compatnes_with_idx, densify_by_compatness, densify_by_shrink_then_compatness,
                            densify_with_new_params, densify_on_optimizer, update_params_with_dict, the `svec` setter,
                            compat_penalty_loss, NN_penalty_loss, densify (the step-gated dispatcher)
  utils/ops.py              distance_to_gaussian_surface, K_nearest_neighbors, nearest_neighbor

Stubs: the ones of make_densify_golden.py (CPU `torch` proxy, `qvec2rotmat_batched` = oracle.quat_to_rotmat for unit
quaternions, `C` = identity) plus `knn_points`: pytorch3d is absent, so `oracle.knn_points` (brute force, ties by index)
stands in behind pytorch3d's batched signature `knn_points(p1[1,Q,3], p2[1,N,3], K, return_nn=True)`.
Result: tests/golden/compatness.npz
"""
import ast
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
import make_densify_golden as mdg  # noqa: E402

REF_GS = "/root/reference/gs/gaussian_splatting.py"
REF_OPS = "/root/reference/utils/ops.py"
METHODS = ["densify_by_compatnes_with_idx", "densify_by_compatness", "densify_by_shrink_then_compatness",
           "densify_with_new_params", "densify_on_optimizer", "update_params_with_dict", "compat_penalty_loss",
           "NN_penalty_loss", "densify", "reset_densify_info"]
OPS = ["distance_to_gaussian_surface", "K_nearest_neighbors", "nearest_neighbor"]


def load(namespace):
    tree = ast.parse(open(REF_OPS).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in OPS:
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF_OPS, "exec"), namespace)  # decorators kept
    assert all(n in namespace for n in OPS)
    tree = ast.parse(open(REF_GS).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GaussianSplattingRenderer")
    found, svec_setter = {}, None
    for node in cls.body:
        if not isinstance(node, ast.FunctionDef):
            continue
        decos = [ast.unparse(d) for d in node.decorator_list]
        if node.name == "svec" and "svec.setter" in decos:
            node.decorator_list = []
            node.name = "_svec_setter"
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF_GS, "exec"), namespace)
            svec_setter = namespace["_svec_setter"]
        elif node.name in METHODS and "property" not in decos:
            node.decorator_list = []
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF_GS, "exec"), namespace)
            found[node.name] = namespace[node.name]
    assert not set(METHODS) - set(found), set(METHODS) - set(found)
    assert svec_setter is not None
    return found, svec_setter


class Writer:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, step):
        self.scalars[tag] = float(value)


def make_host(Host, ns, state, lr):
    h = Host()
    h.N = state["mean"].shape[0]
    for f, raw in ns["field2raw"].items():
        setattr(h, raw, nn.Parameter(state[f].clone()))
    h.optimizer = torch.optim.Adam(
        [{"params": [getattr(h, ns["field2raw"][f])], "lr": lr[f], "name": f} for f in Host.fields], lr=0.0, eps=1e-15)
    g = torch.Generator().manual_seed(5)
    for _ in range(2):  # populated Adam moments
        for f in Host.fields:
            p = getattr(h, ns["field2raw"][f])
            p.grad = torch.randn(p.shape, generator=g) * 0.01
        h.optimizer.step()
    for f in Host.fields:  # the gradients recorded below are those of the loss alone
        getattr(h, ns["field2raw"][f]).grad = None
    h.reset_densify_info()
    return h


def main():
    import oracle

    def knn_points(p1, p2, K, return_nn=True):  # pytorch3d's batched signature over the brute-force search
        d2, idx = oracle.knn_points(p1[0], p2[0], K)
        return d2[None], idx[None], p2[0][idx][None]

    g = torch.Generator().manual_seed(123)
    ns = {"torch": mdg.TorchProxy(g, []), "nn": nn, "F": F, "Optional": Optional,
          "qvec2rotmat_batched": oracle.quat_to_rotmat, "C": lambda v, step, _=None: v, "step_check": mdg.step_check,
          "console": type("Con", (), {"print": staticmethod(lambda *a, **k: None)})(), "knn_points": knn_points,
          "pytorch3d_capable": True,
          "field2raw": dict(mean="mean", qvec="qvec", svec="svec_before_activation", color="color_before_activation",
                            alpha="alpha_before_activation")}
    methods, svec_setter = load(ns)

    class Host(mdg.Host):
        rotmat = property(lambda self: oracle.quat_to_rotmat(self.qvec))
        svec = property(lambda self: torch.exp(self.svec_before_activation), svec_setter)

    for name, fn in methods.items():
        setattr(Host, name, fn)
    N = 500
    # a cloud with gaps: points on a jittered lattice, scales well below the spacing for some, above it for others
    lattice = torch.stack(torch.meshgrid(*[torch.arange(8.0)] * 3, indexing="ij"), -1).reshape(-1, 3)[:N] * 0.25
    state = {"mean": lattice + 0.05 * torch.randn(N, 3, generator=g),
             "qvec": F.normalize(torch.randn(N, 4, generator=g), dim=-1),
             "svec": torch.log(0.02 + 0.22 * torch.rand(N, 3, generator=g)),
             "color": torch.randn(N, 3, generator=g), "alpha": 2.0 * torch.randn(N, generator=g)}
    lr = dict(mean=0.005, qvec=0.003, svec=0.003, color=0.01, alpha=0.003)
    out = {f"in_{k}": v.clone() for k, v in state.items()}

    # ---- utils/ops.py on their own
    h = make_host(Host, ns, state, lr)
    snap0 = {}
    mdg.snapshot(h, "s0", snap0)
    out.update(snap0)
    nn_pos, idx4 = ns["K_nearest_neighbors"](h.mean, K=4)
    out["knn_idx_K4"] = idx4.clone()
    out["knn_nn_K4"] = nn_pos.clone()
    nn1, idx1 = ns["nearest_neighbor"](h.mean)
    out["nn_idx"], out["nn_pos"] = idx1.clone(), nn1.clone()
    out["surface_self_to_nn"] = ns["distance_to_gaussian_surface"](h.mean, h.svec, h.rotmat, nn1).detach().clone()
    # ---- densify_by_compatnes_with_idx for the nearest neighbour
    new = h.densify_by_compatnes_with_idx(idx4[:, 0])
    for k, v in new.items():
        out[f"with_idx0_{k}"] = v.detach().clone()
    # ---- densify_by_compatness(K=3) incl. the optimizer surgery
    n_new = h.densify_by_compatness(K=3)
    mdg.snapshot(h, "s1", out)
    out["s1_num"] = torch.tensor([n_new])
    # ---- densify_by_shrink_then_compatness(1.5, K=2) on a fresh copy
    h2 = make_host(Host, ns, state, lr)
    n_new2 = h2.densify_by_shrink_then_compatness(1.5, K=2)
    mdg.snapshot(h2, "s2", out)
    out["s2_num"] = torch.tensor([n_new2])
    # ---- the dispatcher: densify(step) with type "compatness" / "shrink_then_compatness" (:790-810)
    trace = []
    for kind in ("compatness", "shrink_then_compatness"):
        h3 = make_host(Host, ns, state, lr)
        h3.densify_cfg = mdg.Cfg(enabled=True, type=kind, warm_up=100, end=1000, period=100, use_legacy=False, K=2,
                                 surface_shrink=1.25)
        h3.cfg = mdg.Cfg(densify=h3.densify_cfg)
        for step in (0, 99, 100, 150, 200, 1100):
            h3.mean_2d_grad_accum = torch.ones(h3.N)
            h3.cnt = torch.ones(h3.N)
            n_before = h3.N
            h3.densify(step, verbose=False)
            trace.append([step, n_before, h3.N])
        mdg.snapshot(h3, f"s3_{kind}", out)
    out["dispatch_trace"] = torch.tensor(trace)
    # ---- penalties: value + gradients w.r.t. the raw leaves
    for kind in ("l1", "l2"):
        h4 = make_host(Host, ns, state, lr)
        h4.cfg = mdg.Cfg(penalty=mdg.Cfg(compat=mdg.Cfg(value=0.7, type=kind), NN=mdg.Cfg(value=0.3)))
        w = Writer()
        loss = h4.compat_penalty_loss(10, w)
        loss.backward()
        out[f"compat_{kind}_loss"] = loss.detach().reshape(1).clone()
        out[f"compat_{kind}_effective_rate"] = torch.tensor([w.scalars["auxiliary/effective_rate"]])
        out[f"compat_{kind}_g_mean"] = h4.mean.grad.clone()
        out[f"compat_{kind}_g_svec"] = h4.svec_before_activation.grad.clone()
        out[f"compat_{kind}_g_qvec"] = h4.qvec.grad.clone()
    h5 = make_host(Host, ns, state, lr)
    h5.cfg = mdg.Cfg(penalty=mdg.Cfg(NN=mdg.Cfg(value=0.3)))
    loss = h5.NN_penalty_loss(10, Writer())
    loss.backward()
    out["NN_loss"] = loss.detach().reshape(1).clone()
    out["NN_g_mean"] = h5.mean.grad.clone()

    path = os.path.join(ROOT, "tests", "golden", "compatness.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print("N", N, "compatness K=3 ->", n_new, "new; shrink_then K=2 ->", n_new2, "new; dispatch", trace,
          "| compat l1", float(out["compat_l1_loss"]), "rate", float(out["compat_l1_effective_rate"]),
          "NN", float(out["NN_loss"]), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
