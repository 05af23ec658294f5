"""Generate the densify / prune golden fixture from the UNMODIFIED reference methods (TEST INFRASTRUCTURE).

Runs in the dev container only (needs /root/reference).  `gs/gaussian_splatting.py` cannot be imported here (kornia,
torchtyping, omegaconf, matplotlib ... are absent), so the script parses the file with `ast`, compiles the function
definitions of the densify / prune / optimizer-surgery methods AS THEY ARE into a scratch namespace and binds them to
a minimal host object that carries the attributes they touch.  Nothing of the reference's source is written anywhere;
only the resulting tensors are stored in `tests/golden/densify_official.npz`.

Stubs (everything else is the reference's own code):
  * `torch` is proxied so that the hard-coded `device="cuda"` allocations land on the CPU and `torch.randn` returns
    recorded noise (stored in the fixture);
  * `qvec2rotmat_batched` (kornia 0.6.0, absent) = the restatement in oracle/__init__.py -- quaternions in the fixture
    are unit length, where the convention is unambiguous;
  * `C(value, step, _)` (utils/misc.py schedule helper) = identity on plain numbers.
"""
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/gs/gaussian_splatting.py"
METHODS = ["prune_optimizer", "densify_on_optimizer", "densify_with_new_params", "prune_by_mask", "densify_by_split",
           "densify_by_clone", "get_params_by_mask", "update_params_with_dict", "reset_densify_info",
           "update_densify_info", "prune_by_scale", "prune_by_alpha", "prune_by_svec", "densify", "prune",
           "densify_by_scale", "densify_by_all"]


class TorchProxy:
    """`torch` as the extracted methods see it: CPU allocations, recorded noise."""

    def __init__(self, gen, noise_log):
        self._gen, self._log = gen, noise_log

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def _cpu(kw):
        kw = dict(kw)
        if "device" in kw:
            kw["device"] = "cpu"
        return kw

    def zeros(self, *a, **kw):
        return torch.zeros(*a, **self._cpu(kw))

    def randn(self, *shape, **kw):
        t = torch.randn(*shape, generator=self._gen)
        self._log.append(t.clone())
        return t


def load_methods(namespace):
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GaussianSplattingRenderer")
    found = {}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in METHODS:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, REF, "exec"), namespace)
            found[node.name] = namespace[node.name]
    missing = set(METHODS) - set(found)
    assert not missing, missing
    return found


class Host:
    """the attributes of GaussianSplattingRenderer the methods above read and write"""
    fields = ["mean", "qvec", "svec", "color", "alpha"]
    pbr = False
    device = "cpu"

    svec = property(lambda self: torch.exp(self.svec_before_activation))
    alpha = property(lambda self: torch.sigmoid(self.alpha_before_activation))
    color = property(lambda self: torch.sigmoid(self.color_before_activation))
    svec_inv_act = staticmethod(torch.log)


def snapshot(h, tag, out):
    for f, raw in (("mean", "mean"), ("qvec", "qvec"), ("svec", "svec_before_activation"),
                   ("color", "color_before_activation"), ("alpha", "alpha_before_activation")):
        p = getattr(h, raw)
        out[f"{tag}_{f}"] = p.detach().clone()
        grp = next(g for g in h.optimizer.param_groups if g["name"] == f)
        st = h.optimizer.state[grp["params"][0]]
        out[f"{tag}_{f}_exp_avg"] = st["exp_avg"].clone()
        out[f"{tag}_{f}_exp_avg_sq"] = st["exp_avg_sq"].clone()
    for s in ("max_radii2d", "mean_2d_grad_accum", "cnt"):
        out[f"{tag}_{s}"] = getattr(h, s).clone().float()
    out[f"{tag}_N"] = torch.tensor([h.N])


def step_check(step, step_size, run_at_zero=False):  # gs/renderer.py:27-31 (checked in tests/test_host_mirrors_cpu.py)
    if step_size == 0:
        return False
    return (run_at_zero or step != 0) and step % step_size == 0


class Cfg(dict):
    """OmegaConf stand-in: attribute and .get access"""
    __getattr__ = dict.__getitem__


def main():
    import oracle

    g = torch.Generator().manual_seed(77)
    noise_log = []
    ns = {"torch": TorchProxy(g, noise_log), "nn": nn, "qvec2rotmat_batched": oracle.quat_to_rotmat,
          "C": lambda v, step, _=None: v, "step_check": step_check,
          "console": type("Con", (), {"print": staticmethod(lambda *a, **k: None)})(),
          "field2raw": dict(mean="mean", qvec="qvec", svec="svec_before_activation", color="color_before_activation",
                            alpha="alpha_before_activation")}
    for name, fn in load_methods(ns).items():
        setattr(Host, name, fn)
    h = Host()
    N = 400
    h.N = N
    h.mean = nn.Parameter(0.5 * torch.randn(N, 3, generator=g))
    h.qvec = nn.Parameter(torch.nn.functional.normalize(torch.randn(N, 4, generator=g), dim=-1))
    h.svec_before_activation = nn.Parameter(torch.log(0.004 + 0.03 * torch.rand(N, 3, generator=g)))
    h.color_before_activation = nn.Parameter(torch.randn(N, 3, generator=g))
    h.alpha_before_activation = nn.Parameter(2.0 * torch.randn(N, generator=g))
    h.densify_cfg = types.SimpleNamespace(split_thresh=0.02, split_shrink=0.8, n_splits=2, mean2d_thresh=0.02)
    h.prune_cfg = types.SimpleNamespace(radii2d_thresh=1.0, alpha_thresh=0.05, radii3d_thresh=0.012)
    lr = dict(mean=0.005, qvec=0.003, svec=0.003, color=0.01, alpha=0.003)
    h.optimizer = torch.optim.Adam(
        [{"params": [getattr(h, ns["field2raw"][f])], "lr": lr[f], "name": f} for f in Host.fields], lr=0.0, eps=1e-15)
    out = {}
    # three Adam steps so that the moments are populated
    for _ in range(3):
        for f in Host.fields:
            p = getattr(h, ns["field2raw"][f])
            p.grad = torch.randn(p.shape, generator=g) * 0.01
        h.optimizer.step()
    h.reset_densify_info()
    h.mean_2d_grad_accum = torch.rand(N, generator=g) * 0.1
    h.cnt = torch.randint(0, 4, (N,), generator=g).float()  # zeros -> NaN -> 0 in densify()
    h.mean_2d_grad_accum[h.cnt == 0] = 0.0
    h.max_radii2d = torch.rand(N, generator=g) * 1.5
    snapshot(h, "s0", out)
    # densify(), type "official" (gs/gaussian_splatting.py:770-778): the four statements of that branch
    grads = h.mean_2d_grad_accum / h.cnt
    grads[grads.isnan()] = 0.0
    out["s0_grads"] = grads.clone()
    n_clone = h.densify_by_clone(grads, h.densify_cfg.mean2d_thresh)
    snapshot(h, "s1", out)
    n_split = h.densify_by_split(grads, h.densify_cfg.mean2d_thresh, h.densify_cfg.n_splits)
    snapshot(h, "s2", out)
    h.reset_densify_info()
    out["s2_counts"] = torch.tensor([n_clone, n_split])
    out["noise"] = torch.cat(noise_log) if noise_log else torch.zeros(0, 3)
    # prune(): by screen radius, opacity, 3-D scale (:1152-1176)
    h.max_radii2d = torch.rand(h.N, generator=g) * 1.2
    out["s2_max_radii2d_for_prune"] = h.max_radii2d.clone()
    n1 = h.prune_by_scale(0)
    snapshot(h, "s3", out)
    n2 = h.prune_by_alpha(0)
    snapshot(h, "s4", out)
    n3 = h.prune_by_svec(0)
    snapshot(h, "s5", out)
    out["prune_counts"] = torch.tensor([n1, n2, n3])
    # ---- the step-gated dispatchers densify(step) / prune(step) (:751-817, :1152-1176) on a fresh copy of s0
    h2 = Host()
    h2.N = N
    for f, raw in ns["field2raw"].items():
        setattr(h2, raw, nn.Parameter(out[f"s0_{f}"].clone()))
    h2.optimizer = torch.optim.Adam(
        [{"params": [getattr(h2, ns["field2raw"][f])], "lr": lr[f], "name": f} for f in Host.fields], lr=0.0, eps=1e-15)
    for f in Host.fields:  # same moments as s0
        p_ = getattr(h2, ns["field2raw"][f])
        h2.optimizer.state[p_] = {"step": torch.tensor(3.0), "exp_avg": out[f"s0_{f}_exp_avg"].clone(),
                                  "exp_avg_sq": out[f"s0_{f}_exp_avg_sq"].clone()}
    h2.densify_cfg = Cfg(enabled=True, type="official", warm_up=2000, end=4999, period=500, mean2d_thresh=0.02,
                         split_thresh=0.02, n_splits=2, split_shrink=0.8, use_legacy=False)
    h2.prune_cfg = Cfg(enabled=True, warm_up=0, end=15000, period=100, radii2d_thresh=1.0, alpha_thresh=0.05,
                       radii3d_thresh=0.012)
    h2.cfg = Cfg(densify=h2.densify_cfg)
    trace = []
    g2 = torch.Generator().manual_seed(78)
    noise2 = []
    ns["torch"]._gen, ns["torch"]._log = g2, noise2
    for step in (0, 1999, 2000, 2100, 2500, 5000):
        h2.mean_2d_grad_accum = out["s0_mean_2d_grad_accum"].clone()[: h2.N] if h2.N <= N else torch.rand(h2.N, generator=g) * 0.1
        h2.cnt = out["s0_cnt"].clone()[: h2.N] if h2.N <= N else torch.ones(h2.N)
        h2.max_radii2d = torch.linspace(0.0, 1.3, h2.N)
        n_before = h2.N
        h2.densify(step, verbose=False)
        n_mid = h2.N
        h2.max_radii2d = torch.linspace(0.0, 1.3, h2.N)
        h2.prune(step, verbose=False)
        trace.append([step, n_before, n_mid, h2.N])
    out["dispatch_trace"] = torch.tensor(trace)
    out["dispatch_noise"] = torch.cat(noise2) if noise2 else torch.zeros(0, 3)
    snapshot(h2, "s9", out)
    print("dispatch trace (step, N before, N after densify, N after prune):", trace)
    path = os.path.join(ROOT, "tests", "golden", "densify_official.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print("N0", N, "clone", n_clone, "split", n_split, "-> N", int(out["s2_N"]), "prune", n1, n2, n3, "-> N", h.N,
          "noise rows", out["noise"].shape[0], "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
