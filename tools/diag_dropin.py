"""Diagnostic for the drop-in difference found at c3 / 30000 Gaussians / 320^2 (svec x2): run the reference's own
render_one over `_gs.so` and over libgsb200 (optionally a variant library via GSB200_LIB) and save inputs + outputs.
usage: python tools/diag_dropin.py <tag>"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gsgen_b200.scenes import make_scene  # noqa: E402
from tests import refpy  # noqa: E402

tag = sys.argv[1]
DEV = "cuda"
entries = refpy.load_entries()
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import _gs  # noqa: E402
from gsgen_b200.backend import _backend as ours  # noqa: E402

rec_a, rec_b = refpy.Recording(_gs), refpy.Recording(ours)
tm = refpy.TorchNoProfiler()
ns_a, ns_b = refpy.namespace(rec_a, entries, tm), refpy.namespace(rec_b, entries, tm)
sc = make_scene("c3", N=30000, reso=320)
sc.svec = (sc.svec * 2.0).contiguous()
cam, c2w = sc.cams[0], sc.c2ws[0]
H, W = cam.h, cam.w
g = torch.Generator().manual_seed(31)
bg = torch.rand(H, W, 3, generator=g)
weights = {k: torch.randn(H, W, 3 if k == "rgb" else 1, generator=g) for k in ("rgb", "depth", "opacity", "z_var")}
out_a, _, side_a = refpy.run_render_one(ns_a, sc, cam, c2w, DEV, bg, weights)
out_b, _, side_b = refpy.run_render_one(ns_b, sc, cam, c2w, DEV, bg, weights)
out_b2, _, _ = refpy.run_render_one(ns_b, sc, cam, c2w, DEV, bg, weights)
ca, cb = rec_a.calls["tile_based_vol_rendering_start_end_with_T"], rec_b.calls["tile_based_vol_rendering_start_end_with_T"]
d = (out_a["rgb"] - out_b["rgb"]).abs().amax(-1)
print(tag, "lib", os.environ.get("GSB200_LIB", "default"), "max diff", float(d.max()), "pixels > 1e-4:", int((d > 1e-4).sum()),
      "run-to-run ours max diff", float((out_b["rgb"] - out_b2["rgb"]).abs().max()),
      "opacity diff", float((out_a["opacity"] - out_b["opacity"]).abs().max()))
# the compat op alone on the reference arm's tensors with BOTH id lists (ours / reference's)
mean, cov, col, al, st, en, ids_a = ca[:7]
ids_b = cb[6]
topleft = ca[8]
common = ca[9:17]
res = {}
for name, be, ids in (("ref_refids", _gs, ids_a), ("ours_refids", ours, ids_a), ("ours_ourids", ours, ids_b), ("ref_ourids", _gs, ids_b)):
    o = torch.zeros(H, W, 3, device=DEV)
    T = torch.ones(H, W, 1, device=DEV)
    be.tile_based_vol_rendering_start_end_with_T(mean, cov, col, al, st, en, ids, o, topleft, *common, T)
    torch.cuda.synchronize()
    res[name] = (o.cpu().numpy(), T.cpu().numpy())
for a_, b_ in (("ref_refids", "ours_refids"), ("ref_refids", "ref_ourids"), ("ours_refids", "ours_ourids"), ("ref_ourids", "ours_ourids")):
    dd = np.abs(res[a_][0] - res[b_][0]).max(-1)
    print("  ", a_, "vs", b_, "max", float(dd.max()), "pixels > 1e-4:", int((dd > 1e-4).sum()))
neq = (ids_a != ids_b)
print("   ids differ at", int(neq.sum()), "positions of", ids_a.numel())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"diag_{tag}.npz"),
                    mean2d=mean.detach().cpu().numpy(), cov2d=cov.detach().cpu().numpy(), color=col.detach().cpu().numpy(),
                    alpha=al.detach().cpu().numpy(), start=st.cpu().numpy(), end=en.cpu().numpy(), ids_ref=ids_a.cpu().numpy(),
                    ids_ours=ids_b.cpu().numpy(), topleft=topleft.cpu().numpy(), depth=rec_a.calls["tile_culling_aabb_start_end"][5].detach().cpu().numpy(),
                    **{f"rgb_{k}": v[0] for k, v in res.items()}, **{f"T_{k}": v[1] for k, v in res.items()})
