"""The per-Gaussian stage -- frustum planes, projection + its autograd, AABB tile rectangles and the duplicate count --
against vectors produced by the reference's OWN torch functions (tests/golden/make_pergaussian_golden.py compiles
`CameraInfo`, `project_gaussians`, `jacobian`, `qsvec2rotmat_batched`, `tile_culling_aabb_count` unmodified out of the
reference files).  Checked here: the oracle's restatement, the product's host-side camera code, and the host build of
the math header the CUDA kernels use (projection forward / backward, bit-exact AABBs)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from gsgen_b200.camera import CameraInfo
from tests.util import fp, ip

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pergaussian_ref.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _cams(gold, tag, oracle_mod):
    fx, fy, cx, cy, w, h, near, far = gold[f"{tag}_cam"].tolist()
    return (oracle_mod.Cam(fx, fy, cx, cy, int(w), int(h), near, far),
            CameraInfo(fx, fy, cx, cy, int(w), int(h), near, far))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_frustum_and_rays(gold, oracle_mod, tag):
    ocam, pcam = _cams(gold, tag, oracle_mod)
    c2w = gold[f"{tag}_c2w"]
    for who in (lambda: oracle_mod.get_frustum(ocam, c2w), lambda: pcam.get_frustum(c2w)):
        normals, pts = who()
        assert torch.allclose(normals, gold[f"{tag}_frustum_normals"], rtol=1e-6, atol=1e-7)
        assert torch.allclose(pts, gold[f"{tag}_frustum_pts"], rtol=1e-6, atol=1e-7)
    for rays in (oracle_mod.get_rays_d(ocam, c2w), pcam.get_rays_d(c2w)):
        assert torch.allclose(rays[::7, ::5], gold[f"{tag}_rays_d"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("mode", ["detach", "full"])
def test_oracle_projection_and_autograd(gold, oracle_mod, tag, mode):
    k = f"{tag}_{mode}"
    mean = gold[f"{tag}_mean"].clone().requires_grad_()
    qvec = gold[f"{tag}_qvec"].clone().requires_grad_()
    svec = gold[f"{tag}_svec"].clone().requires_grad_()
    m2, cov, JW, dp = oracle_mod.project_gaussians(mean, qvec, svec, gold[f"{tag}_c2w"], mode == "detach")
    assert torch.allclose(m2, gold[f"{k}_mean2d"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(cov, gold[f"{k}_cov2d"], rtol=1e-4, atol=1e-9)
    assert torch.allclose(dp, gold[f"{k}_depth"], rtol=1e-6, atol=1e-7)
    if mode == "detach":
        assert torch.allclose(JW, gold[f"{k}_JW"], rtol=1e-5, atol=1e-6)
    ((m2 * gold[f"{k}_g_mean2d"]).sum() + (cov * gold[f"{k}_g_cov2d"]).sum() + (dp * gold[f"{k}_g_depth"]).sum()).backward()
    for ours, name in ((mean.grad, "mean"), (qvec.grad, "qvec"), (svec.grad, "svec")):
        ref = gold[f"{k}_grad_{name}"]
        assert float((ours - ref).norm() / ref.norm()) < 1e-5, name


@pytest.mark.parametrize("tag", ["a", "b"])
def test_aabb_rectangles_and_duplicate_count_bit_exact(gold, oracle_mod, hostmath, tag):
    ocam, _ = _cams(gold, tag, oracle_mod)
    front = gold[f"{tag}_front"]
    m2 = gold[f"{tag}_detach_mean2d"][front].contiguous()
    cov = gold[f"{tag}_detach_cov2d"][front].contiguous()
    D, tl, br = oracle_mod.tile_culling_aabb_count(m2, cov, 16, ocam, 6.0)
    assert D == int(gold[f"{tag}_D"]) and torch.equal(tl, gold[f"{tag}_aabb_tl"].to(tl.dtype)) \
        and torch.equal(br, gold[f"{tag}_aabb_br"].to(br.dtype))
    # the kernels' integer rectangle (gsb200_math.cuh::aabb_tiles, host build)
    n = m2.shape[0]
    htl, hbr = torch.zeros(n, 2, dtype=torch.int32), torch.zeros(n, 2, dtype=torch.int32)
    hostmath.hm_aabb(n, fp(m2), fp(cov.reshape(n, 4).contiguous()), ctypes.c_float(6.0), ctypes.c_float(ocam.fx),
                     ctypes.c_float(ocam.fy), ctypes.c_float(ocam.cx), ctypes.c_float(ocam.cy), ocam.w, ocam.h, 16,
                     ip(htl), ip(hbr))
    assert torch.equal(htl.long(), gold[f"{tag}_aabb_tl"].long()) and torch.equal(hbr.long(), gold[f"{tag}_aabb_br"].long())
    assert int(((hbr - htl + 1).prod(dim=1)).sum()) == int(gold[f"{tag}_D"])


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("mode", ["detach", "full"])
def test_kernel_math_projection_forward_backward(gold, hostmath, tag, mode):
    """gsb200_math.cuh::project_gaussian / project_gaussian_bwd (the code k_preprocess / k_project_bwd_fused run),
    compiled for the host, against the reference functions' outputs and autograd gradients."""
    k = f"{tag}_{mode}"
    mean, qvec, svec = gold[f"{tag}_mean"], gold[f"{tag}_qvec"], gold[f"{tag}_svec"]
    c2w = gold[f"{tag}_c2w"][:3, :4].contiguous()
    n = mean.shape[0]
    m2, cov, dp = torch.empty(n, 2), torch.empty(n, 4), torch.empty(n)
    hostmath.hm_project_fwd(n, fp(mean), fp(qvec), fp(svec), fp(c2w), fp(m2), fp(cov), fp(dp))
    assert torch.allclose(m2, gold[f"{k}_mean2d"], rtol=2e-5, atol=1e-6)
    assert torch.allclose(dp, gold[f"{k}_depth"].reshape(-1), rtol=2e-6, atol=1e-6)
    ref_cov = gold[f"{k}_cov2d"].reshape(n, 4)
    rel = (cov - ref_cov).abs().max(dim=1).values / ref_cov.abs().max(dim=1).values
    assert float(rel.max()) < 5e-5
    gx, gq, gs = torch.empty(n, 3), torch.empty(n, 4), torch.empty(n, 3)
    hostmath.hm_project_bwd(n, fp(mean), fp(qvec), fp(svec), fp(c2w), 1 if mode == "detach" else 0,
                            fp(gold[f"{k}_g_mean2d"]), fp(gold[f"{k}_g_cov2d"].reshape(n, 4).contiguous()),
                            fp(gold[f"{k}_g_depth"].reshape(-1).contiguous()), fp(gx), fp(gq), fp(gs))
    for ours, name in ((gx, "mean"), (gq, "qvec"), (gs, "svec")):
        ref = gold[f"{k}_grad_{name}"]
        assert float((ours - ref).norm() / ref.norm()) < 1e-4, name
