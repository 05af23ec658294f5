"""K nearest neighbours of the Gaussian means -- the host-side mirror of `utils/ops.py:103-134` over `gsb200_knn`
(csrc/knn.cu: uniform-grid shell search, exact) and the geometry helper the compactness rules use.

Reference interface kept (same names, argument meaning and return values):
  * `nearest_neighbor(mean)`                     utils/ops.py:103-114  -> (position of the nearest OTHER point [N,3], its index [N])
  * `K_nearest_neighbors(mean, K, query, return_dist)`  :117-134      -> columns 1..K-1 of knn_points(query, mean, K): the
    reference always drops column 0 (the query itself when `query is None`), so `K` neighbours need `K + 1`
    (gs/gaussian_splatting.py:683 passes `K=K + 1`).
  * `distance_to_gaussian_surface(mean, svec, rotmat, query)`  :137-158

Both searches sit on `pytorch3d.ops.knn_points` in the reference (third party, optional import, unpinned, absent in this
image): squared Euclidean distances ascending, int64 indices.  `knn_points` below has that contract; ties are ordered by
point index.  CUDA tensors only -- there is no CPU fallback (the CPU test-suite plugs the oracle's brute-force search
into `GaussianStore(knn_fn=...)`, like `render_fn` for the renderer)."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib

MAX_K = 32


def knn_points(query: Optional[torch.Tensor], points: torch.Tensor, K: int, return_dist: bool = True):
    """(dist2 [Q,K] fp32 or None, idx [Q,K] int64): the K points nearest to each query, ascending; `query=None` = the
    points themselves (column 0 is then the point itself).  Slots beyond the number of points: idx -1, dist2 +inf."""
    if not points.is_cuda:
        raise RuntimeError("gsgen_b200.knn needs CUDA tensors (libgsb200 has no CPU path)")
    if not 1 <= int(K) <= MAX_K:
        raise RuntimeError(f"K = {K} outside [1, {MAX_K}]")
    dev = points.device
    pts = points.detach().to(torch.float32).contiguous()
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise RuntimeError(f"points must be [N,3], got {tuple(pts.shape)}")
    q = None
    if query is not None:
        q = query.detach().to(dev, torch.float32).contiguous()
        if q.dim() != 2 or q.shape[1] != 3:
            raise RuntimeError(f"query must be [Q,3], got {tuple(q.shape)}")
    nq = pts.shape[0] if q is None else q.shape[0]
    idx = torch.empty(nq, int(K), dtype=torch.int64, device=dev)
    d2 = torch.empty(nq, int(K), dtype=torch.float32, device=dev) if return_dist else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().gsb200_knn(
            _lib.ctx(dev), _lib.fptr(pts, "points"), ctypes.c_uint32(pts.shape[0]), _lib.fptr(q, "query"),
            ctypes.c_uint32(nq), ctypes.c_int32(int(K)), _lib.ptr(idx, torch.int64, "idx"), _lib.fptr(d2, "dist2"),
            _lib.stream_ptr(dev)))
    return d2, idx


@torch.no_grad()
def K_nearest_neighbors(mean: torch.Tensor, K: int, query: Optional[torch.Tensor] = None, return_dist: bool = False,
                        knn=knn_points):
    """utils/ops.py:117-134: `knn_points(query, mean, K, return_nn=True)` with column 0 dropped ->
    (neighbour positions [Q,K-1,3], indices [Q,K-1] (, squared distances [Q,K-1]))."""
    d2, idx = knn(query, mean, K, return_dist)
    idx = idx[:, 1:]
    nn = mean.detach()[idx.clamp_min(0)]
    if not return_dist:
        return nn, idx
    return nn, idx, d2[:, 1:]


@torch.no_grad()
def nearest_neighbor(mean: torch.Tensor, knn=knn_points):
    """utils/ops.py:103-114: `knn_points(mean, mean, K=2)`, column 1 (column 0 is the point itself)."""
    _, idx = knn(None, mean, 2, False)
    idx = idx[:, 1]
    return mean.detach()[idx.clamp_min(0)], idx


def distance_to_gaussian_surface(mean, svec, rotmat, query):
    """utils/ops.py:137-158, expression for expression: the "radius" of the ellipsoid (axes svec, frame rotmat) in the
    direction of `query`, from the spherical angles of that direction in the Gaussian's frame -- including the
    reference's `d2**2 * sin_theta**2` (d2 is already a squared length there) and its 1e-10 guards."""
    xyz = query - mean
    xyz = torch.einsum("bij,bj->bi", rotmat.transpose(-1, -2), xyz)
    xyz = F.normalize(xyz, dim=-1)
    z, y, x = xyz[..., 2], xyz[..., 1], xyz[..., 0]
    r_xy = torch.sqrt(x ** 2 + y ** 2 + 1e-10)
    cos_theta, sin_theta = z, r_xy
    cos_phi, sin_phi = x / r_xy, y / r_xy
    d2 = svec[..., 0] ** 2 * cos_phi ** 2 + svec[..., 1] ** 2 * sin_phi ** 2
    r2 = svec[..., 2] ** 2 * cos_theta ** 2 + d2 ** 2 * sin_theta ** 2
    return torch.sqrt(r2 + 1e-10)
