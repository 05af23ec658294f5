"""Design study (CPU, numpy, no GPU): can the SH colour dot product -- the largest block of instructions in both
composite kernels (39 of 98 warp instructions per blended pair in the forward at degree 3) -- be replaced by a per-tile
SECOND-ORDER expansion of the basis in the pixel offset?

    Y_k(u, v) ~= sum_j Bm[k][j] phi_j(u, v),  phi = (1, u, v, u^2, uv, v^2)
    => s_c = sum_j E[c][j] phi_j with E = sh . Bm contracted once per (tile, Gaussian): 15 FMAs per pair instead of 48,
       and the backward would accumulate 3 x 6 moments instead of 3 x 16 sums.

Answer recorded in DESIGN.md (round-1 negative result): it works to ~1e-6 when the matrix the SH kernels read is a
ROTATION, but the reference hands the kernels the first nine floats of the [3,4] c2w (translation elements mixed in,
SURVEY App. A.7), and for that matrix the direction swings ~2.4x faster across the image: the quadratic is 25-300x
less accurate and most tiles miss a 5e-6 budget.  A drop-in has to reproduce the nine-float read, so the idea was
dropped before any kernel time was spent on it.

    python tools/sh_tile_expansion_error.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gsgen_b200.scenes import make_scene  # noqa: E402


def sh_basis16(d):
    """real SH up to degree 3 in the reference's order and signs (shencoder.h:24-56); d [...,3] unit vectors"""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    return np.stack([
        0.28209479177387814 * np.ones_like(x), -0.48860251190291987 * y, 0.48860251190291987 * z,
        -0.48860251190291987 * x, 1.0925484305920792 * xy, -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * y * (-3.0 * x2 + y2),
        2.8906114426405538 * xy * z, 0.45704579946446572 * y * (1.0 - 5.0 * z2),
        0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)], axis=-1)


def tile_errors(cam, M):
    """max |Y_k - 9-point-stencil quadratic| per 16x16 tile for the 3x3 matrix M the kernels multiply (x, y, 1) by"""
    def basis_at(gx, gy):  # pixel position as the kernels form it: (g - c) / f, no half-pixel offset
        p = np.stack([(gx - cam.cx) / cam.fx, (gy - cam.cy) / cam.fy, np.ones_like(gx)], -1) @ M.T
        return sh_basis16(p / np.linalg.norm(p, axis=-1, keepdims=True))

    H, W = cam.h, cam.w
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    Y = basis_at(xs.astype(np.float64), ys.astype(np.float64))
    lx, ly = np.meshgrid(np.arange(16), np.arange(16))
    u, v = (lx - 7.5).reshape(-1), (ly - 7.5).reshape(-1)
    A = np.stack([np.ones_like(u), u, v, u * u, u * v, v * v], -1)
    h, errs = 8.0, []
    for ty in range(H // 16):
        for tx in range(W // 16):
            cx, cy = 16 * tx + 7.5, 16 * ty + 7.5
            f = {(i, j): basis_at(np.float64(cx + i * h), np.float64(cy + j * h)) for i in (-1, 0, 1) for j in (-1, 0, 1)}
            st = np.stack([f[0, 0], (f[1, 0] - f[-1, 0]) / (2 * h), (f[0, 1] - f[0, -1]) / (2 * h),
                           (f[1, 0] - 2 * f[0, 0] + f[-1, 0]) / (2 * h * h),
                           (f[1, 1] - f[1, -1] - f[-1, 1] + f[-1, -1]) / (4 * h * h),
                           (f[0, 1] - 2 * f[0, 0] + f[0, -1]) / (2 * h * h)], 0)
            errs.append(np.abs(A @ st - Y[16 * ty:16 * ty + 16, 16 * tx:16 * tx + 16].reshape(-1, 16)).max())
    return np.array(errs)


if __name__ == "__main__":
    for cfg in ("c2", "c3", "c4", "c5"):
        sc = make_scene(cfg, N=8)
        cam, c2w = sc.cams[0], sc.c2ws[0].numpy().astype(np.float64)
        for name, M in (("rotation c2w[:3,:3]     ", c2w[:3, :3]),
                        ("nine-float read of c2w  ", c2w.reshape(-1)[:9].reshape(3, 3))):
            e = tile_errors(cam, M)
            print(f"{cfg} {cam.w}x{cam.h}  {name} max basis error {e.max():.1e}   tiles within 5e-6: {(e <= 5e-6).mean():.3f}"
                  f"   within 2e-5: {(e <= 2e-5).mean():.3f}")
