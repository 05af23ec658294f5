"""Golden vectors for the PER-GAUSSIAN torch stage from the UNMODIFIED reference functions (TEST INFRASTRUCTURE; dev
container only -- needs /root/reference):

    utils/camera.py   CameraInfo.get_frustum, .camera_space_to_pixel_space, .get_rays_d
    gs/renderer.py    jacobian, project_pts, project_gaussians   (+ their autograd)
    utils/transforms.py  qsvec2rotmat_batched
    gs/culling.py     tile_culling_aabb_count

The modules cannot be imported here (kornia, torchtyping, plyfile ... are absent), so the function / class definitions
are compiled out of the files with `ast` exactly as they are (type annotations and the @lineprofiler decorator
dropped) and executed on seeded CPU inputs.  The only stub is kornia's `quaternion_to_rotation_matrix` (= the
restatement in oracle/__init__.py): quaternions in the fixture are unit length, where that convention is pinned by the
reference's own `qvec2rotmat` (kernels.h:49-58).  Output: tests/golden/pergaussian_ref.npz.
"""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def compile_defs(path, names, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            for sub in ast.walk(node):
                if isinstance(sub, ast.FunctionDef):
                    sub.returns = None
                    for a in sub.args.args + sub.args.kwonlyargs:
                        a.annotation = None
                    sub.decorator_list = [d for d in sub.decorator_list
                                          if not (isinstance(d, ast.Name) and d.id == "lineprofiler")]
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing


def main():
    import oracle
    from gsgen_b200.scenes import make_scene

    ns = {"torch": torch, "np": np, "F": torch.nn.functional,
          "QuaternionCoeffOrder": type("Q", (), {"WXYZ": "wxyz"}),
          "quaternion_to_rotation_matrix": lambda q, order: oracle.quat_to_rotmat(q),
          "console": type("C", (), {"print": staticmethod(lambda *a, **k: None)})()}
    compile_defs(f"{REF}/utils/transforms.py", ["qsvec2rotmat_batched"], ns)
    compile_defs(f"{REF}/gs/renderer.py", ["jacobian", "project_pts", "project_gaussians"], ns)
    compile_defs(f"{REF}/utils/camera.py", ["CameraInfo"], ns)
    compile_defs(f"{REF}/gs/culling.py", ["tile_culling_aabb_count"], ns)

    out = {}
    for tag, cfg, N, reso in (("a", "c1", 1200, 200), ("b", "c3", 1500, 168)):
        sc = make_scene(cfg, N=N, reso=reso)
        cam, c2w = sc.cams[0], sc.c2ws[0]
        rcam = ns["CameraInfo"](cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h, cam.near_plane, cam.far_plane)
        normals, pts = rcam.get_frustum(c2w)
        g = torch.Generator().manual_seed(17)
        for detach in (True, False):
            mean = sc.mean.clone().requires_grad_()
            qvec = sc.qvec.clone().requires_grad_()  # unit quaternions
            svec = sc.svec.clone().requires_grad_()
            m2, cov, JW, dp = ns["project_gaussians"](mean, qvec, svec, c2w, detach)
            gm2, gcov = torch.randn(m2.shape, generator=g), torch.randn(cov.shape, generator=g)
            gdp = torch.randn(dp.shape, generator=g)
            ((m2 * gm2).sum() + (cov * gcov).sum() + (dp * gdp).sum()).backward()
            k = f"{tag}_{'detach' if detach else 'full'}"
            if detach:
                out[f"{k}_JW"] = JW.detach()
            out.update({f"{k}_mean2d": m2.detach(), f"{k}_cov2d": cov.detach(),
                        f"{k}_depth": dp.detach(), f"{k}_g_mean2d": gm2, f"{k}_g_cov2d": gcov, f"{k}_g_depth": gdp,
                        f"{k}_grad_mean": mean.grad.clone(), f"{k}_grad_qvec": qvec.grad.clone(),
                        f"{k}_grad_svec": svec.grad.clone()})
        m2, cov, _, dp = ns["project_gaussians"](sc.mean, sc.qvec, sc.svec, c2w, True)
        front = dp.reshape(-1) > 0.05  # the AABB of a Gaussian behind the camera is not meaningful
        D, tl, br = ns["tile_culling_aabb_count"](m2[front], cov[front], 16, rcam, 6.0)
        out.update({f"{tag}_mean": sc.mean, f"{tag}_qvec": sc.qvec, f"{tag}_svec": sc.svec, f"{tag}_c2w": c2w,
                    f"{tag}_cam": torch.tensor([cam.fx, cam.fy, cam.cx, cam.cy, cam.w, cam.h, cam.near_plane,
                                                cam.far_plane], dtype=torch.float64),
                    f"{tag}_frustum_normals": normals, f"{tag}_frustum_pts": pts, f"{tag}_front": front,
                    f"{tag}_aabb_tl": tl, f"{tag}_aabb_br": br, f"{tag}_D": torch.tensor([int(D)]),
                    f"{tag}_rays_d": rcam.get_rays_d(c2w)[::7, ::5].contiguous()})
        print(tag, cfg, "N", N, "front", int(front.sum()), "D", int(D))
    path = os.path.join(ROOT, "tests", "golden", "pergaussian_ref.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print("wrote", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
