"""TEST INFRASTRUCTURE -- compiles the reference's OWN Python for the hot path into code objects, straight from
/root/reference into oracle/_ref/ref_py.bin (git-ignored like oracle/_ref/_gs.so, NOT gpurun-ignored: it travels to the
B200 box, where /root/reference does not exist).  No reference source is copied into this repo: the output holds
compiled CPython code objects (marshal), the Python counterpart of the `_gs.so` that oracle/build_ref.sh compiles from
the reference's CUDA sources.

What it is for (round-1 verdict item 9, "drop-in proof on the GPU"): tests/test_dropin_gpu.py executes these code
objects on the B200 twice -- once with `_backend` = the unmodified reference extension (`_gs.so`), once with
`_backend` = gsgen_b200.backend._backend -- i.e. the reference's own `GaussianSplattingRenderer.render_one`,
`SHRenderer.forward` and the autograd Functions of gs/renderer.py run UNCHANGED over libgsb200.so, and the two arms are
compared image for image and gradient for gradient.

Definitions compiled (as they are; only type annotations and the `@lineprofiler` decorator are dropped, because
`torchtyping` / `line_profiler` are not installed in this image):
  utils/transforms.py        qsvec2rotmat_batched
  gs/renderer.py             jacobian, project_pts, project_gaussians, _render_with_T, _render_scalar, _render_sh,
                             _render_sh_bg
  utils/camera.py            CameraInfo
  gs/culling.py              tile_culling_aabb_count
  gs/gaussian_splatting.py   GaussianSplattingRenderer.render_one, .get_with_overrides
  gs/sh_renderer.py          SHRenderer.forward, .get_with_overrides

    python oracle/build_ref_py.py          # no-op when /root/reference is absent (GPU box)
"""
from __future__ import annotations

import ast
import importlib.util
import marshal
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(HERE, "_ref", "ref_py.bin")

SPEC = [
    ("utils/transforms.py", None, ["qsvec2rotmat_batched"]),
    ("gs/renderer.py", None, ["jacobian", "project_pts", "project_gaussians", "_render_with_T", "_render_scalar",
                              "_render_sh", "_render_sh_bg"]),
    ("utils/camera.py", None, ["CameraInfo"]),
    ("gs/culling.py", None, ["tile_culling_aabb_count"]),
    ("gs/gaussian_splatting.py", "GaussianSplattingRenderer", ["render_one", "get_with_overrides"]),
    ("gs/sh_renderer.py", "SHRenderer", ["forward", "get_with_overrides"]),
]


def compile_defs(path: str, names, cls=None):
    """[(name, code object)] of the named top-level definitions (or methods of `cls`) of `path`."""
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    found = {}
    for node in body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names and node.name not in found:
            for sub in ast.walk(node):
                if isinstance(sub, ast.FunctionDef):
                    sub.returns = None
                    for a in sub.args.args + sub.args.kwonlyargs:
                        a.annotation = None
                    sub.decorator_list = [d for d in sub.decorator_list
                                          if not (isinstance(d, ast.Name) and d.id == "lineprofiler")]
            found[node.name] = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
    missing = [n for n in names if n not in found]
    if missing:
        raise RuntimeError(f"{path}: {missing} not found")
    return [(n, found[n]) for n in names]


def build(force: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(REF, "gs")):
        print(f"build_ref_py: {REF}/gs not present (GPU box?) -- using prebuilt {OUT} if any")
        return OUT if os.path.exists(OUT) else None
    srcs = [os.path.join(REF, rel) for rel, _, _ in SPEC] + [os.path.abspath(__file__)]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(s) for s in srcs):
        print(f"build_ref_py: {OUT} up to date")
        return OUT
    entries = []
    for rel, cls, names in SPEC:
        for name, code in compile_defs(os.path.join(REF, rel), names, cls):
            entries.append({"file": rel, "cls": cls, "name": name, "code": code})
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    blob = {"python": tuple(sys.version_info[:3]), "magic": importlib.util.MAGIC_NUMBER, "entries": entries}
    with open(OUT, "wb") as f:
        marshal.dump(blob, f)
    print(f"build_ref_py: wrote {OUT} ({os.path.getsize(OUT)} bytes, {len(entries)} definitions)")
    return OUT


def load(path: str = OUT):
    """entries of ref_py.bin, or None when the file is absent / was compiled by another CPython"""
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        blob = marshal.load(f)
    if blob.get("magic") != importlib.util.MAGIC_NUMBER:
        return None
    return blob["entries"]


if __name__ == "__main__":
    build(force="--force" in sys.argv)
