"""Generate the `.splat` golden file with the UNMODIFIED reference function `utils/export.py:to_splat`
(TEST INFRASTRUCTURE; dev container only -- needs /root/reference).  `utils/export.py` cannot be imported (plyfile,
kornia ... are absent), so the function definition is compiled out of the file with `ast` as it is and run on a
checkpoint written to a temp directory; its output file is copied to tests/golden/ref_params_300.splat next to the
parameters that produced it (tests/golden/ref_params_300.npz)."""
import ast
import os
import shutil
import struct
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/utils/export.py"


def main():
    tree = ast.parse(open(REF).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "to_splat")
    ns = {"np": np, "torch": torch, "struct": struct, "Path": Path, "get_ckpt_path": lambda p: Path(p),
          "console": type("C", (), {"print": staticmethod(lambda *a, **k: None)})()}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    g = torch.Generator().manual_seed(2024)
    N = 300
    params = dict(mean=torch.randn(N, 3, generator=g), qvec=torch.randn(N, 4, generator=g) * 2.0,
                  svec=torch.log(0.005 + 0.05 * torch.rand(N, 3, generator=g)), color=2.0 * torch.randn(N, 3, generator=g),
                  alpha=3.0 * torch.randn(N, generator=g))
    # edge cases of the casts: saturated sigmoid, axis-aligned unit quaternions (q*128+128 = 256 -> uint8 wrap), ties
    params["color"][0] = torch.tensor([40.0, -40.0, 0.0])
    params["alpha"][1] = 40.0
    params["qvec"][2] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    params["qvec"][3] = torch.tensor([0.0, -1.0, 0.0, 0.0])
    for f in ("mean", "qvec", "svec", "color", "alpha"):  # two identical Gaussians: equal sort keys
        params[f][10] = params[f][4]
    with tempfile.TemporaryDirectory() as d:
        ck = os.path.join(d, "ck.pt")
        torch.save({"cfg": {"prompt": {"prompt": "ref params"}}, "params": params}, ck)
        ns["to_splat"](ck, d)
        out = os.path.join(d, "splat", "ref_params.splat")
        shutil.copy(out, os.path.join(ROOT, "tests", "golden", "ref_params_300.splat"))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_params_300.npz"),
                        **{k: v.numpy() for k, v in params.items()})
    print("wrote", os.path.getsize(os.path.join(ROOT, "tests", "golden", "ref_params_300.splat")), "bytes")


if __name__ == "__main__":
    main()
