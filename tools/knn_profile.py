"""one warm-up + one measured gsb200_knn over 1M points uniform in the unit ball, K=4 (run under ncu for the launch list)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsgen_b200.knn import knn_points  # noqa: E402

g = torch.Generator().manual_seed(2)
n = 1_000_000
v = torch.randn(n, 3, generator=g)
pts = (v / v.norm(dim=1, keepdim=True) * torch.rand(n, 1, generator=g) ** (1 / 3)).cuda()
for _ in range(2):
    d2, idx = knn_points(None, pts, 4)
torch.cuda.synchronize()
print("ok", int(idx[:, 1].min()), float(d2[:, 1].mean()))
