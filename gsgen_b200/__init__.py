"""gsgen_b200 -- B200-native (sm_100a) differentiable Gaussian-splatting rasterizer behind the
`gs` API of gsgen3d/gsgen.  The hot path lives in `csrc/` (hand-written CUDA + a C-ABI shared
library, `include/gsb200.h`); this package is the thin Python host side that mirrors the
reference's operator interface (`gs/backend.py`, `gs/renderer.py`, `gs/culling.py`,
`GaussianSplattingRenderer.render_one`).  There is no CPU fallback: every op raises if
`libgsb200.so` is missing or the tensors are not CUDA tensors.
"""
__version__ = "0.1.0"
