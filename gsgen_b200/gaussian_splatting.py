"""Import-path alias: the reference's callers do `from gs.gaussian_splatting import GaussianSplattingRenderer`
(trainer.py:17, vis.py:3, utils/export.py:6, ...); with this module the level-4 integration is that one import line
changed to `from gsgen_b200.gaussian_splatting import GaussianSplattingRenderer`.  The class itself lives in
gsgen_b200/splatting.py."""
from .splatting import GaussianSplattingRenderer  # noqa: F401

__all__ = ["GaussianSplattingRenderer"]
