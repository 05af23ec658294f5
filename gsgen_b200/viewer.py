"""The renderer side of the interactive viewer: what `ViserViewer.update`
(utils/viewer/viser_viewer.py:129-171) does for every connected client, without the viser server (UI / transport are
out of scope, §6 of DESIGN.md; viser is not in this image either):

    camera_info = CameraInfo.from_fov_camera(fov, aspect, resolution, near, far)
    c2w         = [qvec2rotmat(camera.wxyz) | camera.position]                       (viser_viewer.py:14-19)
    frame       = (renderer.render_one(c2w, camera_info)["rgb"].detach().cpu().clamp(0, 1).numpy() * 255).astype(uint8)

under `torch.no_grad()`, plus the frames-per-second bookkeeping over the last three frames.  Here the clamp, the scale
and the uint8 cast run on the device and 3 B/pixel instead of 12 cross PCIe -- the bytes are the same (clamp, fp32
multiply and the truncating cast are exactly the operations numpy performs on the host copy).
"""
from __future__ import annotations

import time
from collections import deque
from typing import Deque, Optional

import numpy as np
import torch

from .camera import CameraInfo


def qvec2rotmat(qvec) -> np.ndarray:
    """utils/transforms.py:12-31 ((w,x,y,z) unit quaternion -> rotation, numpy)"""
    w, x, y, z = (float(v) for v in qvec)
    return np.array([[1 - 2 * y ** 2 - 2 * z ** 2, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x ** 2 - 2 * z ** 2, 2 * y * z - 2 * w * x],
                     [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x ** 2 - 2 * y ** 2]])


def get_c2w(wxyz, position) -> np.ndarray:
    """viser_viewer.py:14-19"""
    c2w = np.zeros([3, 4], dtype=np.float32)
    c2w[:3, :3] = qvec2rotmat(wxyz)
    c2w[:3, 3] = position
    return c2w


class ViewerLoop:
    """renderer: anything with `render_one(c2w, camera_info) -> {"rgb": [H,W,3]}` (GaussianSplattingRenderer in eval
    mode: no gradient arena, no densification statistics are touched)."""

    def __init__(self, renderer, resolution: int = 512, near_plane: float = 0.01, far_plane: float = 100.0):
        self.renderer = renderer
        self.resolution, self.near_plane, self.far_plane = resolution, near_plane, far_plane
        self.render_times: Deque[float] = deque(maxlen=3)  # viser_viewer.py:33
        self._host: Optional[torch.Tensor] = None

    @property
    def fps(self) -> float:
        return 1.0 / float(np.mean(self.render_times)) if self.render_times else 0.0

    @torch.no_grad()
    def render_frame(self, fov: float, aspect: float, wxyz, position) -> np.ndarray:
        """one client's frame: uint8 [H,W,3]"""
        cam = CameraInfo.from_fov_camera(fov, aspect, self.resolution, self.near_plane, self.far_plane)
        c2w = torch.from_numpy(get_c2w(wxyz, position))  # host pose: the fused path takes it by value (no H2D, no sync)
        start = time.time()
        rgb = self.renderer.render_one(c2w, cam)["rgb"].detach()
        frame = (rgb.clamp(min=0.0, max=1.0) * 255.0).to(torch.uint8)
        if frame.is_cuda:
            if self._host is None or self._host.shape != frame.shape:
                self._host = torch.empty(frame.shape, dtype=torch.uint8).pin_memory()
            self._host.copy_(frame, non_blocking=True)
            torch.cuda.current_stream(frame.device).synchronize()
            out = self._host.numpy().copy()
        else:
            out = frame.numpy()
        self.render_times.append(time.time() - start)
        return out
