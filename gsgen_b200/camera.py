"""Camera model on the host side of the hot path.

Mirrors the parts of the reference's `utils/camera.py:219-368` (`CameraInfo`) and
`data/__init__.py:14-29` (`get_c2w_from_up_and_look_at`) that feed the rasterizer: intrinsics,
the 6-plane frustum (`get_frustum`, :260-294), camera-plane -> pixel conversion (:301-314) and the
per-pixel ray directions (:327-346).  Same attribute / method names so the reference trainer's
`camera_info` objects and ours are interchangeable (duck typing).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


class CameraInfo:
    """OpenCV pinhole camera (+x right, +y down, +z forward), reference utils/camera.py:219-231."""

    def __init__(self, fx, fy, cx, cy, w, h, near_plane=0.01, far_plane=100.0) -> None:
        self.fx, self.fy = float(fx), float(fy)
        self.cx, self.cy = float(cx), float(cy)
        self.w, self.h = int(w), int(h)
        self.yfov = 2 * np.arctan(self.h / (2 * self.fy))
        self.aspect = self.w / self.h
        self.near_plane = float(near_plane)
        self.far_plane = float(far_plane)

    @classmethod
    def from_reso(cls, reso: int, focal: float = 1.0, near_plane=0.01, far_plane=100.0):
        """`CameraPoseProvider` convention (data/__init__.py:188-197): fx=fy=focal*reso, cx=cy=reso/2."""
        return cls(focal * reso, focal * reso, reso / 2.0, reso / 2.0, reso, reso, near_plane, far_plane)

    @classmethod
    def from_fov_camera(cls, fov, aspect, resolution, near_plane, far_plane):
        """utils/camera.py:316-325 (the viewer's camera: vertical fov in radians, width = `resolution`)"""
        W = resolution
        H = int(resolution / aspect)
        cx = W / 2
        cy = H / 2
        fx = cx / np.tan(fov / 2)
        fy = cy / np.tan(fov / 2)
        return cls(fx, fy, cx, cy, W, H, near_plane, far_plane)

    # ---- resolution changes the trainer applies to its cameras (utils/camera.py:230-258; trainer.py:786-787 calls
    # cam_info.set_reso(reso) for the up-sampling stage) ----------------------------------------------------------
    def _refresh(self):
        self.yfov = 2 * np.arctan(self.h / (2 * self.fy))
        self.aspect = self.w / self.h

    def downsample(self, scale):
        self.fx /= scale; self.fy /= scale; self.cx /= scale; self.cy /= scale
        self.w //= scale; self.h //= scale
        self._refresh()

    def upsample(self, scale):
        """(the reference refreshes neither yfov nor aspect here; both are invariant under a uniform integer scale)"""
        self.fx *= scale; self.fy *= scale; self.cx *= scale; self.cy *= scale
        self.w *= int(scale); self.h *= int(scale)

    def set_reso(self, reso: int):
        self.fx = self.fy = reso
        self.cx = self.cy = reso / 2.0
        self.w = self.h = reso
        self._refresh()

    def get_camera_intrinsic(self, device="cuda"):
        """utils/camera.py:361-368"""
        return torch.tensor([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]]).to(device)

    def get_frustum(self, c2w: torch.Tensor):
        """-> (normals[6,3], pts[6,3]) fp32 on c2w's device; reference utils/camera.py:260-294."""
        up = -c2w[:, 1]
        right = c2w[:, 0]
        lookat = c2w[:, 2]
        t = c2w[:, 3]
        half_vside = self.far_plane * np.tan(self.yfov * 0.5)
        half_hside = half_vside * self.aspect
        near_point = self.near_plane * lookat
        far_point = self.far_plane * lookat
        left_normal = torch.linalg.cross(far_point - half_hside * right, up)
        right_normal = torch.linalg.cross(up, far_point + half_hside * right)
        up_normal = torch.linalg.cross(far_point + half_vside * up, right)
        down_normal = torch.linalg.cross(right, far_point - half_vside * up)
        pts = torch.stack([near_point + t, far_point + t, t, t, t, t], dim=0)
        normals = torch.stack([lookat, -lookat, left_normal, right_normal, up_normal, down_normal], dim=0)
        normals = F.normalize(normals, dim=-1)
        return normals.contiguous(), pts.contiguous()

    def camera_space_to_pixel_space(self, pts):
        """Reference utils/camera.py:301-314 (in-place multiply-add, truncating int32 cast)."""
        if pts.shape[1] == 3:
            pts = pts[:, :2] / pts[:, 2:]
        assert pts.shape[1] == 2
        pts[:, 0] = pts[:, 0] * self.fx + self.cx
        pts[:, 1] = pts[:, 1] * self.fy + self.cy
        return pts.to(torch.int32)

    def get_rays_d(self, c2w):
        """[H,W,3] un-normalised world ray directions; reference utils/camera.py:327-346."""
        xp = (torch.arange(0, self.w, dtype=torch.float32, device=c2w.device) - self.cx) / self.fx
        yp = (torch.arange(0, self.h, dtype=torch.float32, device=c2w.device) - self.cy) / self.fy
        xp, yp = torch.meshgrid(xp, yp, indexing="ij")
        xyz = torch.stack([xp.reshape(-1), yp.reshape(-1), torch.ones_like(xp.reshape(-1))], dim=-1)
        return torch.einsum("ij,bj->bi", c2w[:3, :3], xyz).reshape(self.w, self.h, 3).transpose(0, 1)

    @property
    def n_tiles(self):
        ts = 16
        return (self.h // ts + (self.h % ts > 0), self.w // ts + (self.w % ts > 0))


def get_c2w_from_up_and_look_at(up, look_at, pos) -> np.ndarray:
    """Reference data/__init__.py:14-29: columns = (right, down, lookat, position)."""
    up = np.asarray(up, dtype=np.float64)
    look_at = np.asarray(look_at, dtype=np.float64)
    pos = np.asarray(pos, dtype=np.float64)
    up = up / np.linalg.norm(up)
    z = look_at - pos
    z = z / np.linalg.norm(z)
    y = -up
    x = np.cross(y, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    c2w = np.zeros([3, 4], dtype=np.float32)
    c2w[:3, 0] = x
    c2w[:3, 1] = y
    c2w[:3, 2] = z
    c2w[:3, 3] = pos
    return c2w


def orbit_c2w(distance: float, elevation_deg: float, azimuth_deg: float, center=(0.0, 0.0, 0.0)) -> torch.Tensor:
    """Random-orbit pose convention of `CameraPoseProvider.sample_one` (data/__init__.py:151-184):
    up = +z, look at `center`, pos = d*(cos e cos a, cos e sin a, sin e)."""
    e, a = np.deg2rad(elevation_deg), np.deg2rad(azimuth_deg)
    pos = np.array([distance * np.cos(e) * np.cos(a), distance * np.cos(e) * np.sin(a), distance * np.sin(e)])
    return torch.from_numpy(get_c2w_from_up_and_look_at(np.array([0.0, 0.0, 1.0]), np.asarray(center), pos))
