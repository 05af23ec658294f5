"""Wire / on-disk formats next to the hot path: the `.splat` records of the web viewer and the
17-float `.ply` of the official 3DGS tools, written from the raw leaves (checkpoint `params` or a `GaussianStore`).

Reference: `utils/export.py` -- `to_splat` (:212-284) packs one 32-byte record per Gaussian in a Python loop of
`struct.pack` calls after sorting the indices with a Python key function; `to_ply` (:159-209) builds a structured
array through a list of tuples and hands it to `plyfile`.  Here both are whole-array numpy encodes (the formats are
plain byte layouts, no arithmetic beyond the reference's own casts, which are reproduced expression for expression
so that the bytes are identical):

  .splat record (32 B, little endian):  3 x f32 position | 3 x f32 exp(svec) | u8 RGB = trunc(sigmoid(color) * 255)
                                        u8 A = trunc(sigmoid(alpha) * 255)   | 4 x u8 trunc(q/|q| * 128 + 128)
                 order: descending `prod(exp(svec)) * A` (A the uint8), ties in index order (stable).
  .ply vertex (68 B): x y z nx ny nz red green blue opacity scale_0..2 rot_0..3 as f32, with the reference's choices:
                 normals 0, `color * 255` / `alpha` / `svec` RAW (no activation) -- utils/export.py:193-198.

This is host-side format code: where the reference is Python the host side is Python (numpy); no GPU involved.
`.ply`: plyfile is not in this image, so the header is restated from the PLY specification as plyfile writes it
("parity unpinned" for the header text; the payload layout is pinned by the field list above).
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

PLY_FIELDS = ("x", "y", "z", "nx", "ny", "nz", "red", "green", "blue", "opacity", "scale_0", "scale_1", "scale_2",
              "rot_0", "rot_1", "rot_2", "rot_3")  # utils/export.py:171-189
SPLAT_RECORD = np.dtype([("pos", "<f4", 3), ("scale", "<f4", 3), ("rgba", "u1", 4), ("rot", "u1", 4)])
assert SPLAT_RECORD.itemsize == 32


def _cpu(params: Dict[str, torch.Tensor], name: str) -> torch.Tensor:
    return params[name].detach().to("cpu", torch.float32)


def splat_records(params: Dict[str, torch.Tensor]) -> np.ndarray:
    """raw leaves {mean [N,3], svec [N,3], color [N,3], alpha [N], qvec [N,4]} -> sorted structured array [N] of
    32-byte records (utils/export.py:246-281)."""
    pos = _cpu(params, "mean").numpy()
    rgb = (torch.sigmoid(_cpu(params, "color")).numpy() * 255.0).astype(np.uint8).clip(0, 255)
    opacity = (torch.sigmoid(_cpu(params, "alpha")).numpy().reshape(-1, 1) * 255.0).astype(np.uint8).clip(0, 255)
    svec = np.exp(_cpu(params, "svec").numpy())
    qvec = _cpu(params, "qvec").numpy()
    qvec = qvec / np.linalg.norm(qvec, axis=1, keepdims=True)
    qvec = (qvec * 128 + 128).astype(np.uint8).clip(0, 255)
    volume = np.prod(svec, axis=1) * opacity[..., 0]
    # sorted(range(n), key=volume.__getitem__, reverse=True): descending, equal keys keep their index order
    order = np.argsort(-volume, kind="stable")
    rec = np.empty(pos.shape[0], dtype=SPLAT_RECORD)
    rec["pos"], rec["scale"] = pos, svec
    rec["rgba"][:, :3], rec["rgba"][:, 3] = rgb, opacity[:, 0]
    rec["rot"] = qvec
    return rec[order]


def splat_bytes(params: Dict[str, torch.Tensor]) -> bytes:
    return splat_records(params).tobytes()


def write_splat(params: Dict[str, torch.Tensor], path: str) -> int:
    """-> number of Gaussians written"""
    rec = splat_records(params)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(rec.tobytes())
    return int(rec.shape[0])


def read_splat(path: str) -> np.ndarray:
    data = np.fromfile(path, dtype=SPLAT_RECORD)
    if data.nbytes != os.path.getsize(path):
        raise RuntimeError(f"{path}: size is not a multiple of the 32-byte record")
    return data


def ply_bytes(params: Dict[str, torch.Tensor]) -> bytes:
    """binary little-endian PLY, one `vertex` element with the 17 float properties of the official tools."""
    pos = _cpu(params, "mean").numpy()
    attributes = np.concatenate((pos, np.zeros_like(pos), _cpu(params, "color").numpy() * 255.0,
                                 _cpu(params, "alpha").numpy().reshape(-1, 1), _cpu(params, "svec").numpy(),
                                 _cpu(params, "qvec").numpy()), axis=1).astype("<f4")
    if attributes.shape[1] != len(PLY_FIELDS):
        raise RuntimeError(f"expected {len(PLY_FIELDS)} columns, got {attributes.shape[1]}")
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % pos.shape[0]
    header += "".join(f"property float {name}\n" for name in PLY_FIELDS) + "end_header\n"
    return header.encode("ascii") + np.ascontiguousarray(attributes).tobytes()


def write_ply(params: Dict[str, torch.Tensor], path: str) -> int:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    data = ply_bytes(params)
    with open(path, "wb") as f:
        f.write(data)
    return int(params["mean"].shape[0])


def read_ply(path: str) -> Dict[str, np.ndarray]:
    """the inverse of write_ply (binary little-endian, float properties only)"""
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    lines = raw[:end].decode("ascii").splitlines()
    if lines[0] != "ply" or not lines[1].startswith("format binary_little_endian"):
        raise RuntimeError(f"{path}: not a binary little-endian PLY")
    n = int(next(l for l in lines if l.startswith("element vertex")).split()[-1])
    names = [l.split()[-1] for l in lines if l.startswith("property float")]
    body = np.frombuffer(raw[end:], dtype="<f4").reshape(n, len(names))
    return {name: body[:, i] for i, name in enumerate(names)}
