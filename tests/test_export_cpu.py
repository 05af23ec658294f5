"""`.splat` / `.ply` writers (gsgen_b200/export.py).  The `.splat` bytes must equal the file the
reference's own `to_splat` wrote for the same parameters (tests/golden/make_splat_golden.py executes that function
unmodified); the `.ply` payload is checked against the field list of `to_ply` and round-tripped."""
import os

import numpy as np
import torch

from gsgen_b200 import export

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _params():
    z = np.load(os.path.join(GOLD, "ref_params_300.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_splat_bytes_equal_reference_file(tmp_path):
    params = _params()
    ref = open(os.path.join(GOLD, "ref_params_300.splat"), "rb").read()
    ours = export.splat_bytes(params)
    assert len(ours) == len(ref) == 300 * 32
    assert ours == ref
    path = str(tmp_path / "out" / "scene.splat")
    assert export.write_splat(params, path) == 300
    rec = export.read_splat(path)
    assert rec.tobytes() == ref
    # order: descending volume * opacity, the two identical Gaussians (4 and 10) adjacent in index order
    vol = rec["scale"].prod(axis=1) * rec["rgba"][:, 3]
    assert np.all(vol[:-1] >= vol[1:])
    pos = params["mean"].numpy()
    where = [int(np.where((rec["pos"] == pos[i]).all(axis=1))[0][0]) for i in (4, 10)]
    same = np.where((rec["pos"] == pos[4]).all(axis=1))[0]
    assert len(same) == 2 and same[1] == same[0] + 1 and where[0] == same[0]


def test_splat_cast_edge_cases():
    params = _params()
    rec = export.splat_records(params)
    pos = params["mean"].numpy()
    row = lambda i: rec[np.where((rec["pos"] == pos[i]).all(axis=1))[0][0]]
    assert row(0)["rgba"][:3].tolist() == [255, 0, 127]      # sigmoid(40) * 255 = 255, sigmoid(-40) -> 0, 0.5 -> 127
    assert row(1)["rgba"][3] == 255
    assert row(2)["rot"].tolist() == [0, 128, 128, 128]      # 1 * 128 + 128 = 256 wraps to 0 in the uint8 cast
    assert row(3)["rot"].tolist() == [128, 0, 128, 128]


def test_ply_layout_and_roundtrip(tmp_path):
    params = _params()
    path = str(tmp_path / "scene.ply")
    assert export.write_ply(params, path) == 300
    raw = open(path, "rb").read()
    head = raw[: raw.index(b"end_header\n") + 11].decode()
    assert head.startswith("ply\nformat binary_little_endian 1.0\nelement vertex 300\nproperty float x\n")
    assert head.count("property float") == 17 and len(raw) == len(head) + 300 * 17 * 4
    cols = export.read_ply(path)
    assert list(cols) == list(export.PLY_FIELDS)
    m, c = params["mean"].numpy(), params["color"].numpy()
    assert np.array_equal(np.stack([cols["x"], cols["y"], cols["z"]], 1), m)
    assert not cols["nx"].any() and not cols["ny"].any() and not cols["nz"].any()
    assert np.array_equal(np.stack([cols["red"], cols["green"], cols["blue"]], 1), c * 255.0)  # RAW colour * 255
    assert np.array_equal(cols["opacity"], params["alpha"].numpy())                             # RAW alpha
    assert np.array_equal(np.stack([cols[f"scale_{i}"] for i in range(3)], 1), params["svec"].numpy())
    assert np.array_equal(np.stack([cols[f"rot_{i}"] for i in range(4)], 1), params["qvec"].numpy())


def test_export_from_a_store():
    from gsgen_b200.store import GaussianStore

    params = _params()
    st = GaussianStore(params, None, "cpu", capacity=512)
    assert export.splat_bytes(st.params) == open(os.path.join(GOLD, "ref_params_300.splat"), "rb").read()


def test_viewer_loop_frame_equals_the_reference_expression():
    """gsgen_b200.viewer.ViewerLoop.render_frame == the expression of ViserViewer.update (viser_viewer.py:147-155) on the
    same rendered image, camera built by from_fov_camera (utils/camera.py:316-325) and get_c2w (viser_viewer.py:14-19)."""
    import numpy as np

    from gsgen_b200.viewer import ViewerLoop, get_c2w

    seen = {}

    class R:
        def render_one(self, c2w, cam):
            seen["c2w"], seen["cam"] = c2w, cam
            g = torch.Generator().manual_seed(0)
            self.img = torch.randn(cam.h, cam.w, 3, generator=g) * 0.7 + 0.5  # values below 0 and above 1 too
            return {"rgb": self.img}

    r = R()
    loop = ViewerLoop(r, resolution=64)
    q = np.array([0.5, 0.5, -0.5, 0.5])
    frame = loop.render_frame(fov=0.9, aspect=1.6, wxyz=q, position=[1.0, 2.0, 3.0])
    want = (r.img.detach().cpu().clamp(min=0.0, max=1.0).numpy() * 255.0).astype(np.uint8)
    assert frame.dtype == np.uint8 and np.array_equal(frame, want)
    cam = seen["cam"]
    assert (cam.w, cam.h) == (64, int(64 / 1.6)) and abs(cam.fy - (cam.h / 2) / np.tan(0.45)) < 1e-9
    c2w = get_c2w(q, [1.0, 2.0, 3.0])
    assert np.allclose(c2w[:, :3] @ c2w[:, :3].T, np.eye(3), atol=1e-6) and np.allclose(c2w[:, 3], [1, 2, 3])
    assert torch.equal(seen["c2w"], torch.from_numpy(c2w)) and loop.fps > 0
