"""Asset ingest (assets.py) against files written the way the reference's assets are stored (sim.py:41-62): ascii and binary PLY
point clouds, an OBJ surface with v/vt/vn face records, the placement transform, the checkpoint naming, and a state dict saved
from the reference's own DeepSDFWithCode loading into the product's DeepSDF front-end (weight-norm keys)."""
import os
import struct

import numpy as np
import pytest

from meshless_inflatable_softbody_b200 import assets


def _write_ply(path, pts, fmt):
    with open(path, "wb") as f:
        hdr = f"ply\nformat {fmt} 1.0\ncomment test\nelement vertex {len(pts)}\nproperty float x\nproperty float y\nproperty float z\n" \
              f"property uchar red\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n"
        f.write(hdr.encode())
        for p in pts:
            if fmt == "ascii":
                f.write(f"{float(p[0])!r} {float(p[1])!r} {float(p[2])!r} 255\n".encode())
            else:
                f.write(struct.pack(("<" if "little" in fmt else ">") + "fffB", *[float(v) for v in p], 255))


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_ply_points_round_trip(tmp_path, fmt):
    pts = np.random.default_rng(0).normal(size=(57, 3)).astype(np.float32)
    _write_ply(tmp_path / "a.ply", pts, fmt)
    got = assets.read_ply_points(str(tmp_path / "a.ply"))
    assert got.shape == (57, 3) and np.array_equal(got.astype(np.float32), pts)


def test_obj_faces_and_fan_triangulation(tmp_path):
    (tmp_path / "m.obj").write_text("# test\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvn 0 0 1\nf 1/1/1 2/1/1 3/1/1\nf 1 3 4\nf 1//1 2//1 3//1 4//1\nf -4 -3 -2\n")
    v, f = assets.read_obj_mesh(str(tmp_path / "m.obj"))
    assert v.shape == (4, 3)
    assert f.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2], [0, 2, 3], [0, 1, 2]]


def test_scene_assembly_matches_sim_py(tmp_path):
    rng = np.random.default_rng(1)
    outer, inner = rng.normal(size=(30, 3)) * 5, rng.normal(size=(50, 3)) * 3
    d = tmp_path / "pcd" / "bunny"
    os.makedirs(d)
    _write_ply(d / "point_cloud_downsampled.ply", outer.astype(np.float32), "binary_little_endian")
    _write_ply(d / "bunny_inner.ply", inner.astype(np.float32), "ascii")
    (d / "outer.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    np.save(d / "uv.npy", rng.uniform(size=(30, 2)))
    a = assets.load_asset(str(tmp_path / "pcd"), "bunny")
    # sim.py:46-53 restated
    R = np.array([[1., 0., 0.], [0., 0., -1.], [0., 1., 0.]])
    pts = np.vstack([outer.astype(np.float32).astype(np.float64) * 0.01, inner.astype(np.float32).astype(np.float64) * 0.01])
    assert a.out_num == 30 and a.faces.tolist() == [[0, 1, 2]] and a.uv.shape == (30, 2)
    assert np.allclose(a.points_model, pts, atol=1e-9) and np.allclose(a.points_world, pts @ R + np.array([0., .07, 0.]), atol=1e-7)
    # the inverse placement used for contact queries maps world back to model space
    from meshless_inflatable_softbody_b200.deepsdf import world_to_model_xform
    xf = np.asarray(world_to_model_xform(), np.float64)
    A, t = xf[:9].reshape(3, 3), xf[9:]
    assert np.allclose((a.points_world.astype(np.float64) - t) @ A.T, a.points_model, atol=1e-6)


def test_checkpoint_naming(tmp_path):
    d = tmp_path / "model" / "bunny"
    os.makedirs(d)
    assert assets.checkpoint_path(str(tmp_path / "model"), "bunny").endswith("model_10000.pth")     # sim.py:59-60 fallback
    np.save(d / "min_loss_index.npy", np.array(4200))
    assert assets.checkpoint_path(str(tmp_path / "model"), "bunny").endswith("model_4200.pth")


@pytest.mark.gpu
def test_reference_style_checkpoint_loads_into_the_engine(tmp_path):
    """torch.save(state_dict) of a weight-normalised 9 x 1024 network with the reference's key names -> DeepSDF(state_dict): same
    values as the oracle's forward on the same weights."""
    import torch
    from meshless_inflatable_softbody_b200 import DeepSDF
    from oracle import deepsdf_oracle as do
    torch.manual_seed(3)
    m = do.reference_like_module()
    path = tmp_path / "model_10000.pth"
    torch.save(m.state_dict(), path)
    st = torch.load(path, map_location="cpu")
    net = DeepSDF(st)
    p = np.random.default_rng(2).uniform(-0.5, 0.5, size=(300, 3)).astype(np.float32)
    want = do.forward({k: v.numpy() for k, v in st.items()}, p, np.float64)[:, 0]
    got = net(p).cpu().numpy()[:, 0]
    assert np.abs(got - want).max() <= 4e-6 * np.abs(want).max() + 1e-7
