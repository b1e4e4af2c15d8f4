"""Asset ingest of the reference scene set-up (sim.py:41-62), without open3d / trimesh (not installed here): the point clouds
(`point_cloud_downsampled.ply`, `<name>_inner.ply`), the surface mesh (`outer.obj`), `uv.npy` and the DeepSDF checkpoint
(`model_<k>.pth`, `min_loss_index.npy`) of one asset folder pair.  Readers cover what those files use: PLY vertex elements with
x/y/z properties (ascii or binary little/big endian, extra properties ignored) and OBJ `v` / `f` records (`f` entries may be
`v`, `v/vt` or `v/vt/vn`, polygons are fan-triangulated like trimesh does).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

# sim.py:46-52: points_np = points * 0.01, then points_np @ R + (0, 0.07, 0); model space = before the rotation and the lift
ASSET_SCALE = 0.01
ASSET_R = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]])
ASSET_LIFT = np.array([0.0, 0.07, 0.0])

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def read_ply_points(path: str) -> np.ndarray:
    """(n, 3) float64 vertex positions of a PLY file (what o3d.io.read_point_cloud(...).points holds, sim.py:41-42,47-48)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, n_vertex, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: header not terminated")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n_vertex = int(tok[2])
                elif n_vertex == 0:
                    raise ValueError(f"{path}: the vertex element must come first")
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list property in the vertex element")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        names = [p[0] for p in props]
        if not all(a in names for a in "xyz"):
            raise ValueError(f"{path}: vertex element without x / y / z")
        if fmt == "ascii":
            rows = [f.readline().split() for _ in range(n_vertex)]
            data = np.array([[float(r[names.index(a)]) for a in "xyz"] for r in rows], dtype=np.float64).reshape(-1, 3)
        else:
            end = "<" if fmt == "binary_little_endian" else ">"
            dt = np.dtype([(nm, end + ty) for nm, ty in props])
            raw = np.frombuffer(f.read(dt.itemsize * n_vertex), dtype=dt, count=n_vertex)
            data = np.stack([raw[a].astype(np.float64) for a in "xyz"], 1)
    return data


def read_obj_mesh(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """(vertices (n, 3) float64, faces (m, 3) int64, 0-based) of an OBJ file (trimesh.load_mesh(...).faces, sim.py:43-44)."""
    verts, faces = [], []
    with open(path, "r") as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "f":
                idx = [int(t.split("/")[0]) for t in tok[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]          # negative = relative to the vertices read so far
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(verts, np.float64).reshape(-1, 3), np.asarray(faces, np.int64).reshape(-1, 3)


def place_asset(points: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(model-space points, world-space points) from raw asset coordinates: x 0.01, then @ R + lift (sim.py:46-52).  The DeepSDF is
    queried at the model-space points (sim.py:49,100), the simulation starts from the world-space ones (sim.py:85)."""
    model = np.asarray(points, np.float64) * ASSET_SCALE
    return model, model @ ASSET_R + ASSET_LIFT


@dataclass
class Asset:
    points_model: np.ndarray          # (n, 3) float32, outer shell first
    points_world: np.ndarray          # (n, 3) float32 = init_position (sim.py:85)
    out_num: int                      # number of outer-shell particles = mesh vertices (sim.py:53)
    faces: Optional[np.ndarray]       # (m, 3) triangle indices into the first out_num particles (sim.py:44)
    uv: Optional[np.ndarray]          # texture coordinates (sim.py:45)


def load_asset(pcd_folder: str, name: str) -> Asset:
    """The scene of sim.py:41-53 from `<pcd_folder>/<name>/`."""
    d = os.path.join(pcd_folder, name)
    outer = read_ply_points(os.path.join(d, "point_cloud_downsampled.ply"))
    inner = read_ply_points(os.path.join(d, f"{name}_inner.ply"))
    model, world = place_asset(np.vstack([outer, inner]))
    faces = read_obj_mesh(os.path.join(d, "outer.obj"))[1] if os.path.exists(os.path.join(d, "outer.obj")) else None
    uv = np.load(os.path.join(d, "uv.npy")) if os.path.exists(os.path.join(d, "uv.npy")) else None
    return Asset(model.astype(np.float32), world.astype(np.float32), len(outer), faces, uv)


def checkpoint_path(model_folder: str, name: str) -> str:
    """sim.py:57-61: model_<min_loss_index>.pth, index from min_loss_index.npy, 10000 if that file is missing."""
    d = os.path.join(model_folder, name)
    try:
        k = int(np.load(os.path.join(d, "min_loss_index.npy")))
    except Exception:
        k = 10000
    return os.path.join(d, f"model_{k}.pth")
