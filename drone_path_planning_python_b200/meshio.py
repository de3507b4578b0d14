"""Host-side mesh ingest: what ``Fcl_mesh.load_stl`` / ``create_indexed_triangles`` do
before FCL sees the mesh (src/RigidBodyPlanners/fcl_checker.py:19-40).  Pure file parsing
and rounding of at most a few dozen triangles — one-off setup, not part of the hot path.
"""
from __future__ import annotations

import os
import struct

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "stl_meshes.npz")


def read_stl(path) -> np.ndarray:
    """Triangles ``[T, 3, 3]`` float32 of a binary or ASCII STL file (the ``vectors`` array
    numpy-stl returns at fcl_checker.py:20)."""
    with open(path, "rb") as fh:
        blob = fh.read()
    if len(blob) >= 84:
        (count,) = struct.unpack_from("<I", blob, 80)
        if len(blob) == 84 + 50 * count:
            out = np.empty((count, 3, 3), dtype=np.float32)
            for i in range(count):
                out[i] = np.frombuffer(blob, dtype="<f4", count=9, offset=84 + 50 * i + 12).reshape(3, 3)
            return out
    tris = []
    corner = []
    for raw in blob.decode("ascii", errors="replace").splitlines():
        tok = raw.split()
        if len(tok) == 4 and tok[0].lower() == "vertex":
            corner.append((float(tok[1]), float(tok[2]), float(tok[3])))
            if len(corner) == 3:
                tris.append(corner)
                corner = []
    if not tris:
        raise ValueError("%s: not a binary or ASCII STL file" % path)
    return np.asarray(tris, dtype=np.float32)


def write_stl(path, triangles, header=b"drone_path_planning_python_b200") -> None:
    """Binary STL writer (normals left zero, as viewers recompute them)."""
    tri = np.asarray(triangles, dtype="<f4").reshape(-1, 3, 3)
    with open(path, "wb") as fh:
        fh.write(header[:80].ljust(80, b"\0"))
        fh.write(struct.pack("<I", tri.shape[0]))
        for t in tri:
            fh.write(struct.pack("<3f", 0.0, 0.0, 0.0))
            fh.write(t.tobytes())
            fh.write(b"\0\0")


def ingest_mesh(vectors):
    """Vertex table + indexed triangles with the reference's rounding: unique corners,
    both the table and the corners rounded to 2 decimals in float32, indices by exact match
    (fcl_checker.py:21-37).  Returns ``(verts[V,3] float32, vecs[T,3,3] float32,
    tris[T,3] float64)`` — ``tris`` is float-typed exactly as the reference's."""
    vectors = np.asarray(vectors, dtype=np.float32).reshape(-1, 3, 3)
    verts = np.around(np.unique(vectors.reshape(-1, 3), axis=0), 2)
    vecs = np.around(vectors, 2)
    tris = np.zeros((len(vecs), 3))
    for i, tri in enumerate(vecs):
        for j, corner in enumerate(tri):
            match = np.flatnonzero((verts == corner).all(axis=1))
            if match.size != 1:
                raise ValueError("corner %s of triangle %d matches %d vertices after rounding"
                                 % (corner, i, match.size))
            tris[i, j] = match[0]
    return verts, vecs, tris


def triangle_soup(verts, tris) -> np.ndarray:
    """``[T, 3, 3]`` float64 corners — the geometry handed to the device."""
    return np.asarray(verts, dtype=np.float64)[np.asarray(tris, dtype=np.int64)]


def shipped_mesh_names():
    with np.load(_DATA) as z:
        return sorted(z.files)


def shipped_mesh(name: str) -> np.ndarray:
    """Raw float32 triangles of one of the reference's ``resources/stl`` meshes
    (e.g. ``"env-scene-ltu-experiment"``, ``"custom_triangle_robot"``)."""
    with np.load(_DATA) as z:
        return z[name].copy()
