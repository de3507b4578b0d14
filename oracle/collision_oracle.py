"""CPU oracle for the collision half of the hot path.

TEST INFRASTRUCTURE ONLY — see the header of ``oracle/minsnap_oracle.py`` for who
may import ``oracle/``.

PARITY UNPINNED.  The reference delegates this arithmetic to FCL (C++) through the
``python-fcl`` binding (module ``fcl``; version not pinned anywhere in the
reference: no requirements file, not in package.xml:15-24, not in setup.py) and to
``numpy-stl`` for mesh ingest.  Neither is in ``/root/reference`` nor installed in
the build image, the reference records no expected collision answer anywhere
(src/RigidBodyPlanners/fcl_checker.py:124-138 prints the result of one query and
nothing keeps it), so this restatement cannot be checked against the reference's
own outputs.  What it follows:

* mesh ingest, rounding and pose conventions: the reference's own call sites,
  src/RigidBodyPlanners/fcl_checker.py:19-59,93-103 and
  src/RigidBodyPlanners/RB_planning_sep_coll_check.py:208-215;
* the published algorithm of FCL's mesh-mesh leaf test for a default
  ``CollisionRequest`` (no contacts requested): two BVH models collide iff some
  triangle pair intersects, and a pair is tested with the 17-axis separating-axis
  test of ``Intersect::intersect_Triangle`` / ``project6`` — both triangles
  translated by ``-P1``; axes = the two face normals, the nine edge x edge
  products, and the six edge x normal products; a pair is *separated* on an axis
  only if ``min1 > max2`` or ``min2 > max1`` (strict), so touching counts as
  collision.  FCL's bounding-volume hierarchy only prunes pairs; it does not
  change the answer away from the touching boundary.

The only weak anchor available: every state of the shipped planned
path was accepted by FCL as collision-free — ``tests/test_collision_oracle.py``
checks that, plus hand-constructed positives.

``margin`` (below) measures how far a pose is from the touching boundary so tests
can exclude the epsilon band the north star allows.
"""
from __future__ import annotations

import math
import struct

import numpy as np


# --------------------------------------------------------------------------- a10
def read_stl_triangles(path) -> np.ndarray:
    """Triangles of a binary or ASCII STL file as float32 ``[T, 3, 3]`` — the
    ``mesh.Mesh.from_file(...).vectors`` of numpy-stl used at
    src/RigidBodyPlanners/fcl_checker.py:20."""
    with open(path, "rb") as fh:
        data = fh.read()
    if len(data) >= 84:
        (count,) = struct.unpack_from("<I", data, 80)
        if 84 + 50 * count == len(data):
            rec = np.dtype([("n", "<f4", (3,)), ("v", "<f4", (3, 3)), ("a", "<u2")])
            return np.frombuffer(data, dtype=rec, count=count, offset=84)["v"].copy()
    tris, cur = [], []
    for line in data.decode("ascii", errors="replace").splitlines():
        parts = line.split()
        if len(parts) == 4 and parts[0] == "vertex":
            cur.append([float(parts[1]), float(parts[2]), float(parts[3])])
            if len(cur) == 3:
                tris.append(cur)
                cur = []
    return np.asarray(tris, dtype=np.float32).reshape(-1, 3, 3)


def ingest_mesh(vectors_f32: np.ndarray):
    """``Fcl_mesh.load_stl`` + ``create_indexed_triangles``
    (src/RigidBodyPlanners/fcl_checker.py:19-40): unique vertices, then BOTH the
    vertex table and the triangle corners rounded to 2 decimals *in float32*,
    then corner -> vertex index by exact equality.  Returns
    ``(verts[V,3] float64, tris[T,3] int64)``; FCL receives the float32 values
    widened to double."""
    vectors_f32 = np.asarray(vectors_f32, dtype=np.float32)
    flat = vectors_f32.reshape(-1, 3)
    verts = np.around(np.unique(flat, axis=0), 2)
    vecs = np.around(vectors_f32, 2)
    tris = np.zeros((len(vecs), 3), dtype=np.int64)
    for i, tri in enumerate(vecs):
        for j, p in enumerate(tri):
            (idx,) = np.where(np.all(p == verts, axis=1))
            if len(idx) != 1:
                # the reference assigns ``index[0]`` (an array) into a scalar slot
                # and raises for anything but exactly one match
                raise ValueError("rounded vertex does not match exactly one table entry")
            tris[i, j] = idx[0]
    return verts.astype(np.float64), tris


def mesh_triangles(vectors_f32) -> np.ndarray:
    """Triangle soup ``[T, 3, 3]`` float64 as FCL sees it after ``ingest_mesh``."""
    verts, tris = ingest_mesh(vectors_f32)
    return verts[tris]


# --------------------------------------------------------------------------- poses
def quat_to_matrix(q_xyzw):
    """Rotation matrix of the unit quaternion the reference hands to
    ``fcl.Transform(q_wxyz, T)`` (fcl_checker.py:54-59)."""
    x, y, z, w = [float(c) for c in q_xyzw]
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def pose_matrices(poses):
    """``poses[P, 4] = (x, y, z, yaw)`` (``isStateValid``,
    RB_planning_sep_coll_check.py:208-215: ``q = quaternion_from_euler(0,0,yaw)``)
    or ``poses[P, 7] = (x, y, z, qx, qy, qz, qw)`` -> ``(R[P,3,3], T[P,3])``."""
    poses = np.asarray(poses, dtype=np.float64)
    P, dim = poses.shape
    R = np.zeros((P, 3, 3))
    if dim == 4:
        half = poses[:, 3] / 2.0
        qz, qw = np.sin(half), np.cos(half)
        qx = qy = np.zeros(P)
    elif dim == 7:
        qx, qy, qz, qw = poses[:, 3], poses[:, 4], poses[:, 5], poses[:, 6]
    else:
        raise ValueError("pose_dim must be 4 or 7")
    R[:, 0, 0] = 1 - 2 * (qy * qy + qz * qz)
    R[:, 0, 1] = 2 * (qx * qy - qz * qw)
    R[:, 0, 2] = 2 * (qx * qz + qy * qw)
    R[:, 1, 0] = 2 * (qx * qy + qz * qw)
    R[:, 1, 1] = 1 - 2 * (qx * qx + qz * qz)
    R[:, 1, 2] = 2 * (qy * qz - qx * qw)
    R[:, 2, 0] = 2 * (qx * qz - qy * qw)
    R[:, 2, 1] = 2 * (qy * qz + qx * qw)
    R[:, 2, 2] = 1 - 2 * (qx * qx + qy * qy)
    return R, poses[:, :3].copy()


# --------------------------------------------------------------------------- a11
def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _dot(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def sat_pair(p, q):
    """17-axis separating-axis test of triangles ``p[..., 3, 3]`` vs
    ``q[..., 3, 3]`` (broadcast over leading dims), FCL ``intersect_Triangle`` /
    ``project6`` semantics.  Returns ``(hit, gap)``: ``hit`` bool, ``gap`` the
    largest normalised separation over the axes (<= 0 when intersecting; the
    distance scale of the epsilon band)."""
    o = p[..., 0, :]
    p1 = p[..., 0, :] - o
    p2 = p[..., 1, :] - o
    p3 = p[..., 2, :] - o
    q1 = q[..., 0, :] - o
    q2 = q[..., 1, :] - o
    q3 = q[..., 2, :] - o
    e1, e2, e3 = p2 - p1, p3 - p2, p1 - p3
    f1, f2, f3 = q2 - q1, q3 - q2, q1 - q3
    n1 = _cross(e1, e2)
    m1 = _cross(f1, f2)
    axes = [n1, m1]
    for e in (e1, e2, e3):
        for f in (f1, f2, f3):
            axes.append(_cross(e, f))
    axes += [_cross(e1, n1), _cross(e2, n1), _cross(e3, n1),
             _cross(f1, m1), _cross(f2, m1), _cross(f3, m1)]
    shape = np.broadcast(p[..., 0, 0], q[..., 0, 0]).shape
    separated = np.zeros(shape, dtype=bool)
    gap = np.full(shape, -np.inf)
    for ax in axes:
        a1, a2, a3 = _dot(ax, p1), _dot(ax, p2), _dot(ax, p3)
        b1, b2, b3 = _dot(ax, q1), _dot(ax, q2), _dot(ax, q3)
        mx1 = np.maximum(np.maximum(a1, a2), a3)
        mn1 = np.minimum(np.minimum(a1, a2), a3)
        mx2 = np.maximum(np.maximum(b1, b2), b3)
        mn2 = np.minimum(np.minimum(b1, b2), b3)
        separated |= (mn1 > mx2) | (mn2 > mx1)
        length = np.sqrt(_dot(ax, ax))
        raw = np.maximum(mn1 - mx2, mn2 - mx1)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.where(length > 0, raw / length, -np.inf)
        gap = np.maximum(gap, g)
    return ~separated, gap


def collide_poses(robot_tris, env_tris, poses, with_margin=False):
    """``Fcl_checker.check_collision`` (fcl_checker.py:93-100) for a batch of
    robot poses; the environment stays at identity.  ``robot_tris[Tr,3,3]`` and
    ``env_tris[Te,3,3]`` are float64 triangle soups (``mesh_triangles``).
    Returns ``hit[P]`` uint8 and, when asked, ``margin[P]`` = the smallest over
    triangle pairs of the pair's largest axis gap: negative inside a collision,
    positive when free, ~0 at touching."""
    robot_tris = np.asarray(robot_tris, dtype=np.float64)
    env_tris = np.asarray(env_tris, dtype=np.float64)
    R, T = pose_matrices(poses)
    P = R.shape[0]
    # world-frame robot corners: R v + T
    world = np.einsum("pij,tcj->ptci", R, robot_tris) + T[:, None, None, :]
    hit = np.zeros(P, dtype=bool)
    margin = np.full(P, np.inf)
    for r in range(robot_tris.shape[0]):
        for e in range(env_tris.shape[0]):
            h, g = sat_pair(world[:, r], env_tris[e][None])
            hit |= h
            margin = np.minimum(margin, g)
    if with_margin:
        return hit.astype(np.uint8), margin
    return hit.astype(np.uint8)


def check_collision(robot_tris, env_tris, T, q_xyzw=(0, 0, 0, 1)):
    """Single query, the reference's call shape."""
    pose = np.array([[T[0], T[1], T[2], q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]]])
    return int(collide_poses(robot_tris, env_tris, pose)[0])


def yaw_pose_quat(yaw):
    """xyzw quaternion of ``quaternion_from_euler(0, 0, yaw)``."""
    return (0.0, 0.0, math.sin(yaw / 2.0), math.cos(yaw / 2.0))
