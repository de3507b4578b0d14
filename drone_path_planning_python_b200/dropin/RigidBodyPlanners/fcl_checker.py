"""B200 drop-in for ``RigidBodyPlanners.fcl_checker``
(reference: src/RigidBodyPlanners/fcl_checker.py).

``Fcl_mesh`` / ``Fcl_checker`` keep the reference's constructor arguments, attributes
(``verts``, ``vecs``, ``tris``, ``collision_object``) and methods, so
``PlannerSepCollision.isStateValid`` (RB_planning_sep_coll_check.py:208-226) runs unchanged —
but there is no python-fcl underneath: the meshes live on the GPU and every query is answered
by ``mst_collide_poses``.  ``check_collision_batch`` is the extra, batched entry point (all
interpolated states of a motion in one launch).
"""
import numpy as np

import drone_path_planning_python_b200 as _mst
from drone_path_planning_python_b200 import meshio as _meshio


class Fcl_mesh():
    """reference: fcl_checker.py:13-59"""

    def __init__(self, filename) -> None:
        self.load_stl(filename)
        self.create_indexed_triangles(self.verts, self.vecs)
        self.create_fcl_mesh()

    def load_stl(self, filename):
        # unique vertices and corners, both rounded to 2 decimals in float32 (:20-23)
        vectors = _meshio.read_stl(filename)
        verts, vecs, _ = _meshio.ingest_mesh(vectors)
        self.verts, self.vecs = verts, vecs
        return verts, vecs

    def create_indexed_triangles(self, vertices, vectors):
        """Indexed triangle mesh from a vertex table and the triangle corners (:28-40)."""
        tris = np.zeros([len(vectors), 3])
        for i, tri in enumerate(vectors):
            for j, corner in enumerate(tri):
                match = np.flatnonzero(np.all(corner == vertices, axis=1))
                if match.size != 1:
                    raise ValueError("setting an array element with a sequence.")  # what numpy raises in the reference
                tris[i][j] = match[0]
        self.tris = tris
        return tris

    def create_fcl_mesh(self):
        """Upload the mesh; the returned object plays the role of fcl.BVHModel (:42-52)."""
        soup = _meshio.triangle_soup(self.verts, self.tris)
        self.m = _mst.Mesh(soup)
        self._pose = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])   # (x, y, z, qx, qy, qz, qw)
        self.collision_object = self
        return self.m

    def set_transform(self, T=[0, 0, 0], q=[0, 0, 0, 1]):
        # q arrives as xyzw like in the reference, which reorders it for fcl (:54-59)
        pose = self._pose
        pose[0], pose[1], pose[2] = T[0], T[1], T[2]
        pose[3], pose[4], pose[5], pose[6] = q[0], q[1], q[2], q[3]

    @property
    def T(self):
        return self._pose[:3]

    @property
    def q(self):
        return self._pose[3:]

    def pose(self):
        return self._pose


def visualize_meshes(filenames):
    """Plot STL files (reference :62-82); matplotlib imported lazily."""
    from matplotlib import pyplot as plt
    from mpl_toolkits.mplot3d.art3d import Poly3DCollection

    ax = plt.figure().add_subplot(projection="3d")
    corners = []
    for filename in filenames:
        triangles = _meshio.read_stl(filename)
        corners.append(triangles.reshape(-1, 3))
        ax.add_collection3d(Poly3DCollection(triangles, alpha=0.6))
    if corners:   # collections do not autoscale the axes
        pts = np.concatenate(corners)
        for setter, lo, hi in zip((ax.set_xlim, ax.set_ylim, ax.set_zlim), pts.min(0), pts.max(0)):
            setter(lo, hi)
    for label, setter in zip("XYZ", (ax.set_xlabel, ax.set_ylabel, ax.set_zlabel)):
        setter(label)
    plt.show()


class Fcl_checker():
    """reference: fcl_checker.py:85-103"""

    def __init__(self, env_mesh_file, robot_mesh_file) -> None:
        self.env = Fcl_mesh(env_mesh_file)
        self.robot = Fcl_mesh(robot_mesh_file)
        self.request = None   # fcl.CollisionRequest() in the reference: default request, no contacts
        self.result = None
        # the single query, bound once: mst_collide_pose_sync(robot, env, pose[7], 7, &hit)
        import ctypes
        lib = _mst._abi.load()
        _mst._abi.require_cuda()
        out = ctypes.c_int(-1)
        fn, rh, eh, ref = lib.mst_collide_pose_sync, self.robot.m.handle, self.env.m.handle, ctypes.byref(out)

        def query(pose, _fn=fn, _rh=rh, _eh=eh, _ref=ref, _out=out, _check=_mst._abi.check):
            rc = _fn(_rh, _eh, pose.ctypes.data, 7, _ref)
            if rc != 0:
                _check(rc, "mst_collide_pose_sync")
            return _out.value
        self._query = query

    def check_collision(self, T=None, q=[0, 0, 0, 1]):
        if T is not None:
            self.robot.set_transform(T, q)
        # one launch + one stream synchronisation (mst_collide_pose_sync); fcl.collide returns the
        # number of contacts: 0 or 1 for the default request
        return self._query(self.robot._pose)

    def set_robot_transform(self, T, q=[0, 0, 0, 1]):
        self.robot.set_transform(T, q)

    def check_collision_batch(self, poses):
        """``poses[P, 4]`` (x, y, z, yaw — ``isStateValid``'s state) or ``[P, 7]``
        (x, y, z, qx, qy, qz, qw) -> uint8 numpy array of collision flags, one launch."""
        return _mst.collide_poses(self.robot.m, self.env.m, np.asarray(poses, dtype=np.float64)).cpu().numpy()

    def check_motions(self, states_a, states_b, steps):
        """Batched motion validation: ``states_a[M, 4]`` -> ``states_b[M, 4]`` (x, y, z, yaw), each
        checked at ``steps`` linearly interpolated states (end state included).  Returns a bool
        numpy array, True where the motion is collision-free."""
        bad = _mst.collide_motions(self.robot.m, self.env.m, np.asarray(states_a, dtype=np.float64),
                                   np.asarray(states_b, dtype=np.float64), steps)
        return ~bad.cpu().numpy().astype(bool)
