#!/usr/bin/env python3
"""B200 drop-in for scripts/drones_pols_generator.py: drone path -> polynomial pieces.

Same module-level names as the reference (``callback1``, ``callback2``, ``path_to_pol``,
``listener``, ``piece_pols_pub``).  ``path_to_pol`` keeps the reference's conventions
(:40-90): uniform stamps ``t_i = i * (10 / len(poses))``, yaw from the pose quaternion, a
float32 ``(n, 33)`` matrix ``[T | x | y | z | yaw]`` written with ``np.savetxt`` and published
as ``TrajectoryPolynomialPieceMarios`` — but yaw extraction, the solve and the float32 packing
all run on the GPU (``mst_formation_waypoints`` with a zero offset, ``mst_solve_batch``,
``mst_pack_pol_matrix``).  ``paths_to_matrices`` is the batched form.
"""
import os

import numpy as np

import drone_path_planning_python_b200 as _mst

from ._ros_compat import HAVE_ROS, Path, rospy
from .drones_traj_generator import _path_array

try:  # the custom message lives in the user's workspace (reference :11-14)
    from execution.msg import TrajectoryPolynomialPieceMarios
except ImportError:
    try:
        from crazyswarm.msg import TrajectoryPolynomialPieceMarios
    except ImportError:
        from types import SimpleNamespace

        def TrajectoryPolynomialPieceMarios():
            return SimpleNamespace(cf_id=0, poly_x=[], poly_y=[], poly_z=[], poly_yaw=[], durations=[])

TOTAL_DURATION = 10  # secs (reference :44)
# the reference hard-codes the author's home directory (:79); overridable here
OUTPUT_DIR = os.environ.get("DRONE_POL_MATRIX_DIR",
                            "/home/marios/thesis_ws/src/drone_path_planning/resources/trajectories/")


def callback1(path):
    if callback1.counter == 0:
        path_to_pol(path, 1)
        callback1.counter += 1


callback1.counter = 0


def callback2(path):
    if callback2.counter == 0:
        path_to_pol(path, 2)
        callback2.counter += 1


callback2.counter = 0


def paths_to_matrices(poses, total_duration=TOTAL_DURATION):
    """``poses[B, m, 7]`` (x, y, z, qx, qy, qz, qw) -> float32 ``[B, m-1, 33]`` matrices.
    One launch each for yaw extraction, the min-snap solve of all B x 4 axes, and the packing."""
    poses = np.asarray(poses, dtype=np.float64)
    B, m, _ = poses.shape
    wp = _mst.formation_waypoints(poses, np.zeros((1, 3)), K=4)          # [B, m, 4] with yaw
    step = total_duration / m                                           # divides by the pose count (:46)
    stamps = np.array([[step * i for i in range(m)]])                   # shared by every path
    stamps = np.repeat(stamps, B, axis=0)
    coef, dur, info = _mst.solve_batch(wp, stamps)
    if int((info != 0).sum()) != 0:
        raise np.linalg.LinAlgError("Singular matrix")
    return _mst.pack_pol_matrix(coef, dur).cpu().numpy()


def paths_to_csv(poses, directory, first_id=1, total_duration=TOTAL_DURATION):
    """Batched ``path_to_pol`` file output: ``poses[B, m, 7]`` -> ``Pol_matrix_{first_id + b}.csv`` in
    ``directory``, each byte-identical to what ``np.savetxt(..., delimiter=",")`` writes for that
    drone's float32 matrix; solve, packing and text formatting are one launch each for all B."""
    matrices = paths_to_matrices(poses, total_duration)
    files = _mst.pol_matrix_csv(matrices)
    names = []
    for b, blob in enumerate(files):
        name = os.path.join(directory, "Pol_matrix_{}.csv".format(first_id + b))
        with open(name, "wb") as fh:
            fh.write(blob)
        names.append(name)
    return matrices, names


def path_to_pol(path, cfid: int):
    print("Path received...")
    matrix = paths_to_matrices(_path_array(path)[None])[0]

    try:
        # the bytes np.savetxt(file, matrix, delimiter=",") writes ('%.18e' fields), formatted on the GPU
        with open(os.path.join(OUTPUT_DIR, "Pol_matrix_{}.csv".format(cfid)), "wb") as fh:
            fh.write(_mst.pol_matrix_csv(matrix)[0])
    except OSError as exc:  # the reference would crash on a machine without that directory
        print("could not write Pol_matrix_{}.csv: {}".format(cfid, exc))

    pol_to_send = TrajectoryPolynomialPieceMarios()
    pol_to_send.cf_id = cfid
    pol_to_send.poly_x = list(matrix[:, 1:9].flatten())
    pol_to_send.poly_y = list(matrix[:, 9:17].flatten())
    pol_to_send.poly_z = list(matrix[:, 17:25].flatten())
    pol_to_send.poly_yaw = list(matrix[:, 25:33].flatten())
    pol_to_send.durations = list(matrix[:, 0].flatten())

    if piece_pols_pub is not None:
        piece_pols_pub.publish(pol_to_send)
        print("Published polynomial piece...")
    return matrix, pol_to_send


def listener():
    rospy.init_node('drones_path_listener')
    rospy.Subscriber('drone1Path', Path, callback1)
    rospy.Subscriber('drone2Path', Path, callback2)
    rospy.spin()


piece_pols_pub = rospy.Publisher('piece_pol', TrajectoryPolynomialPieceMarios, queue_size=10) if HAVE_ROS else None

if __name__ == '__main__':
    listener()
