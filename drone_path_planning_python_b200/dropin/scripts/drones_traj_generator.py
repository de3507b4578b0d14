#!/usr/bin/env python3
"""B200 drop-in for scripts/drones_traj_generator.py: rigid-body path -> per-drone paths.

Same module-level names as the reference (``drone_positions``, ``drone_pose``,
``drone_pose2``, ``get_drone_positions``, ``transform``, ``callback``, ``listener``,
``trajPub1``, ``trajPub2``).  ``transform(path)`` no longer loops over poses calling
``tf2_geometry_msgs.do_transform_pose`` (reference :56-89): all poses x all drone offsets go
through ``mst_formation_waypoints`` in one launch; ``transform_many`` does the same for a whole
batch of rigid-body paths and any number of drones.
"""
import numpy as np

import drone_path_planning_python_b200 as _mst

from ._ros_compat import HAVE_ROS, Path, make_pose, rospy

drone_positions = [
    [0.5, 0, 0],
    [-0.5, 0, 0]
]

drone_pose = make_pose(drone_positions[0][0], drone_positions[0][1], drone_positions[0][2], (0, 0, 0, 1), 'rb_path')
drone_pose2 = make_pose(drone_positions[1][0], drone_positions[1][1], drone_positions[1][2], (0, 0, 0, 1), 'rb_path')


def get_drone_positions(drone_positions):
    """One offset pose per drone (the reference's version, :41-53, appends the same object
    repeatedly and is never called; this one returns distinct poses)."""
    return [make_pose(p[0], p[1], p[2], (0, 0, 0, 1), 'rb_path') for p in drone_positions]


def _path_array(path):
    """nav_msgs/Path -> [m, 7] (x, y, z, qx, qy, qz, qw)."""
    out = np.zeros((len(path.poses), 7))
    for i, ps in enumerate(path.poses):
        p, q = ps.pose.position, ps.pose.orientation
        out[i] = (p.x, p.y, p.z, q.x, q.y, q.z, q.w)
    return out


def transform_many(rb_poses, offsets=drone_positions):
    """``rb_poses[F, m, 7]`` (or ``[F, m, 4]`` = x, y, z, yaw) -> ``[F * D, m, 4]`` drone
    waypoints (x, y, z, yaw), ready for ``solve_batch(share_time_group=D)``."""
    return _mst.formation_waypoints(rb_poses, np.asarray(offsets, dtype=np.float64), K=4)


def transform(path, inverse=False):
    """Both drones' paths for one rigid-body path (reference :56-89).  Every output pose keeps
    the rigid body's orientation (q_rb x identity) and ``p = R(q_rb) @ offset + t_rb``."""
    rb = _path_array(path)
    wp = transform_many(rb[None], drone_positions).cpu().numpy()   # [2, m, 4]
    paths = []
    for d in range(len(drone_positions)):
        out = Path()
        out.header.frame_id = "world"
        if HAVE_ROS:
            out.header.stamp = rospy.get_rostime()
        for i in range(rb.shape[0]):
            out.poses.append(make_pose(wp[d, i, 0], wp[d, i, 1], wp[d, i, 2], rb[i, 3:7], "world"))
        paths.append(out)
    return paths[0], paths[1]


def callback(path):
    print("Path received...")
    print(len(path.poses))
    drone_path1, drone_path2 = transform(path)

    trajPub1.publish(drone_path1)
    trajPub2.publish(drone_path2)


def listener():
    rospy.init_node('rb_path_listener', anonymous=True)
    rospy.Subscriber('rigiBodyPath', Path, callback)
    rospy.spin()


trajPub1 = rospy.Publisher('drone1Path', Path, queue_size=10) if HAVE_ROS else None
trajPub2 = rospy.Publisher('drone2Path', Path, queue_size=10) if HAVE_ROS else None
if __name__ == '__main__':
    listener()
