"""B200 drop-in for ``trajectory_visualising.visualization``
(reference: src/trajectory_visualising/visualization.py).

``get_nav_path_msg(tr, timestep, offset)`` keeps the reference's contract (:39-71): one
``PoseStamped`` per sample time ``np.arange(0, tr.duration, timestep)``, position = flat-output
position + offset, orientation = ``quaternion_from_euler(0, 0, -yaw)`` (the yaw sign flip of
:62-63), frame ``world`` — but the whole trajectory is sampled by ONE ``mst_flat_outputs`` launch
(``Trajectory.eval_many``) instead of a Python loop of ``tr.eval(t)`` calls.
"""
import math

import numpy as np

try:  # pragma: no cover - only on a ROS machine
    import rospy
    from geometry_msgs.msg import PoseStamped
    from nav_msgs.msg import Path
    _HAVE_ROS = True
except ImportError:  # ROS message stand-ins: same attribute structure, no transport
    from types import SimpleNamespace
    rospy = None
    _HAVE_ROS = False

    def PoseStamped():
        return SimpleNamespace(header=SimpleNamespace(frame_id="", stamp=None),
                               pose=SimpleNamespace(position=SimpleNamespace(x=0.0, y=0.0, z=0.0),
                                                    orientation=SimpleNamespace(x=0.0, y=0.0, z=0.0, w=1.0)))

    def Path():
        return SimpleNamespace(header=SimpleNamespace(frame_id="", stamp=None), poses=[])

try:
    from .uav_trajectory import *  # noqa: F401,F403  (same import dance as the reference, :8-11)
    from .uav_trajectory import Trajectory, TrajectoryOutput
except ImportError:
    from uav_trajectory import *  # noqa: F401,F403
    from uav_trajectory import Trajectory, TrajectoryOutput  # noqa: F401


def sample_times(tr: Trajectory, timestep: float) -> np.ndarray:
    """The reference's sampling grid: ``np.arange(0, tr.duration, timestep)`` (:53)."""
    return np.arange(0, tr.duration, timestep)


def sample_path(tr: Trajectory, timestep: float, offset=(0, 0, 0)):
    """Batched core of ``get_nav_path_msg``: ``(positions[S, 3], quaternions_xyzw[S, 4])`` of every
    sample, one kernel launch.  Quaternion = ``tf.transformations.quaternion_from_euler(0, 0, -yaw)``,
    i.e. ``(0, 0, sin(-yaw/2), cos(-yaw/2))``."""
    ts = sample_times(tr, timestep)
    rows = tr.eval_many(ts)                                   # [S, 13] pos vel acc omega yaw
    pos = rows[:, 0:3] + np.asarray(offset, dtype=np.float64).reshape(1, 3)
    # tf halves the angle and calls math.sin / math.cos (libm), so the same calls are used here
    quat = np.array([[0.0, 0.0, math.sin(-yaw / 2.0), math.cos(-yaw / 2.0)] for yaw in rows[:, 12]]).reshape(-1, 4)
    return pos, quat


def visualize_python(tr: Trajectory, timestep: float):
    """3-D line plot of the sampled positions (reference :17-36); matplotlib imported lazily."""
    import matplotlib.pyplot as plt
    from mpl_toolkits.mplot3d import Axes3D  # noqa: F401

    size = int(tr.duration / timestep + 0.5)
    print("size:", size)
    pos, _ = sample_path(tr, timestep)
    for row in pos:
        print(row[0], row[1], row[2])
    fig = plt.figure()
    ax = fig.add_subplot(111, projection='3d')
    ax.plot(pos[:, 0], pos[:, 1], pos[:, 2])
    plt.show()


def get_nav_path_msg(tr: Trajectory, timestep: float, offset=[0, 0, 0]):
    """
    Publish the ROS message containing the waypoints
    """
    msg = Path()
    msg.header.frame_id = "world"
    if _HAVE_ROS:
        msg.header.stamp = rospy.Time.now()

    size = int(tr.duration / timestep + 0.5)
    print("size:", size)
    pos, quat = sample_path(tr, timestep, offset)
    for p, q in zip(pos, quat):
        pose = PoseStamped()
        pose.pose.position.x, pose.pose.position.y, pose.pose.position.z = float(p[0]), float(p[1]), float(p[2])
        pose.pose.orientation.x = float(q[0])
        pose.pose.orientation.y = float(q[1])
        pose.pose.orientation.z = float(q[2])
        pose.pose.orientation.w = float(q[3])
        msg.poses.append(pose)

    if _HAVE_ROS:
        rospy.loginfo("Published {} waypoints.".format(len(msg.poses)))
    return msg


def quaternion_from_yaw(yaw: float):
    """``tf.transformations.quaternion_from_euler(0, 0, yaw)`` (xyzw) for callers without tf."""
    return [0.0, 0.0, math.sin(yaw * 0.5), math.cos(yaw * 0.5)]
