"""ROS message stand-ins so the node scripts import and run without a ROS installation
(tests, batch jobs).  When rospy / geometry_msgs / nav_msgs are importable the real classes
are used and nothing here matters."""
from types import SimpleNamespace

try:  # pragma: no cover - exercised only on a ROS machine
    import rospy  # noqa: F401
    from geometry_msgs.msg import Point, PoseStamped, Quaternion  # noqa: F401
    from nav_msgs.msg import Path  # noqa: F401
    HAVE_ROS = True
except ImportError:
    rospy = None
    HAVE_ROS = False

    def Point(x=0.0, y=0.0, z=0.0):
        return SimpleNamespace(x=x, y=y, z=z)

    def Quaternion(x=0.0, y=0.0, z=0.0, w=1.0):
        return SimpleNamespace(x=x, y=y, z=z, w=w)

    def PoseStamped():
        return SimpleNamespace(header=SimpleNamespace(frame_id="", stamp=None),
                               pose=SimpleNamespace(position=Point(), orientation=Quaternion()))

    def Path():
        return SimpleNamespace(header=SimpleNamespace(frame_id="", stamp=None), poses=[])


def make_pose(x, y, z, q_xyzw, frame_id="world"):
    ps = PoseStamped()
    ps.header.frame_id = frame_id
    ps.pose.position = Point(float(x), float(y), float(z))
    ps.pose.orientation = Quaternion(float(q_xyzw[0]), float(q_xyzw[1]), float(q_xyzw[2]), float(q_xyzw[3]))
    return ps
