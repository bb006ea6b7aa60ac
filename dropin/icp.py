"""`from icp import ICP` for the ROS nodes: put this directory in front of course_agv_slam/scripts
on PYTHONPATH and slam_ekf.py / localization.py pick up the GPU class."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from b2slam.icp import ICP as _GpuICP  # noqa: E402


class ICP(_GpuICP):
    """Constructed with no arguments by the nodes ([SLAM]:35); reads /icp/* like the original."""

    def __init__(self):
        _GpuICP.__init__(self, use_ros_params=True)
