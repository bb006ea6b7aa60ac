"""`from mapping import Mapping` for slam_ekf.py ([SLAM]:15,33): same constructor and update()."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from b2slam.mapping import Mapping  # noqa: E402,F401
