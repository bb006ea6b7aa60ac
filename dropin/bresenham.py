"""`import bresenham as drawing` ([MAP]:4): drawing.bresenham(start, end).path on the GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from b2slam.bresenham import bresenham, paths  # noqa: E402,F401
