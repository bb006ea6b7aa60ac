"""b2slam: B200-native ICP scan matching and occupancy-grid mapping (import as `b2slam`).

Drop-in replacements for the reference's hot-path classes
(`course_agv_slam/scripts/icp.py`, `mapping.py`, `bresenham.py`); see DESIGN.md.
Submodules are imported lazily so that `import b2slam.synth` works without the CUDA library.
`b2slam.bresenham` is the module (the reference does `import bresenham as drawing`).
"""
__version__ = "0.1.0"


def __getattr__(name):
    if name == "ICP":
        from b2slam.icp import ICP
        return ICP
    if name == "Mapping":
        from b2slam.mapping import Mapping
        return Mapping
    raise AttributeError(name)
